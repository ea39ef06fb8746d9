"""Shared test utilities: golden-fixture loading and oracle drivers."""
import glob
import os

import numpy as np
import torch

from oracle import view_synthesis as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _fid(s):
    return "s" if s == "s" else int(s)


class Golden:
    """One committed fixture produced by tests/golden/make_golden.py from the reference."""

    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.frame_ids = [_fid(str(s)) for s in z["frame_ids"]]
        fl = [bool(int(x)) for x in z["flags"]] + [False, False, False]
        self.avg_reprojection, self.disable_automasking, self.no_ssim, self.v1_multiscale, self.posecnn = fl[:5]
        self.predictive_mask = fl[5]
        self.B, _, self.H, self.W = z["in__color__0__0"].shape
        self.n_src = len(self.frame_ids) - 1
        self.n_id = 0 if self.disable_automasking else (1 if self.avg_reprojection else self.n_src)
        # --scales (options.py:64): the fixtures of the default 0 1 2 3 carry no "scales" entry
        self.scales = [int(s) for s in z["scales"]] if "scales" in z.files else [0, 1, 2, 3]

    def cfg(self, **kw):
        return O.OracleConfig(height=self.H, width=self.W, frame_ids=tuple(self.frame_ids),
                              avg_reprojection=self.avg_reprojection,
                              disable_automasking=self.disable_automasking, no_ssim=self.no_ssim,
                              v1_multiscale=self.v1_multiscale, posecnn=self.posecnn,
                              predictive_mask=self.predictive_mask, scales=tuple(self.scales), **kw)

    def t(self, key, dtype=torch.float32):
        return torch.from_numpy(np.asarray(self.z[key])).to(dtype)

    def inputs(self, dtype=torch.float32):
        d = {}
        for k in self.z.files:
            if not k.startswith("in__"):
                continue
            parts = k[4:].split("__")
            if parts[0] == "color":
                key = ("color", _fid(parts[1]), int(parts[2]))
            elif parts[0] in ("K", "inv_K"):
                key = (parts[0], int(parts[1]))
            else:
                key = parts[0]
            d[key] = self.t(k, dtype)
        return d

    def noise(self, dtype=torch.float32):
        return [self.t("noise__%d" % s, dtype)[:, :max(self.n_id, 1)] for s in self.scales]

    def leaves(self, dtype=torch.float32):
        """disp_s and (axisangle, translation) leaves with requires_grad."""
        lv = {}
        for s in self.scales:
            lv[("disp", s)] = self.t("disp__%d" % s, dtype).requires_grad_(True)
        if self.predictive_mask:
            for s in self.scales:
                lv[("mask", s)] = self.t("mask__%d" % s, dtype).requires_grad_(True)
        for f in self.frame_ids[1:]:
            if f == "s":
                continue
            lv[("axisangle", f)] = self.t("axisangle__%s" % f, dtype).requires_grad_(True)
            lv[("translation", f)] = self.t("translation__%s" % f, dtype).requires_grad_(True)
        return lv


def run_oracle(g: Golden, dtype=torch.float32, **cfgkw):
    """Oracle fwd+bwd on a golden case; returns dict with losses, grads, side outputs."""
    cfg = g.cfg(**cfgkw)
    inputs = g.inputs(dtype)
    lv = g.leaves(dtype)
    outs = {}
    for s in g.scales:
        outs[("disp", s)] = lv[("disp", s)]
    for f in g.frame_ids[1:]:
        if f == "s":
            continue
        T = O.transformation_from_parameters(lv[("axisangle", f)], lv[("translation", f)], invert=(f < 0))
        T.retain_grad()
        outs[("cam_T_cam", 0, f)] = T
        outs[("axisangle", 0, f)] = lv[("axisangle", f)].reshape(-1, 1, 1, 3)
        outs[("translation", 0, f)] = lv[("translation", f)].reshape(-1, 1, 1, 3)
    if g.predictive_mask:
        outs["predictive_mask"] = {("disp", s): lv[("mask", s)] for s in g.scales}
    O.generate_images_pred(inputs, outs, cfg)
    for s in g.scales:
        outs[("depth", 0, s)].retain_grad()
    losses = O.compute_losses(inputs, outs, cfg, g.noise(dtype))
    losses["loss"].backward()
    return dict(losses=losses, outs=outs, leaves=lv, inputs=inputs, cfg=cfg)


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-300))
