"""Generate golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

For every case it imports /root/reference/trainer.py (after stubbing the three
absent, irrelevant modules tensorboardX / IPython / skimage),
binds ``Trainer.generate_images_pred``, ``compute_reprojection_loss`` and
``compute_losses`` onto a namespace built like trainer.py:145-159 with
``device = cpu`` (what ``--no_cuda`` selects, trainer.py:41), runs them plus
``losses["loss"].backward()`` on a seeded synthetic batch
(monodepth2_b200/synthetic.py) and stores inputs and results in
``tests/golden/<case>.npz``.  The tie-break noise of trainer.py:468-469 is drawn
beforehand and injected by temporarily replacing ``torch.randn``.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

CASES = {
    # name: dict(B,H,W, frame_ids, kind, seed, flags, jitter)
    "mono_iid": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="iid", seed=0, flags=[]),
    "mono_structured": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=5, flags=[]),
    "mono_jitterK": dict(B=2, H=64, W=96, frame_ids=[0, -1, 1], kind="iid", seed=1, flags=[], jitter=True),
    "stereo_iid": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="iid", seed=2, flags=["--use_stereo"]),
    "stereo_structured": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=6,
                              flags=["--use_stereo"]),
    "avg_reprojection": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=3,
                             flags=["--avg_reprojection"]),
    "disable_automasking": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=4,
                                flags=["--disable_automasking"]),
    "no_ssim": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=8, flags=["--no_ssim"]),
    "v1_multiscale": dict(B=2, H=64, W=96, frame_ids=[0, -1, 1], kind="structured", seed=9,
                          flags=["--v1_multiscale"]),
    "posecnn": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=10,
                    flags=["--pose_model_type", "posecnn"]),
    "predictive_mask": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=12,
                            flags=["--disable_automasking", "--predictive_mask"]),
    # four source frames (options.py:80-84 takes any --frame_ids; the maximum the C ABI supports: MD2_MAX_SRC)
    "five_frames": dict(B=2, H=32, W=64, frame_ids=[0, -2, -1, 1, 2], kind="structured", seed=13, flags=[]),
    # --scales subsets (options.py:64: any list of pyramid levels; trainer.py:345,413 iterate it, the dataloader
    # holds levels 0..3 regardless, trainer.py:127-135): md2_problem.scale_level.  Level 0 has to be in the list: the
    # reference builds backproject_depth / project_3d for opt.scales only and indexes [0] (trainer.py:151-159,377)
    "scales_0_2": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=14, flags=["--scales", "0", "2"]),
    "scales_0_1_3_stereo": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=15,
                                flags=["--scales", "0", "1", "3", "--use_stereo"]),
    "scales_0_2_posecnn": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=16,
                               flags=["--scales", "0", "2", "--pose_model_type", "posecnn"]),
    "scales_0_3_predictive_mask": dict(B=2, H=48, W=80, frame_ids=[0, -1, 1], kind="structured", seed=17,
                                       flags=["--scales", "0", "3", "--disable_automasking", "--predictive_mask"]),
    "stereo_only": dict(B=2, H=32, W=64, frame_ids=[0], kind="structured", seed=7,
                        flags=["--use_stereo", "--frame_ids", "0"]),
}


def import_reference():
    for name, attrs in (("tensorboardX", {"SummaryWriter": object}),
                        ("IPython", {"embed": lambda *a, **k: None}),
                        ("skimage", {}), ("skimage.transform", {})):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.path.insert(0, REF)
    import trainer as T  # noqa: the reference module, unmodified
    from options import MonodepthOptions
    return T, MonodepthOptions


def build_self(T, opt, dtype=torch.float32):
    """A namespace standing in for the Trainer instance (trainer.py:38-52,145-159)."""
    dev = torch.device("cpu")
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        s = SimpleNamespace(opt=opt, device=dev, num_scales=len(opt.scales))
        s.ssim = T.SSIM().to(dev)
        s.backproject_depth, s.project_3d = {}, {}
        for scale in opt.scales:
            h, w = opt.height // (2 ** scale), opt.width // (2 ** scale)
            s.backproject_depth[scale] = T.BackprojectDepth(opt.batch_size, h, w).to(dev).to(dtype)
            s.project_3d[scale] = T.Project3D(opt.batch_size, h, w).to(dev)
    finally:
        torch.set_default_dtype(prev)
    for name in ("generate_images_pred", "compute_reprojection_loss", "compute_losses"):
        setattr(s, name, types.MethodType(getattr(T.Trainer, name), s))
    return s


def run_reference(T, MonodepthOptions, case, dtype=torch.float32, batch=None):
    from monodepth2_b200.synthetic import make_batch
    argv = ["x", "--height", str(case["H"]), "--width", str(case["W"]),
            "--batch_size", str(case["B"]), "--no_cuda"] + list(case["flags"])
    if "--frame_ids" not in case["flags"]:
        argv += ["--frame_ids"] + [str(f) for f in case["frame_ids"]]
    old = sys.argv
    sys.argv = argv
    try:
        opt = MonodepthOptions().parse()
    finally:
        sys.argv = old
    if opt.use_stereo:
        opt.frame_ids.append("s")       # trainer.py:51-52
    frame_ids = list(opt.frame_ids)
    n_src = len(frame_ids) - 1
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else n_src)
    if batch is None:
        ms = bool(opt.v1_multiscale)
        batch = make_batch(case["B"], case["H"], case["W"], frame_ids, 4, case["seed"], case["kind"],
                           jitter_K=case.get("jitter", False), n_id=max(n_id, 1), all_scale_K=ms,
                           multiscale_noise=ms)
    inputs, outputs, pose, noise = batch
    inputs = {k: v.to(dtype) for k, v in inputs.items()}
    leaves = {}
    outs = {}
    scales = list(opt.scales)
    for s in scales:
        d = outputs[("disp", s)].to(dtype).clone().requires_grad_(True)
        leaves[("disp", s)] = d
        outs[("disp", s)] = d
    for f, (aa, tr) in pose.items():
        a = aa.to(dtype).reshape(-1, 1, 3).clone().requires_grad_(True)
        t = tr.to(dtype).reshape(-1, 1, 3).clone().requires_grad_(True)
        leaves[("axisangle", f)] = a
        leaves[("translation", f)] = t
        Tm = T.transformation_from_parameters(a, t, invert=(f < 0))
        Tm.retain_grad()
        outs[("cam_T_cam", 0, f)] = Tm
        # what PoseDecoder / PoseCNN emit (pose_decoder.py:49-54): (B, n, 1, 3); the posecnn branch of
        # generate_images_pred (trainer.py:366-375) rebuilds T from these
        outs[("axisangle", 0, f)] = a.reshape(-1, 1, 1, 3)
        outs[("translation", 0, f)] = t.reshape(-1, 1, 1, 3)
    if opt.predictive_mask:
        # what the mask decoder emits (trainer.py:96-98, 251-252): one sigmoid map per source and scale
        gm = torch.Generator().manual_seed(1000 + case["seed"])
        outs["predictive_mask"] = {}
        for s in range(4):
            if s not in scales:
                continue
            m = torch.sigmoid(2.0 * torch.randn(case["B"], n_src, case["H"] >> s, case["W"] >> s, generator=gm))
            m = m.to(dtype).requires_grad_(True)
            leaves[("mask", s)] = m
            outs["predictive_mask"][("disp", s)] = m
    me = build_self(T, opt, dtype)
    me.generate_images_pred(inputs, outs)
    for s in scales:
        outs[("depth", 0, s)].retain_grad()
    draws = [noise[s].to(dtype) for s in scales]      # one draw per entry of opt.scales, in list order
    real_randn = torch.randn
    it = iter(draws)

    def fake_randn(shape, *a, **k):
        z = next(it)
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z
    torch.randn = fake_randn
    # trainer.py:458 hard-codes `.cuda()` in the predictive-mask branch; on this CPU run it is made a
    # no-op from outside (the reference source stays unmodified)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        losses = me.compute_losses(inputs, outs)
    finally:
        torch.randn = real_randn
        torch.Tensor.cuda = real_cuda
    losses["loss"].backward()
    return dict(opt=opt, frame_ids=frame_ids, inputs=inputs, outs=outs, leaves=leaves,
                losses=losses, noise=noise, n_id=n_id, pose=pose)


def pack(res):
    d = {}
    fids = res["frame_ids"]
    d["frame_ids"] = np.array([str(f) for f in fids])
    opt = res["opt"]
    d["flags"] = np.array([int(opt.avg_reprojection), int(opt.disable_automasking), int(opt.no_ssim),
                           int(opt.v1_multiscale), int(opt.pose_model_type == "posecnn"),
                           int(opt.predictive_mask)])
    for k, v in res["inputs"].items():
        name = "in__" + "__".join(str(x) for x in (k if isinstance(k, tuple) else (k,)))
        d[name] = v.detach().numpy()
    scales = list(opt.scales)
    if scales != [0, 1, 2, 3]:
        d["scales"] = np.array(scales)
    for s in scales:
        d["disp__%d" % s] = res["leaves"][("disp", s)].detach().numpy()
        d["grad_disp__%d" % s] = res["leaves"][("disp", s)].grad.numpy()
        d["grad_depth__%d" % s] = res["outs"][("depth", 0, s)].grad.numpy()
        d["noise__%d" % s] = res["noise"][s].numpy()
        d["loss__%d" % s] = res["losses"]["loss/%d" % s].detach().numpy()
        key = "identity_selection/%d" % s
        if key in res["outs"]:
            d["idsel__%d" % s] = res["outs"][key].detach().numpy().astype(np.uint8)
    for s in range(4):
        if ("mask", s) in res["leaves"]:
            d["mask__%d" % s] = res["leaves"][("mask", s)].detach().numpy()
            d["grad_mask__%d" % s] = res["leaves"][("mask", s)].grad.numpy()
    d["loss"] = res["losses"]["loss"].detach().numpy()
    s0 = scales[0]
    d["depth__%d" % s0] = res["outs"][("depth", 0, s0)].detach().numpy()
    for f in fids[1:]:
        d["color__%s__%d" % (f, s0)] = res["outs"][("color", f, s0)].detach().numpy()
        if f != "s":
            d["axisangle__%s" % f] = res["leaves"][("axisangle", f)].detach().numpy()
            d["translation__%s" % f] = res["leaves"][("translation", f)].detach().numpy()
            d["grad_axisangle__%s" % f] = res["leaves"][("axisangle", f)].grad.numpy()
            d["grad_translation__%s" % f] = res["leaves"][("translation", f)].grad.numpy()
            d["cam_T_cam__%s" % f] = res["outs"][("cam_T_cam", 0, f)].detach().numpy()
            gT = res["outs"][("cam_T_cam", 0, f)].grad          # None under posecnn (T is rebuilt per scale)
            d["grad_cam_T_cam__%s" % f] = gT.numpy() if gT is not None else np.zeros((1,), np.float32)
    return d


def main():
    torch.manual_seed(0)
    T, MonodepthOptions = import_reference()
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        res = run_reference(T, MonodepthOptions, case)
        d = pack(res)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print("%-22s loss=%.8f  %s  %.0f KB" % (name, float(d["loss"]),
                                               list(d["frame_ids"]), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
