"""Decision-locked fp64 restatement of the path.  TEST INFRASTRUCTURE ONLY (protocol P4).

The loss is piece-wise smooth: its gradient depends on discrete decisions (the bilinear cell and
border-clip masks of grid_sample, the per-pixel argmin over [identity, reprojection] candidates, the
live/clamped state of the SSIM clamp, the signs inside the L1 and smoothness absolute values).  Two
correct fp32 implementations take a few of those decisions differently (rounding), which changes
aggregated gradients by ~1e-2 although every per-pixel term is right.  This module evaluates the same
algorithm (trainer.py:341-496, layers.py:16-25,139-248) in fp64 with the decisions *given* - exported
by the kernels' host emulator - so that what remains is pure arithmetic: the result must agree with the
kernel's aggregated gradients to 1e-4.  With ``decisions=None`` it takes its own decisions and must
agree with the plain fp64 oracle (oracle/view_synthesis.py) to ~1e-12.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import view_synthesis as O


def _gather_taps(img: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """img (B,3,H,W), integer x,y (B,H,W) -> (B,3,H,W) texels."""
    B, C, H, W = img.shape
    idx = (y * W + x).reshape(B, 1, H * W).expand(B, C, H * W)
    return img.reshape(B, C, H * W).gather(2, idx).reshape(B, C, H, W)


def locked_loss(inputs: Dict, outputs: Dict, cfg: O.OracleConfig, noise: Optional[List[torch.Tensor]],
                decisions: Optional[Dict] = None, export: Optional[Dict] = None) -> Dict[str, torch.Tensor]:
    """fp64 loss with discrete decisions taken from ``decisions`` (or from the forward itself when None;
    ``export`` then receives them).  Keys of ``decisions``: x0,y0,mx,my [(s,f)] (B,H,W); tag [s] (B,H,W);
    live [(s,f)] (B,3,H,W) bool; l1sgn [(s,f)] (B,3,H,W); smx,smy [s] signs of the smoothness differences."""
    H, W = cfg.height, cfg.width
    srcs = list(cfg.frame_ids[1:])
    target = inputs[("color", 0, 0)]
    dt = target.dtype
    S = len(cfg.scales)
    sx = W / (W - 1.0) if not cfg.align_corners else 1.0
    sy = H / (H - 1.0) if not cfg.align_corners else 1.0
    off = -0.5 if not cfg.align_corners else 0.0
    losses = {}
    total = 0
    for i, s in enumerate(cfg.scales):
        disp_s = outputs[("disp", s)]
        D = F.interpolate(disp_s, [H, W], mode="bilinear", align_corners=False)
        _, depth = O.disp_to_depth(D, cfg.min_depth, cfg.max_depth)
        pts = O.backproject(depth, inputs[("inv_K", 0)])
        rls = []
        for fi, f in enumerate(srcs):
            T = inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]
            P = torch.matmul(inputs[("K", 0)], T)[:, :3, :]
            cam = torch.matmul(P, pts)
            u = (cam[:, 0] / (cam[:, 2] + cfg.eps)).reshape(-1, H, W)
            v = (cam[:, 1] / (cam[:, 2] + cfg.eps)).reshape(-1, H, W)
            ix, iy = u * sx + off, v * sy + off
            if decisions is None:
                mx = (ix > 0) & (ix < W - 1)
                my = (iy > 0) & (iy < H - 1)
                x0 = ix.detach().clamp(0, W - 1).floor().long()
                y0 = iy.detach().clamp(0, H - 1).floor().long()
                if export is not None:
                    export.setdefault("x0", {})[(s, f)] = x0
                    export.setdefault("y0", {})[(s, f)] = y0
                    export.setdefault("mx", {})[(s, f)] = mx
                    export.setdefault("my", {})[(s, f)] = my
            else:
                x0, y0 = decisions["x0"][(s, f)].long(), decisions["y0"][(s, f)].long()
                mx, my = decisions["mx"][(s, f)].bool(), decisions["my"][(s, f)].bool()
            ixc = torch.where(mx, ix, ix.detach().clamp(0, W - 1))
            iyc = torch.where(my, iy, iy.detach().clamp(0, H - 1))
            wx = (ixc - x0.to(dt)).unsqueeze(1)
            wy = (iyc - y0.to(dt)).unsqueeze(1)
            x1, y1 = (x0 + 1).clamp(max=W - 1), (y0 + 1).clamp(max=H - 1)
            img = inputs[("color", f, 0)]
            nw, ne = _gather_taps(img, x0, y0), _gather_taps(img, x1, y0)
            sw, se = _gather_taps(img, x0, y1), _gather_taps(img, x1, y1)
            top = nw + wx * (ne - nw)
            bot = sw + wx * (se - sw)
            pred = top + wy * (bot - top)
            # L1 with locked signs
            diff = pred - target
            if decisions is None:
                sg = torch.sign(diff.detach())
                if export is not None:
                    export.setdefault("l1sgn", {})[(s, f)] = sg
            else:
                sg = decisions["l1sgn"][(s, f)].to(dt)
            l1 = (sg * diff).mean(1, keepdim=True)
            if cfg.no_ssim:
                rl = l1
            else:
                xp = F.pad(pred, (1, 1, 1, 1), mode="reflect")
                yp = F.pad(target, (1, 1, 1, 1), mode="reflect")
                mu_x, mu_y = F.avg_pool2d(xp, 3, 1), F.avg_pool2d(yp, 3, 1)
                sig_x = F.avg_pool2d(xp * xp, 3, 1) - mu_x * mu_x
                sig_y = F.avg_pool2d(yp * yp, 3, 1) - mu_y * mu_y
                sig_xy = F.avg_pool2d(xp * yp, 3, 1) - mu_x * mu_y
                n = (2 * mu_x * mu_y + cfg.ssim_c1) * (2 * sig_xy + cfg.ssim_c2)
                d = (mu_x * mu_x + mu_y * mu_y + cfg.ssim_c1) * (sig_x + sig_y + cfg.ssim_c2)
                raw = (1 - n / d) / 2
                if decisions is None:
                    live = (raw.detach() >= 0) & (raw.detach() <= 1)
                    if export is not None:
                        export.setdefault("live", {})[(s, f)] = live
                else:
                    live = decisions["live"][(s, f)].bool()
                ssim = torch.where(live, raw, raw.detach().clamp(0, 1))
                rl = 0.85 * ssim.mean(1, keepdim=True) + 0.15 * l1
            rls.append(rl)
        reproj = torch.cat(rls, 1)
        if cfg.avg_reprojection:
            reproj = reproj.mean(1, keepdim=True)
        # winner: locked tag (-1 = an identity candidate, whose value does not depend on the leaves)
        if decisions is None:
            if not cfg.disable_automasking:
                ident = torch.cat([O.reprojection_loss(inputs[("color", f, 0)], target, cfg) for f in srcs], 1)
                if cfg.avg_reprojection:
                    ident = ident.mean(1, keepdim=True)
                ident = ident + noise[i].to(dt) * 0.00001
                combined = torch.cat([ident, reproj.detach()], 1)
                idxs = combined.argmin(1) if combined.shape[1] > 1 else torch.zeros_like(combined[:, 0]).long()
                tag = idxs - ident.shape[1]
                tag = torch.where(tag >= 0, tag, torch.full_like(tag, -1))
            else:
                tag = reproj.detach().argmin(1) if reproj.shape[1] > 1 else torch.zeros_like(reproj[:, 0]).long()
            if export is not None:
                export.setdefault("tag", {})[s] = tag
        else:
            tag = decisions["tag"][s].long()
        sel = torch.zeros_like(reproj[:, 0])
        for k in range(reproj.shape[1]):
            sel = sel + torch.where(tag == k, reproj[:, k], torch.zeros_like(sel))
        photo = sel.mean()          # identity-won pixels contribute a constant: irrelevant for gradients
        # smoothness with locked signs
        color = inputs[("color", 0, s)]
        mean_disp = disp_s.mean(2, True).mean(3, True)
        nd = disp_s / (mean_disp + 1e-7)
        dx = nd[..., :, :-1] - nd[..., :, 1:]
        dy = nd[..., :-1, :] - nd[..., 1:, :]
        if decisions is None:
            sgx, sgy = torch.sign(dx.detach()), torch.sign(dy.detach())
            if export is not None:
                export.setdefault("smx", {})[s] = sgx
                export.setdefault("smy", {})[s] = sgy
        else:
            sgx, sgy = decisions["smx"][s].to(dt), decisions["smy"][s].to(dt)
        wxi = torch.exp(-(color[..., :, :-1] - color[..., :, 1:]).abs().mean(1, keepdim=True))
        wyi = torch.exp(-(color[..., :-1, :] - color[..., 1:, :]).abs().mean(1, keepdim=True))
        smooth = (sgx * dx * wxi).mean() + (sgy * dy * wyi).mean()
        loss = photo + cfg.disparity_smoothness * smooth / (2 ** s)
        losses["loss/{}".format(s)] = loss
        total = total + loss
    losses["loss"] = total / S
    return losses
