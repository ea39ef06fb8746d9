"""CPU oracle for the monodepth2 view-synthesis loss path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``monodepth2_b200/`` imports it.

It restates, in plain PyTorch (runs on CPU in fp32 or fp64), the algorithm of

  * ``Trainer.generate_images_pred``      /root/reference/trainer.py:341-391
  * ``Trainer.compute_reprojection_loss`` /root/reference/trainer.py:393-405
  * ``Trainer.compute_losses``            /root/reference/trainer.py:407-496
  * ``layers.py``: disp_to_depth :16-25, transformation_from_parameters :28-45,
    get_translation_matrix :48-61, rot_from_axisangle :64-103,
    BackprojectDepth :139-168, Project3D :171-193, get_smooth_loss :202-215,
    SSIM :218-248

The arithmetic of the reference lives in its third-party dependency PyTorch
(README.md:38 pins "pytorch=0.4.1" by prose only; installed here: torch 2.11):
``F.grid_sample`` (bilinear, padding_mode="border", align_corners left at the
installed default = False), ``F.interpolate`` (bilinear, align_corners=False),
``AvgPool2d(3, 1)``, ``ReflectionPad2d(1)``, ``torch.min(dim=1)``.  The oracle calls
the same library entry points at the same call sites, so it is the reference's
algorithm executed by the reference's own numerical library.

Pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the pin is the reference itself: ``tests/golden/make_golden.py``
imports /root/reference in the build container, runs the *unmodified*
``Trainer.generate_images_pred`` + ``compute_losses`` + ``backward`` on seeded
inputs and commits the results as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this oracle against every one of them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """The subset of ``options.py`` flags that shape the path (options.py:52-119)."""
    height: int = 192
    width: int = 640
    scales: Sequence[int] = (0, 1, 2, 3)
    frame_ids: Sequence = (0, -1, 1)          # "s" appended by the caller for --use_stereo
    min_depth: float = 0.1
    max_depth: float = 100.0
    disparity_smoothness: float = 1e-3
    avg_reprojection: bool = False
    disable_automasking: bool = False
    no_ssim: bool = False
    v1_multiscale: bool = False               # trainer.py:347-352,417-420: source_scale = scale
    predictive_mask: bool = False             # trainer.py:447-459 (requires disable_automasking, trainer.py:90-92)
    posecnn: bool = False                     # trainer.py:366-375: translation scaled by the mean inverse depth
    # trainer.py:384-387 leaves align_corners unspecified -> False on torch >= 1.3.
    align_corners: bool = False
    ssim_c1: float = 0.01 ** 2
    ssim_c2: float = 0.03 ** 2
    eps: float = 1e-7


# --------------------------------------------------------------------------- pose
def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """Rodrigues formula, layers.py:64-103.  vec: (B,1,3) -> (B,4,4)."""
    angle = vec.norm(p=2, dim=2, keepdim=True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    rows = [
        [x * (x * C) + ca, x * (y * C) - z * sa, z * (x * C) + y * sa],
        [x * (y * C) + z * sa, y * (y * C) + ca, y * (z * C) - x * sa],
        [z * (x * C) - y * sa, y * (z * C) + x * sa, z * (z * C) + ca],
    ]
    B = vec.shape[0]
    rot3 = torch.stack([torch.stack([e.reshape(B) for e in r], dim=1) for r in rows], dim=1)
    rot = torch.cat([torch.cat([rot3, vec.new_zeros(B, 3, 1)], 2),
                     torch.cat([vec.new_zeros(B, 1, 3), vec.new_ones(B, 1, 1)], 2)], 1)
    return rot


def translation_matrix(t: torch.Tensor) -> torch.Tensor:
    """layers.py:48-61.  t: (B,1,3) or (B,3) -> (B,4,4)."""
    B = t.shape[0]
    eye = torch.eye(4, dtype=t.dtype, device=t.device).expand(B, 4, 4)
    col = torch.cat([t.reshape(B, 3, 1), t.new_zeros(B, 1, 1)], 1)
    pad = torch.cat([t.new_zeros(B, 4, 3), col], 2)
    return eye + pad


def transformation_from_parameters(axisangle, translation, invert=False):
    """layers.py:28-45: T = Trans(t) @ R, or R^T @ Trans(-t) when ``invert``."""
    R = rot_from_axisangle(axisangle)
    t = translation
    if invert:
        return torch.matmul(R.transpose(1, 2), translation_matrix(-t))
    return torch.matmul(translation_matrix(t), R)


# ----------------------------------------------------------------------- geometry
def disp_to_depth(disp, min_depth, max_depth):
    """layers.py:16-25."""
    lo, hi = 1.0 / max_depth, 1.0 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1.0 / scaled


def backproject(depth, inv_K):
    """layers.py:139-168: (B,1,H,W),(B,4,4) -> homogeneous camera points (B,4,H*W)."""
    B, _, H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype), torch.arange(W, dtype=depth.dtype),
                            indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=depth.dtype)], 0)
    pix = pix.to(depth.device).unsqueeze(0).expand(B, 3, H * W)
    cam = torch.matmul(inv_K[:, :3, :3], pix) * depth.reshape(B, 1, -1)
    return torch.cat([cam, cam.new_ones(B, 1, H * W)], 1)


def project(points, K, T, height, width, eps=1e-7):
    """layers.py:171-193: -> sampling grid (B,H,W,2) normalised as the reference does."""
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2:3, :] + eps)
    pix = pix.reshape(B, 2, height, width).permute(0, 2, 3, 1)
    gx = (pix[..., 0] / (width - 1) - 0.5) * 2
    gy = (pix[..., 1] / (height - 1) - 0.5) * 2
    return torch.stack([gx, gy], dim=-1)


# --------------------------------------------------------------------------- loss
def ssim_dissimilarity(x, y, c1=0.01 ** 2, c2=0.03 ** 2):
    """layers.py:218-248: clamp((1 - SSIM)/2, 0, 1) on 3x3 reflect-padded windows."""
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    yp = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(xp, 3, 1)
    mu_y = F.avg_pool2d(yp, 3, 1)
    sig_x = F.avg_pool2d(xp * xp, 3, 1) - mu_x * mu_x
    sig_y = F.avg_pool2d(yp * yp, 3, 1) - mu_y * mu_y
    sig_xy = F.avg_pool2d(xp * yp, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + c1) * (2 * sig_xy + c2)
    d = (mu_x * mu_x + mu_y * mu_y + c1) * (sig_x + sig_y + c2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_loss(pred, target, cfg: OracleConfig):
    """trainer.py:393-405."""
    l1 = (target - pred).abs().mean(1, keepdim=True)
    if cfg.no_ssim:
        return l1
    s = ssim_dissimilarity(pred, target, cfg.ssim_c1, cfg.ssim_c2).mean(1, keepdim=True)
    return 0.85 * s + 0.15 * l1


def smooth_loss(disp, img):
    """layers.py:202-215 (edge-aware first-order smoothness)."""
    dx = (disp[..., :, :-1] - disp[..., :, 1:]).abs()
    dy = (disp[..., :-1, :] - disp[..., 1:, :]).abs()
    ix = (img[..., :, :-1] - img[..., :, 1:]).abs().mean(1, keepdim=True)
    iy = (img[..., :-1, :] - img[..., 1:, :]).abs().mean(1, keepdim=True)
    return (dx * torch.exp(-ix)).mean() + (dy * torch.exp(-iy)).mean()


def generate_images_pred(inputs: Dict, outputs: Dict, cfg: OracleConfig) -> None:
    """trainer.py:341-391 (both source_scale branches and the posecnn translation rescale)."""
    H, W = cfg.height, cfg.width
    for s in cfg.scales:
        disp = outputs[("disp", s)]
        if cfg.v1_multiscale:
            src_scale = s
        else:
            disp = F.interpolate(disp, [H, W], mode="bilinear", align_corners=False)
            src_scale = 0
        hs, ws = H >> src_scale, W >> src_scale
        _, depth = disp_to_depth(disp, cfg.min_depth, cfg.max_depth)
        outputs[("depth", 0, s)] = depth
        for f in cfg.frame_ids[1:]:
            T = inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]
            if cfg.posecnn and f != "s":
                mean_inv_depth = (1 / depth).mean(3, True).mean(2, True)
                T = transformation_from_parameters(outputs[("axisangle", 0, f)][:, 0],
                                                   outputs[("translation", 0, f)][:, 0] * mean_inv_depth[:, 0], f < 0)
            pts = backproject(depth, inputs[("inv_K", src_scale)])
            grid = project(pts, inputs[("K", src_scale)], T, hs, ws, cfg.eps)
            outputs[("sample", f, s)] = grid
            outputs[("color", f, s)] = F.grid_sample(
                inputs[("color", f, src_scale)], grid, mode="bilinear", padding_mode="border",
                align_corners=cfg.align_corners)


def compute_losses(inputs: Dict, outputs: Dict, cfg: OracleConfig,
                   noise: Optional[List[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """trainer.py:407-496.  ``noise[i]`` (one tensor per scale, the shape of the identity
    losses) replaces the ``torch.randn`` draw of trainer.py:468-469 when given."""
    losses: Dict[str, torch.Tensor] = {}
    total = 0
    srcs = list(cfg.frame_ids[1:])
    for i, s in enumerate(cfg.scales):
        src_scale = s if cfg.v1_multiscale else 0
        target = inputs[("color", 0, src_scale)]
        disp = outputs[("disp", s)]
        color = inputs[("color", 0, s)]
        reproj = torch.cat([reprojection_loss(outputs[("color", f, s)], target, cfg) for f in srcs], 1)
        extra = 0
        if cfg.disable_automasking and cfg.predictive_mask:
            # trainer.py:447-459: per-source mask from the mask decoder, up-sampled like the disparity;
            # weights the reprojection losses and is pushed towards 1 by 0.2 * BCE(mask, 1)
            mask = outputs["predictive_mask"][("disp", s)]
            if not cfg.v1_multiscale:
                mask = F.interpolate(mask, [cfg.height, cfg.width], mode="bilinear", align_corners=False)
            reproj = reproj * mask
            extra = 0.2 * F.binary_cross_entropy(mask, torch.ones_like(mask))
        if cfg.avg_reprojection:       # trainer.py:461-462 (after the mask weighting)
            reproj = reproj.mean(1, keepdim=True)
        if not cfg.disable_automasking:
            ident = torch.cat([reprojection_loss(inputs[("color", f, src_scale)], target, cfg) for f in srcs], 1)
            if cfg.avg_reprojection:
                ident = ident.mean(1, keepdim=True)
            z = noise[i] if noise is not None else torch.randn(ident.shape, dtype=ident.dtype,
                                                                device=ident.device)
            ident = ident + z.to(ident.dtype) * 0.00001
            combined = torch.cat([ident, reproj], 1)
        else:
            combined = reproj
        if combined.shape[1] == 1:
            to_optimise = combined
        else:
            to_optimise, idxs = torch.min(combined, dim=1)
        if not cfg.disable_automasking:
            outputs["identity_selection/{}".format(s)] = (idxs > ident.shape[1] - 1).to(disp.dtype)
        loss = to_optimise.mean() + extra
        mean_disp = disp.mean(2, True).mean(3, True)
        norm_disp = disp / (mean_disp + 1e-7)
        loss = loss + cfg.disparity_smoothness * smooth_loss(norm_disp, color) / (2 ** s)
        total = total + loss
        losses["loss/{}".format(s)] = loss
    losses["loss"] = total / len(cfg.scales)
    return losses


def view_synthesis_loss(inputs: Dict, outputs: Dict, cfg: OracleConfig,
                        noise: Optional[List[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """generate_images_pred followed by compute_losses (trainer.py:257-258)."""
    generate_images_pred(inputs, outputs, cfg)
    return compute_losses(inputs, outputs, cfg, noise)


# ------------------------------------------------------------------ monitoring path
def compute_depth_errors(gt, pred):
    """layers.py:251-269."""
    thresh = torch.max((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).float().mean()
    a2 = (thresh < 1.25 ** 2).float().mean()
    a3 = (thresh < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def depth_metrics(depth_pred, depth_gt, crop=(153, 371, 44, 1197)):
    """Trainer.compute_depth_losses (trainer.py:498-526): (7,) tensor abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3."""
    hg, wg = depth_gt.shape[-2:]
    pred = torch.clamp(F.interpolate(depth_pred, [hg, wg], mode="bilinear", align_corners=False), 1e-3, 80).detach()
    mask = depth_gt > 0
    crop_mask = torch.zeros_like(mask)
    crop_mask[:, :, crop[0]:crop[1], crop[2]:crop[3]] = 1
    mask = mask * crop_mask
    gt = depth_gt[mask]
    pred = pred[mask]
    pred = pred * (torch.median(gt) / torch.median(pred))
    pred = torch.clamp(pred, min=1e-3, max=80)
    return torch.stack([torch.as_tensor(v, dtype=torch.float32) for v in compute_depth_errors(gt, pred)])
