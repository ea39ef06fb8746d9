"""Seeded synthetic batches shaped like the reference's input contract.

Key names and tensor conventions follow ``MonoDataset.__getitem__``
(/root/reference/datasets/mono_dataset.py:114-200), the KITTI normalised
intrinsics (/root/reference/datasets/kitti_dataset.py:29-32) and the
``PoseDecoder`` output scale (/root/reference/networks/pose_decoder.py:49).
All tensors are generated on the CPU from a ``torch.Generator`` so that the CPU
oracle, the PyTorch-CUDA path and the fused kernels see identical bits
(IID seed 0, STRUCTURED seed 5).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

KITTI_K = np.array([[0.58, 0, 0.5, 0],
                    [0, 1.92, 0.5, 0],
                    [0, 0, 1, 0],
                    [0, 0, 0, 1]], dtype=np.float32)


def _rodrigues(aa: torch.Tensor) -> torch.Tensor:
    """(B,3) axis-angle -> (B,3,3); same parameterisation as layers.py:64-103."""
    ang = aa.norm(dim=1, keepdim=True)
    ax = aa / (ang + 1e-7)
    ca, sa = torch.cos(ang)[:, 0], torch.sin(ang)[:, 0]
    C = 1 - ca
    x, y, z = ax[:, 0], ax[:, 1], ax[:, 2]
    R = torch.stack([
        torch.stack([x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa], 1),
        torch.stack([x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa], 1),
        torch.stack([z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca], 1)], 1)
    return R


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool) -> torch.Tensor:
    """(B,3),(B,3) -> (B,4,4): Trans(t)·R, or Rᵀ·Trans(−t) when ``invert`` (layers.py:28-45)."""
    B = axisangle.shape[0]
    R = _rodrigues(axisangle)
    T = torch.zeros(B, 4, 4, dtype=axisangle.dtype)
    T[:, 3, 3] = 1
    if invert:
        Rt = R.transpose(1, 2)
        T[:, :3, :3] = Rt
        T[:, :3, 3] = -(Rt @ translation.unsqueeze(-1))[:, :, 0]
    else:
        T[:, :3, :3] = R
        T[:, :3, 3] = translation
    return T


def intrinsics(batch: int, height: int, width: int, gen: torch.Generator = None,
               jitter: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Scale-0 K and inv_K = pinv(K), (B,4,4) fp32 (mono_dataset.py:164-173)."""
    Ks, iKs = [], []
    for _ in range(batch):
        K = KITTI_K.copy()
        K[0, :] *= width
        K[1, :] *= height
        if jitter:
            r = torch.rand(4, generator=gen).numpy()
            K[0, 0] *= 0.9 + 0.2 * r[0]
            K[1, 1] *= 0.9 + 0.2 * r[1]
            K[0, 2] += (r[2] * 0.04 - 0.02) * width
            K[1, 2] += (r[3] * 0.04 - 0.02) * height
        Ks.append(torch.from_numpy(K))
        iKs.append(torch.from_numpy(np.linalg.pinv(K).astype(np.float32)))
    return torch.stack(Ks), torch.stack(iKs)


def _box_blur(x: torch.Tensor, k: int) -> torch.Tensor:
    return F.avg_pool2d(F.pad(x, (k // 2,) * 4, mode="reflect"), k, 1)


def make_batch(batch: int = 12, height: int = 192, width: int = 640,
               frame_ids: Sequence = (0, -1, 1), num_scales: int = 4, seed: int = 0,
               kind: str = "iid", jitter_K: bool = False,
               noise_seed: int = 99, n_id: int = None, all_scale_K: bool = False, multiscale_noise: bool = False):
    """Returns ``(inputs, outputs, pose, noise)`` on the CPU in fp32.

    inputs : {("color", f, s)}, ("K", 0), ("inv_K", 0), "stereo_T"
    outputs: {("disp", s)}, {("cam_T_cam", 0, f)} for temporal sources
    pose   : {f: (axisangle (B,3), translation (B,3))} the leaves behind cam_T_cam
    noise  : list over scales of (B, n_id, H, W) standard-normal tie-break draws
    """
    g = torch.Generator().manual_seed(seed)
    inputs: Dict = {}
    outputs: Dict = {}
    B, H, W = batch, height, width
    if kind == "iid":
        for f in frame_ids:
            for s in range(num_scales):
                inputs[("color", f, s)] = torch.rand(B, 3, H >> s, W >> s, generator=g)
        for s in range(num_scales):
            outputs[("disp", s)] = torch.rand(B, 1, H >> s, W >> s, generator=g)
    elif kind == "structured":
        base = _box_blur(torch.rand(B, 3, H + 64, W + 64, generator=g), 9)
        lo = base.amin(dim=(1, 2, 3), keepdim=True)
        hi = base.amax(dim=(1, 2, 3), keepdim=True)
        base = (base - lo) / (hi - lo)
        shift = {0: 0, -1: 3, 1: -3, "s": 6}
        for f in frame_ids:                       # further temporal neighbours (--frame_ids 0 -2 -1 1 2): 3 px per frame
            if f not in shift:
                shift[f] = -3 * f
        for f in frame_ids:
            x0 = 32 + shift[f]
            frame = base[:, :, 32:32 + H, x0:x0 + W].contiguous()
            for s in range(num_scales):
                inputs[("color", f, s)] = frame if s == 0 else F.avg_pool2d(frame, 2 ** s)
        for s in range(num_scales):
            d = _box_blur(torch.rand(B, 1, H >> s, W >> s, generator=g), 5)
            lo = d.amin(dim=(1, 2, 3), keepdim=True)
            hi = d.amax(dim=(1, 2, 3), keepdim=True)
            outputs[("disp", s)] = (0.2 + 0.6 * (d - lo) / (hi - lo)).contiguous()
    else:
        raise ValueError(kind)

    K, inv_K = intrinsics(B, H, W, g, jitter_K)
    inputs[("K", 0)] = K
    inputs[("inv_K", 0)] = inv_K
    if all_scale_K:      # mono_dataset.py:164-173: one K / inv_K per pyramid level (for --v1_multiscale)
        for s in range(1, num_scales):
            Ks = K.clone()
            Ks[:, 0, :] /= 2 ** s
            Ks[:, 1, :] /= 2 ** s
            inputs[("K", s)] = Ks
            inputs[("inv_K", s)] = torch.from_numpy(
                np.stack([np.linalg.pinv(k.numpy()).astype(np.float32) for k in Ks]))
    stereo_T = torch.eye(4).repeat(B, 1, 1)
    stereo_T[:, 0, 3] = 0.1
    inputs["stereo_T"] = stereo_T

    pose: Dict = {}
    for f in frame_ids[1:]:
        if f == "s":
            continue
        aa = 0.01 * torch.randn(B, 3, generator=g)
        tr = 0.01 * torch.randn(B, 3, generator=g)
        pose[f] = (aa, tr)
        outputs[("cam_T_cam", 0, f)] = pose_matrix(aa, tr, invert=(f < 0))

    n_src = len(frame_ids) - 1
    if n_id is None:
        n_id = n_src
    gn = torch.Generator().manual_seed(noise_seed)
    noise: List[torch.Tensor] = [torch.randn(B, n_id, H >> (s if multiscale_noise else 0),
                                             W >> (s if multiscale_noise else 0), generator=gn)
                                 for s in range(num_scales)]
    return inputs, outputs, pose, noise
