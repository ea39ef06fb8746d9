"""Colour pyramid on the GPU: the device-side replacement for the resize half of
``MonoDataset.preprocess`` (/root/reference/datasets/mono_dataset.py:90-103).

The reference resizes every frame on the host with ``transforms.Resize((h // 2**i, w // 2**i),
interpolation=Image.ANTIALIAS)`` on PIL images, scale 0 from the native frame and scale i from scale i-1;
here the same chain runs as two integer kernels per level on a whole batch of uint8 frames, byte-exact
with Pillow (``md2_resize_lanczos_u8``).  The uint8 levels go straight into the uint8 entry of the fused loss.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _capi


class ColorPyramid:
    """``ColorPyramid(height, width, num_scales)(frames_u8) -> [level 0, ..., level num_scales-1]``.

    ``frames_u8``: uint8 CUDA tensor, (B,H0,W0,3) (numpy view of the PIL images) or (B,3,H0,W0); the levels
    come back in the same layout at (height >> i, width >> i)."""

    def __init__(self, height: int, width: int, num_scales: int = 4):
        self.height, self.width, self.num_scales = int(height), int(width), int(num_scales)
        self.lib = _capi.load_library()
        self._plans: Dict[Tuple[int, int, int, int], C.c_void_p] = {}

    def _plan(self, ih, iw, oh, ow):
        key = (ih, iw, oh, ow)
        p = self._plans.get(key)
        if p is None:
            p = C.c_void_p()
            _capi.check(self.lib, self.lib.md2_resize_plan_create(ih, iw, oh, ow, C.byref(p)), "md2_resize_plan_create")
            self._plans[key] = p
        return p

    def __del__(self):
        for p in getattr(self, "_plans", {}).values():
            try:
                self.lib.md2_resize_plan_destroy(p)
            except Exception:
                pass

    def resize(self, x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4:
            raise RuntimeError("frames must be a 4-d uint8 CUDA tensor (there is no CPU path)")
        hwc = x.shape[-1] == 3 and x.shape[1] != 3
        x = x.contiguous()
        B = x.shape[0]
        ih, iw = (x.shape[1], x.shape[2]) if hwc else (x.shape[2], x.shape[3])
        plan = self._plan(ih, iw, out_h, out_w)
        out = torch.empty((B, out_h, out_w, 3) if hwc else (B, 3, out_h, out_w), dtype=torch.uint8, device=x.device)
        n = C.c_size_t(0)
        _capi.check(self.lib, self.lib.md2_resize_scratch_bytes(plan, B, C.byref(n)), "md2_resize_scratch_bytes")
        scratch = torch.empty(max(n.value, 1), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            st = self.lib.md2_resize_lanczos_u8(plan, x.data_ptr(), out.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                                B, int(hwc), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _capi.check(self.lib, st, "md2_resize_lanczos_u8")
        return out

    def __call__(self, frames_u8: torch.Tensor) -> List[torch.Tensor]:
        levels, cur = [], frames_u8
        for i in range(self.num_scales):          # mono_dataset.py:98-103: scale i from scale i-1
            cur = self.resize(cur, self.height >> i, self.width >> i)
            levels.append(cur)
        return levels


class ColorAug:
    """The colour augmentation of ``MonoDataset`` on the GPU (/root/reference/datasets/mono_dataset.py:60-70,107-109,
    169-176): ``color_aug = transforms.ColorJitter.get_params(brightness, contrast, saturation, hue)`` applied to the
    PIL frames before ``to_tensor`` - here on a batch of uint8 CUDA frames, byte-exact with torchvision's PIL path
    (``md2_color_jitter_u8``).

        aug = ColorAug()
        fn_idx, b, c, s, h = transforms.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
        out_u8 = aug(frames_u8, [(fn_idx, b, c, s, h)] * n)     # one parameter set per image (the reference draws one
                                                                # per dataset item and uses it for all its frames)
    ``frames_u8``: (n,H,W,3) or (n,3,H,W) uint8 CUDA tensor; a factor given as ``None`` is skipped, as in torchvision."""

    def __init__(self):
        self.lib = _capi.load_library()

    @staticmethod
    def pack_params(params) -> torch.Tensor:
        """[(fn_idx, brightness, contrast, saturation, hue), ...] -> (n, 8) int32 CPU tensor laid out as md2_color_jitter
        (order[4], brightness, contrast, saturation as float bits, hue_shift)."""
        import numpy as np
        rows = np.zeros((len(params), 8), dtype=np.int32)
        fl = rows.view(np.float32)
        for n, (fn_idx, b, c, s, h) in enumerate(params):
            fac = {0: b, 1: c, 2: s, 3: h}
            order = [int(i) for i in (fn_idx.tolist() if hasattr(fn_idx, "tolist") else fn_idx)]
            order = [(i if fac[i] is not None else -1) for i in order] + [-1] * (4 - len(order))
            rows[n, :4] = order[:4]
            fl[n, 4] = 1.0 if b is None else float(b)
            fl[n, 5] = 1.0 if c is None else float(c)
            fl[n, 6] = 1.0 if s is None else float(s)
            # torchvision _functional_pil.adjust_hue: np.int32(hue_factor * 255).astype(np.uint8)
            rows[n, 7] = 0 if h is None else int(np.int32(float(h) * 255).astype(np.uint8))
        return torch.from_numpy(rows)

    def __call__(self, frames_u8: torch.Tensor, params) -> torch.Tensor:
        x = frames_u8
        if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4:
            raise RuntimeError("frames must be a 4-d uint8 CUDA tensor (there is no CPU path)")
        hwc = x.shape[-1] == 3 and x.shape[1] != 3
        x = x.contiguous()
        n = x.shape[0]
        h, w = (x.shape[1], x.shape[2]) if hwc else (x.shape[2], x.shape[3])
        p = params if torch.is_tensor(params) else self.pack_params(params)
        if tuple(p.shape) != (n, 8) or p.dtype != torch.int32:
            raise RuntimeError("one parameter set per image: expected an (n, 8) int32 tensor or a list of n tuples")
        p = p.to(x.device, non_blocking=True).contiguous()
        out = torch.empty_like(x)
        scratch = torch.empty(n * 8, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            st = self.lib.md2_color_jitter_u8(x.data_ptr(), out.data_ptr(), p.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                              n, h, w, int(hwc), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _capi.check(self.lib, st, "md2_color_jitter_u8")
        return out
