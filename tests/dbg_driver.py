"""Loads the DEBUG build of the CUDA library (libmd2loss_dbg.so: -DMD2_DBG_DEVICE -DMD2_BOUNDS_CHECK,
include/md2_debug.h) and drives it through the same Python host code as the product library.  Test infrastructure."""
import ctypes as C
import os

import numpy as np
import torch

from monodepth2_b200 import _capi
from emu_driver import DebugSink

DBG_LIB = os.path.join(_capi.LIB_DIR, "libmd2loss_dbg.so")
_lib = None


def debug_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(DBG_LIB):
            raise RuntimeError("libmd2loss_dbg.so not built - run `python -m monodepth2_b200.build --debug` "
                               "(__graft_entry__.build() does)")
        _lib = _capi.load_library(DBG_LIB)
        _lib.md2_debug_set_sink.argtypes = [C.c_void_p]
        _lib.md2_debug_set_sink.restype = C.c_int
        _lib.md2_debug_oob_count.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
        _lib.md2_debug_oob_count.restype = C.c_int
    return _lib


def oob_count(reset=True):
    n = C.c_ulonglong(0)
    _capi.check(debug_lib(), debug_lib().md2_debug_oob_count(C.byref(n), int(reset)), "md2_debug_oob_count")
    return n.value


class DecisionSink:
    """Device arrays the debug kernels export their discrete decisions into (layout of md2::DebugSink)."""

    def __init__(self, B, H, W, n_src, dev="cuda:0"):
        z = lambda shape, dt, fill=0: torch.full(shape, fill, dtype=dt, device=dev)
        self.x0 = z((4, B, n_src, H, W), torch.int16)
        self.y0 = z((4, B, n_src, H, W), torch.int16)
        self.mxy = z((4, B, n_src, H, W), torch.uint8)
        self.tag = z((4, B, H, W), torch.int8, -1)
        self.live = z((4, B, n_src, 3, H, W), torch.uint8)
        self.l1sgn = z((4, B, n_src, 3, H, W), torch.int8)
        self.sizes = [B * (H >> s) * (W >> s) for s in range(4)]
        self.offs = np.concatenate([[0], np.cumsum(self.sizes)[:-1]]).astype(np.int64)
        self.smx = z((int(sum(self.sizes)),), torch.int8)
        self.smy = z((int(sum(self.sizes)),), torch.int8)
        self.c = DebugSink(B=B, H=H, W=W, S=4, nsrc=n_src, x0=self.x0.data_ptr(), y0=self.y0.data_ptr(),
                           mxy=self.mxy.data_ptr(), tag=self.tag.data_ptr(), live=self.live.data_ptr(),
                           l1sgn=self.l1sgn.data_ptr(), smx=self.smx.data_ptr(), smy=self.smy.data_ptr())
        for s in range(4):
            self.c.smoff[s] = int(self.offs[s])

    def __enter__(self):
        _capi.check(debug_lib(), debug_lib().md2_debug_set_sink(C.byref(self.c)), "md2_debug_set_sink")
        return self

    def __exit__(self, *a):
        torch.cuda.synchronize()
        debug_lib().md2_debug_set_sink(None)

    def numpy(self):
        return dict(x0=self.x0.cpu().numpy(), y0=self.y0.cpu().numpy(), mxy=self.mxy.cpu().numpy(),
                    tag=self.tag.cpu().numpy(), live=self.live.cpu().numpy(), l1sgn=self.l1sgn.cpu().numpy(),
                    smx=self.smx.cpu().numpy(), smy=self.smy.cpu().numpy(), smoff=self.offs, smsizes=self.sizes)
