"""Builds and drives the host emulator of the marching kernels (tests/emu/md2_emu.cpp).

Test infrastructure only: it checks the kernel arithmetic and tiling on a box with no
GPU.  The product path (monodepth2_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from monodepth2_b200._capi import MAX_SCALES, Md2Problem, Md2Tensors

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "md2_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libmd2emu.so")
CORE = [os.path.join(HERE, "..", "monodepth2_b200", "csrc", f) for f in ("md2_core.cuh", "md2_plan.h")] + \
       [os.path.join(HERE, "..", "include", "md2_loss.h")]


def build_emu():
    deps = [SRC] + CORE
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           "-o", OUT, SRC])
    return OUT


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class DebugSink(C.Structure):
    """ctypes mirror of md2::DebugSink (monodepth2_b200/csrc/md2_core.cuh, host build only)."""
    _fields_ = [("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("S", C.c_int), ("nsrc", C.c_int),
                ("x0", C.c_void_p), ("y0", C.c_void_p), ("mxy", C.c_void_p), ("tag", C.c_void_p),
                ("live", C.c_void_p), ("l1sgn", C.c_void_p), ("smx", C.c_void_p), ("smy", C.c_void_p),
                ("smoff", C.c_long * 4)]


def quantise_u8(a):
    """float image in [0,1] -> the uint8 image a dataloader would hold before ToTensor"""
    return np.clip(np.rint(np.asarray(a, np.float64) * 255.0), 0, 255).astype(np.uint8)


def run_emu(g, want_grad=True, rows_per_segment=0, align_corners=False, side_outputs=True, decisions=False,
            pose_leaves=False, u8=None):
    """Run the emulator on a Golden fixture; returns dict of numpy outputs.

    u8: None (float images as stored), "f32" (float images = ToTensor of the quantised uint8 frames),
    "chw" / "hwc" (the quantised uint8 frames through the uint8 entry, planar / interleaved)."""
    lib = C.CDLL(build_emu())
    lib.md2_emu_workspace_bytes.argtypes = [C.POINTER(Md2Problem), C.POINTER(C.c_size_t)]
    lib.md2_emu_view_synthesis_loss.argtypes = [C.POINTER(Md2Problem), C.POINTER(Md2Tensors), C.c_void_p, C.c_size_t]
    z = g.z
    B, H, W = g.B, g.H, g.W
    srcs = g.frame_ids[1:]
    scales = list(g.scales)          # slot i of every per-scale array = pyramid level scales[i] (md2_problem.scale_level)
    lv = (C.c_int * 4)(*(scales + [0] * (4 - len(scales)))) if scales != list(range(len(scales))) else (C.c_int * 4)()
    p = Md2Problem(batch=B, height=H, width=W, num_scales=len(scales), num_src=len(srcs), scale_level=lv,
                   automask=int(not g.disable_automasking), avg_reprojection=int(g.avg_reprojection),
                   align_corners=int(align_corners), min_depth=0.1, max_depth=100.0,
                   disparity_smoothness=1e-3, want_grad=int(want_grad), rows_per_segment=rows_per_segment,
                   no_ssim=int(g.no_ssim), posecnn=int(g.posecnn), predictive_mask=int(g.predictive_mask))
    pose_leaves = pose_leaves or g.posecnn      # posecnn rebuilds T per scale from the leaves (trainer.py:366-375)
    keep = []

    def arr(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a

    def img(key):
        """(float pointer, uint8 pointer) of one image tensor under the requested entry"""
        a = z[key]
        if u8 is None:
            return _ptr(arr(a)), None
        q = quantise_u8(a)
        if u8 == "f32":      # torchvision ToTensor: uint8 -> float32, .div(255)
            return _ptr(arr(q.astype(np.float32) / np.float32(255.0))), None
        q = np.ascontiguousarray(q.transpose(0, 2, 3, 1) if u8 == "hwc" else q)
        keep.append(q)
        return None, _ptr(q)

    t = Md2Tensors()
    t.target, t.target_u8 = img("in__color__0__0")
    t.u8_hwc = int(u8 == "hwc")
    out = {"grad_T": {}, "warped": {}}
    for i, f in enumerate(srcs):
        t.source[i], t.source_u8[i] = img("in__color__%s__0" % f)
        if f == "s":
            t.T[i] = _ptr(arr(z["in__stereo_T"]))
            t.pose_requires_grad[i] = 0
        elif pose_leaves:
            # the pose leaves as PoseDecoder lays them out: (B,2,1,3), only [:, 0] used -> stride 6
            aa = np.zeros((B, 2, 1, 3), np.float32); tr = np.zeros((B, 2, 1, 3), np.float32)
            aa[:, 0] = np.asarray(z["axisangle__%s" % f]).reshape(B, 1, 3)
            tr[:, 0] = np.asarray(z["translation__%s" % f]).reshape(B, 1, 3)
            t.axisangle[i], t.translation[i] = _ptr(arr(aa)), _ptr(arr(tr))
            t.pose_stride[i], t.pose_invert[i] = 6, int(f < 0)
            t.pose_requires_grad[i] = 1
            out.setdefault("cam_T_cam", {})[f] = np.zeros((B, 4, 4), np.float32)
            out.setdefault("grad_axisangle", {})[f] = np.zeros((B, 3), np.float32)
            out.setdefault("grad_translation", {})[f] = np.zeros((B, 3), np.float32)
            t.cam_T_cam[i] = _ptr(out["cam_T_cam"][f])
            t.grad_axisangle[i] = _ptr(out["grad_axisangle"][f])
            t.grad_translation[i] = _ptr(out["grad_translation"][f])
        else:
            t.T[i] = _ptr(arr(z["cam_T_cam__%s" % f]))
            t.pose_requires_grad[i] = 1
        if not g.posecnn:
            gT = np.zeros((B, 4, 4), np.float32)
            out["grad_T"][f] = gT
            t.grad_T[i] = _ptr(gT)
    t.K = _ptr(arr(z["in__K__0"]))
    t.inv_K = _ptr(arr(z["in__inv_K__0"]))
    out["grad_disp"], out["grad_updisp"], out["idsel"], out["depth"] = [], [], [], []
    for s, L in enumerate(scales):     # s: slot, L: level; the per-scale output lists are indexed by slot
        t.disp[s] = _ptr(arr(z["disp__%d" % L]))
        t.color[s], t.color_u8[s] = img("in__color__0__%d" % L)
        if g.n_id > 0:
            t.noise[s] = _ptr(arr(z["noise__%d" % L][:, :g.n_id]))
        gd = np.zeros((B, 1, H >> L, W >> L), np.float32)
        out["grad_disp"].append(gd)
        t.grad_disp[s] = _ptr(gd)
        if g.predictive_mask:
            t.pmask[s] = _ptr(arr(z["mask__%d" % L]))
            gm = np.zeros((B, len(srcs), H >> L, W >> L), np.float32)
            out.setdefault("grad_mask", []).append(gm)
            t.grad_pmask[s] = _ptr(gm)
        if side_outputs:
            gu = np.zeros((B, 1, H, W), np.float32)
            out["grad_updisp"].append(gu)
            t.grad_depth_dbg[s] = _ptr(gu)
            m = np.zeros((B, H, W), np.float32)
            out["idsel"].append(m)
            t.identity_selection[s] = _ptr(m)
            d = np.zeros((B, 1, H, W), np.float32)
            out["depth"].append(d)
            t.depth[s] = _ptr(d)
            for i, f in enumerate(srcs):
                w = np.zeros((B, 3, H, W), np.float32)
                out["warped"][(f, L)] = w
                t.warped[i][s] = _ptr(w)
    losses = np.zeros(5, np.float32)
    t.losses = _ptr(losses)
    nbytes = C.c_size_t(0)
    st = lib.md2_emu_workspace_bytes(C.byref(p), C.byref(nbytes))
    assert st == 0, st
    ws = np.zeros(nbytes.value + 64, np.uint8)
    dbg = None
    if decisions:
        assert scales == [0, 1, 2, 3], "the decision sink is laid out for the default scales"
        n_src = len(srcs)
        dbg = dict(x0=np.zeros((4, B, n_src, H, W), np.int16), y0=np.zeros((4, B, n_src, H, W), np.int16),
                   mxy=np.zeros((4, B, n_src, H, W), np.uint8), tag=np.full((4, B, H, W), -1, np.int8),
                   live=np.zeros((4, B, n_src, 3, H, W), np.uint8), l1sgn=np.zeros((4, B, n_src, 3, H, W), np.int8))
        sizes = [B * (H >> s) * (W >> s) for s in range(4)]
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        dbg["smx"] = np.zeros(int(sum(sizes)), np.int8)
        dbg["smy"] = np.zeros(int(sum(sizes)), np.int8)
        sink = DebugSink(B=B, H=H, W=W, S=4, nsrc=n_src, x0=_ptr(dbg["x0"]), y0=_ptr(dbg["y0"]), mxy=_ptr(dbg["mxy"]),
                         tag=_ptr(dbg["tag"]), live=_ptr(dbg["live"]), l1sgn=_ptr(dbg["l1sgn"]),
                         smx=_ptr(dbg["smx"]), smy=_ptr(dbg["smy"]))
        for s in range(4):
            sink.smoff[s] = int(offs[s])
        dbg["smoff"], dbg["smsizes"] = offs, sizes
        lib.md2_emu_set_debug.argtypes = [C.POINTER(DebugSink)]
        lib.md2_emu_set_debug(C.byref(sink))
    try:
        st = lib.md2_emu_view_synthesis_loss(C.byref(p), C.byref(t), _ptr(ws), nbytes.value)
    finally:
        if decisions:
            lib.md2_emu_set_debug(None)
    assert st == 0, st
    out["decisions"] = dbg
    out["losses"] = losses
    return out
