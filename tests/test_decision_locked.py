"""Protocol P4 of SURVEY.md 8(c): decision-locked exactness.

(a) The fp64 decision-locked restatement (oracle/decision_locked.py) with its OWN decisions reproduces
    the plain fp64 oracle's autograd gradients to ~1e-12 (it is the same function).
(b) Fed with the discrete decisions exported by the host emulator of the kernels (same per-lane code as
    the CUDA build), it reproduces the kernel arithmetic's aggregated gradients (disp_s, cam_T_cam) to
    <= 1e-4 relative L2: what separates the kernels from the reference is decision flips, not arithmetic.
"""
import numpy as np
import pytest
import torch

from helpers import Golden, rel_l2, run_oracle
from oracle import decision_locked as DL
from oracle import view_synthesis as O

CASES = ["mono_iid", "mono_structured", "stereo_structured", "avg_reprojection", "disable_automasking", "no_ssim"]


def _leaves64(g):
    inputs = g.inputs(torch.float64)
    outs, lv = {}, {}
    for s in range(4):
        lv[("disp", s)] = g.t("disp__%d" % s, torch.float64).requires_grad_(True)
        outs[("disp", s)] = lv[("disp", s)]
    for f in g.frame_ids[1:]:
        if f != "s":
            lv[("T", f)] = g.t("cam_T_cam__%s" % f, torch.float64).requires_grad_(True)
            outs[("cam_T_cam", 0, f)] = lv[("T", f)]
    return inputs, outs, lv


@pytest.mark.parametrize("name", CASES)
def test_locked_restatement_with_own_decisions_equals_fp64_oracle(name):
    g = Golden(name)
    cfg = g.cfg()
    inputs, outs, lv = _leaves64(g)
    DL.locked_loss(inputs, outs, cfg, g.noise(torch.float64))["loss"].backward()
    inputs2, outs2, lv2 = _leaves64(g)
    O.view_synthesis_loss(inputs2, outs2, cfg, g.noise(torch.float64))["loss"].backward()
    for k in lv:
        assert rel_l2(lv[k].grad, lv2[k].grad) < 1e-11, k


def _check_locked(g, d, grad_disp, grad_T):
    """fp64 restatement fed with the exported decisions `d` against the aggregated gradients that came with them"""
    srcs = g.frame_ids[1:]
    B, H, W = g.B, g.H, g.W
    dec = dict(x0={}, y0={}, mx={}, my={}, tag={}, live={}, l1sgn={}, smx={}, smy={})
    for s in range(4):
        for fi, f in enumerate(srcs):
            dec["x0"][(s, f)] = torch.from_numpy(d["x0"][s, :, fi].astype(np.int64))
            dec["y0"][(s, f)] = torch.from_numpy(d["y0"][s, :, fi].astype(np.int64))
            dec["mx"][(s, f)] = torch.from_numpy((d["mxy"][s, :, fi] & 1) > 0)
            dec["my"][(s, f)] = torch.from_numpy((d["mxy"][s, :, fi] & 2) > 0)
            dec["live"][(s, f)] = torch.from_numpy(d["live"][s, :, fi] > 0)
            dec["l1sgn"][(s, f)] = torch.from_numpy(d["l1sgn"][s, :, fi].astype(np.float64))
        dec["tag"][s] = torch.from_numpy(d["tag"][s].astype(np.int64))
        hs, ws = H >> s, W >> s
        off, n = int(d["smoff"][s]), d["smsizes"][s]
        sx = torch.from_numpy(d["smx"][off:off + n].astype(np.float64)).reshape(B, 1, hs, ws)
        sy = torch.from_numpy(d["smy"][off:off + n].astype(np.float64)).reshape(B, 1, hs, ws)
        dec["smx"][s], dec["smy"][s] = sx[..., :, :-1], sy[..., :-1, :]
    cfg = g.cfg()
    inputs, outs, lv = _leaves64(g)
    DL.locked_loss(inputs, outs, cfg, g.noise(torch.float64), decisions=dec)["loss"].backward()
    for s in range(4):
        assert rel_l2(grad_disp[s], lv[("disp", s)].grad) <= 1e-4, ("disp", s)
    for f in srcs:
        if f != "s":
            ref = lv[("T", f)].grad
            got = torch.as_tensor(grad_T[f]).double()
            assert rel_l2(got, ref) <= 1e-4, ("T", f)
            # per element relative to the tensor's max (the 12 pose-relevant entries)
            assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-4


@pytest.mark.parametrize("name", CASES)
def test_kernel_gradients_are_exact_given_their_decisions(name):
    from emu_driver import run_emu
    g = Golden(name)
    o = run_emu(g, rows_per_segment=16, decisions=True)
    _check_locked(g, o["decisions"], o["grad_disp"], o["grad_T"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("rows", [0, 16])
def test_cuda_kernel_gradients_are_exact_given_their_decisions(name, rows):
    """P4 on the REAL kernels.  The debug build of the library (-DMD2_DBG_DEVICE: same sources,
    the role-specialised kernels write the decisions they take into device arrays) runs the fused call; the fp64
    restatement fed with those decisions must reproduce the CUDA gradients to 1e-4.  The same build checks every
    global index of the marching path (-DMD2_BOUNDS_CHECK): none may be out of range."""
    from dbg_driver import DecisionSink, debug_lib, oob_count
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    g = Golden(name)
    dev = "cuda:0"
    plan = LossPlan(g.B, g.H, g.W, g.frame_ids, avg_reprojection=g.avg_reprojection,
                    disable_automasking=g.disable_automasking, no_ssim=g.no_ssim, rows_per_segment=rows)
    plan.lib = debug_lib()
    inputs = {k: v.to(dev) for k, v in g.inputs().items()}
    outs = {}
    for s in range(4):
        outs[("disp", s)] = g.t("disp__%d" % s).to(dev).requires_grad_(True)
    for f in g.frame_ids[1:]:
        if f != "s":
            outs[("cam_T_cam", 0, f)] = g.t("cam_T_cam__%s" % f).to(dev).requires_grad_(True)
    noise = [n.to(dev) for n in g.noise()] if g.n_id > 0 else None
    oob_count(reset=True)
    with DecisionSink(g.B, g.H, g.W, len(g.frame_ids) - 1, dev) as sink:
        losses = view_synthesis_loss(plan, inputs, outs, noise)
        losses["loss"].backward()
    assert oob_count() == 0
    ref = float(g.z["loss"])
    assert abs(float(losses["loss"].detach()) - ref) <= 1e-5 * abs(ref)
    _check_locked(g, sink.numpy(), [outs[("disp", s)].grad.cpu() for s in range(4)],
                  {f: outs[("cam_T_cam", 0, f)].grad.cpu() for f in g.frame_ids[1:] if f != "s"})

