"""--scales subsets (options.py:64, e.g. ``--scales 0 2``): trainer.py:345,413 iterate opt.scales, the decoder emits
``("disp", s)`` for exactly those levels and the dataloader holds levels 0..3 regardless (trainer.py:127-135).  The C ABI
takes the levels as md2_problem.scale_level (ascending); level 0 has to be in the list, as in the reference
(trainer.py:151-159,377 index backproject_depth[0]).

Golden vectors from the unmodified reference: tests/golden/scales_*.npz (picked up by test_oracle_golden.py,
test_emu_parity.py and test_gpu_parity.py like every other fixture).  Here: the host-side contract without a GPU, and on
the GPU the size-independent property that a subset call equals the matching part of the full call."""
import ctypes as C

import pytest
import torch


def test_plan_keeps_levels_ascending_and_rejects_bad_lists():
    from monodepth2_b200.fused_loss import LossPlan
    p = LossPlan(2, 48, 80, [0, -1, 1], scales=[2, 0])
    assert p.scales == [0, 2] and p.slot(2) == 1
    # the reference draws the tie-break noise in the order of the list (trainer.py:413,468-469): the host side draws in
    # that order and hands the draws over in slot order, so a seeded run sees the reference's values per scale
    from monodepth2_b200.fused_loss import _in_slot_order
    assert p.scales_given == [2, 0] and _in_slot_order(p, ["first draw", "second draw"]) == ["second draw", "first draw"]
    q = LossPlan(2, 48, 80, [0, -1, 1], scales=[0, 1, 3])
    drawn = ["a", "b", "c"]
    assert _in_slot_order(q, drawn) is drawn
    assert list(p.problem(True).scale_level) == [0, 2, 0, 0] and p.problem(True).num_scales == 2
    d = LossPlan(2, 48, 80, [0, -1, 1])
    assert list(d.problem(True).scale_level) == [0, 0, 0, 0] and d.problem(True).num_scales == 4
    assert list(LossPlan(2, 48, 80, [0, -1, 1], scales=[0]).problem(False).scale_level) == [0, 0, 0, 0]
    for bad in ([0, 0], [0, 4], [-1, 0], []):
        with pytest.raises(RuntimeError):
            LossPlan(2, 48, 80, [0, -1, 1], scales=bad)


def test_c_abi_validates_scale_levels():
    from monodepth2_b200 import _capi
    from monodepth2_b200._capi import Md2Problem
    lib = _capi.load_library()
    n, n4 = C.c_size_t(0), C.c_size_t(0)

    def prob(ns, levels, h=192):
        return Md2Problem(batch=12, height=h, width=640, num_scales=ns, num_src=2, automask=1, min_depth=0.1,
                          max_depth=100.0, disparity_smoothness=1e-3, want_grad=1, scale_level=(C.c_int * 4)(*levels))
    assert lib.md2_loss_workspace_bytes(C.byref(prob(4, [0, 0, 0, 0])), C.byref(n4)) == 0
    assert lib.md2_loss_workspace_bytes(C.byref(prob(4, [0, 1, 2, 3])), C.byref(n)) == 0 and n.value == n4.value
    assert lib.md2_loss_workspace_bytes(C.byref(prob(2, [0, 2, 0, 0])), C.byref(n)) == 0 and n.value < n4.value
    assert lib.md2_loss_workspace_bytes(C.byref(prob(2, [0, 3, 0, 0], h=196)), C.byref(n)) == -1   # 196 % 8 != 0
    assert lib.md2_loss_workspace_bytes(C.byref(prob(2, [0, 2, 0, 0], h=196)), C.byref(n)) == 0    # 196 % 4 == 0
    for bad in ([2, 0, 0, 0], [0, 2, 2, 0], [0, 4, 0, 0], [1, 2, 0, 0]):      # descending, duplicate, level 4, no level 0
        assert lib.md2_loss_workspace_bytes(C.byref(prob(sum(1 for i, x in enumerate(bad) if x or i == 0), bad)),
                                            C.byref(n)) == -1, bad


def _run(batch, fids, scales, u8=False, **kw):
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    inputs, outputs, _pose, noise = batch
    dev = "cuda:0"
    B, _, H, W = inputs[("color", 0, 0)].shape
    plan = LossPlan(B, H, W, fids, scales=scales, **kw)
    ins = {}
    for k, v in inputs.items():
        if u8 and isinstance(k, tuple) and k[0] == "color":
            v = (v * 255).round().clamp(0, 255).to(torch.uint8)
        ins[k] = v.to(dev)
    outs = {k: v.to(dev).requires_grad_(True) for k, v in outputs.items()
            if k[0] == "cam_T_cam" or (k[0] == "disp" and k[1] in scales)}
    nz = [noise[s][:, :plan.n_id].contiguous().to(dev) for s in sorted(scales)] if plan.n_id > 0 else None
    side = {"mask_scales": sorted(scales)} if plan.n_id > 0 else {}
    losses = view_synthesis_loss(plan, ins, outs, nz, side)
    losses["loss"].backward()
    torch.cuda.synchronize()
    return losses, outs, side


@pytest.mark.gpu
@pytest.mark.parametrize("scales", [[0, 2], [0, 1, 3], [0, 3], [0]])
@pytest.mark.parametrize("fids,u8", [([0, -1, 1], False), ([0, -1, 1, "s"], False), ([0, -1, 1], True)])
def test_subset_call_equals_the_matching_part_of_the_full_call(scales, fids, u8):
    """BASELINE size (640x192, batch 12).  Per scale the path is independent of the other scales (trainer.py:413-492):
    loss/<s> is the same number, the masks are the same, and d loss / d disp_s differs only by the 1 / len(scales) of
    trainer.py:494, which enters the kernels as one fp32 factor (gscale) - so the gradients agree to a few ulp after
    rescaling.  (The pose gradient sums over the scales of the call: next test.)"""
    from monodepth2_b200.synthetic import make_batch
    batch = make_batch(12, 192, 640, fids, 4, 31, "structured", n_id=len(fids) - 1)
    lf, of, sf = _run(batch, fids, [0, 1, 2, 3], u8)
    ls, os_, ss = _run(batch, fids, scales, u8)
    k = 4.0 / len(scales)
    for s in scales:
        a, b = float(ls["loss/%d" % s]), float(lf["loss/%d" % s])
        assert abs(a - b) <= 2e-6 * abs(b), (s, a, b)       # fp64 atomics: the order of the sums varies run to run
        assert torch.equal(ss["identity_selection/%d" % s], sf["identity_selection/%d" % s]), s
        g1, g0 = os_[("disp", s)].grad, of[("disp", s)].grad * k
        assert float((g1 - g0).abs().max()) <= 1e-5 * float(g0.abs().max()), s
    want = sum(float(ls["loss/%d" % s]) for s in scales) / len(scales)
    assert abs(float(ls["loss"]) - want) <= 2e-6 * abs(want)
    assert sorted(k_ for k_ in ls if k_.startswith("loss/")) == ["loss/%d" % s for s in scales]


@pytest.mark.gpu
def test_pose_gradient_is_additive_over_disjoint_scale_subsets():
    """With L_S the loss of the call over subset S (a mean over its scales, trainer.py:494):
    4 L_{0123} = 2 L_{02} + 3 L_{013} - L_{0}.  The same holds for d loss / d cam_T_cam, which the kernels accumulate
    per call in fp64."""
    from monodepth2_b200.synthetic import make_batch
    fids = [0, -1, 1]
    batch = make_batch(6, 192, 640, fids, 4, 32, "structured", n_id=2)
    g = {}
    for key, sc in (("full", [0, 1, 2, 3]), ("a", [0, 2]), ("b", [0, 1, 3]), ("c", [0])):
        _l, o, _s = _run(batch, fids, sc)
        g[key] = {f: o[("cam_T_cam", 0, f)].grad.double() for f in fids[1:]}
    for f in fids[1:]:
        want = 4.0 * g["full"][f]
        got = 2.0 * g["a"][f] + 3.0 * g["b"][f] - g["c"][f]
        assert float((got - want).norm() / want.norm()) <= 1e-5, f


def test_a_list_without_level_0_is_rejected_like_the_reference_does():
    """No level 0 in the list: the reference fails with KeyError (backproject_depth[0], trainer.py:377); the plan raises
    before anything is launched (and the C ABI answers MD2_ERR_INVALID_ARGUMENT, above).  --v1_multiscale warps at the
    scale itself (trainer.py:347-348), so there any levels do."""
    from monodepth2_b200.fused_loss import LossPlan
    with pytest.raises(RuntimeError):
        LossPlan(2, 48, 80, [0, -1, 1], scales=[1, 2])
    assert LossPlan(2, 48, 80, [0, -1, 1], scales=[1, 2], v1_multiscale=True).scales == [1, 2]
