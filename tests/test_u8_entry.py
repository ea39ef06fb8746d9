"""uint8 entry of the fused call (first slice): frames handed over as the bytes the dataloader
holds before ToTensor (datasets/mono_dataset.py:106-109) must give bit-identical results to the float entry fed
with ToTensor of the same bytes.  CPU: through the host emulator (same load_px / kernel arithmetic); GPU: through
the C ABI."""
import numpy as np
import pytest
import torch

from helpers import Golden
from emu_driver import run_emu, quantise_u8


@pytest.mark.parametrize("case", ["mono_structured", "stereo_iid", "disable_automasking", "scales_0_2"])
@pytest.mark.parametrize("layout", ["hwc", "chw"])
def test_emulator_u8_entry_is_bit_identical_to_float_entry(case, layout):
    g = Golden(case)
    ref = run_emu(g, u8="f32")
    got = run_emu(g, u8=layout)
    assert np.array_equal(ref["losses"], got["losses"])
    for s in range(len(g.scales)):
        assert np.array_equal(ref["grad_disp"][s], got["grad_disp"][s])
        assert np.array_equal(ref["idsel"][s], got["idsel"][s])
    for f in ref["grad_T"]:
        assert np.array_equal(ref["grad_T"][f], got["grad_T"][f])


def _cuda_run(g, mode, dev="cuda:0"):
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    plan = LossPlan(g.B, g.H, g.W, g.frame_ids, avg_reprojection=g.avg_reprojection,
                    disable_automasking=g.disable_automasking, no_ssim=g.no_ssim, scales=g.scales)
    inputs = {}
    for k, v in g.inputs().items():
        if isinstance(k, tuple) and k[0] == "color":
            q = torch.from_numpy(quantise_u8(v.numpy()))
            if mode == "f32":
                v = q.float().div(255)          # ToTensor
            elif mode == "hwc":
                v = q.permute(0, 2, 3, 1).contiguous()
            else:
                v = q
        inputs[k] = v.to(dev)
    outs = {}
    for s in g.scales:
        outs[("disp", s)] = g.t("disp__%d" % s).to(dev).requires_grad_(True)
    for f in g.frame_ids[1:]:
        if f != "s":
            outs[("cam_T_cam", 0, f)] = g.t("cam_T_cam__%s" % f).to(dev).requires_grad_(True)
    noise = [n.to(dev) for n in g.noise()] if g.n_id > 0 else None
    side = {"mask_scales": list(g.scales)} if g.n_id > 0 else None
    losses = view_synthesis_loss(plan, inputs, outs, noise, side)
    losses["loss"].backward()
    torch.cuda.synchronize()
    return losses, outs, side


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mono_iid", "mono_structured", "stereo_iid", "avg_reprojection", "disable_automasking",
                                  "scales_0_2"])
@pytest.mark.parametrize("layout", ["hwc", "chw"])
def test_cuda_u8_entry_is_bit_identical_to_float_entry(case, layout):
    g = Golden(case)
    l0, o0, s0 = _cuda_run(g, "f32")
    l1, o1, s1 = _cuda_run(g, layout)
    for k in l0:
        assert torch.equal(l0[k], l1[k]), k
    if s0 is not None:
        for s in g.scales:
            assert torch.equal(s0["identity_selection/%d" % s], s1["identity_selection/%d" % s])
    # the scalar sums are accumulated with fp64 atomics (order varies run to run): per-pixel gradients are exact
    for s in g.scales:
        assert torch.equal(o0[("disp", s)].grad, o1[("disp", s)].grad), s


def test_device_form_of_the_byte_to_unit_conversion_is_the_exact_quotient():
    """md2_core.cuh u8_unit, device branch: q = v * rn(1/255); q' = fma(fma(-q, 255, v), rn(1/255), q) must equal the
    correctly rounded v / 255 (torchvision's ToTensor) for all 256 byte values.  Emulated here in float64, in which
    every product of two fp32 values and the fused operands are exact before the single rounding of each fma."""
    v = np.arange(256, dtype=np.float32)
    exact = (v / np.float32(255.0)).astype(np.float32)
    r = np.float32(0.003921568859368563)
    assert r == np.float32(1.0) / np.float32(255.0)
    q = (v * r).astype(np.float32)
    rem = (v.astype(np.float64) - q.astype(np.float64) * 255.0).astype(np.float32)          # fma(-q, 255, v)
    q2 = (q.astype(np.float64) + rem.astype(np.float64) * np.float64(r)).astype(np.float32)  # fma(rem, r, q)
    assert (q != exact).sum() > 0          # the plain product is NOT exact: the correction is needed
    assert np.array_equal(q2, exact)
