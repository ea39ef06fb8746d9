"""Host emulation of the marching kernels (same per-lane source as the CUDA build, g++-compiled,
tests/emu) against the reference's golden vectors: checks tiling, halos, reflection and the
adjoint without a GPU.  The emulator is test infrastructure, never a product path."""
import numpy as np
import pytest

from helpers import Golden, golden_cases, rel_l2


@pytest.mark.parametrize("rows", [16, 5])
@pytest.mark.parametrize("name", [c for c in golden_cases() if c != "v1_multiscale"])  # (one call per level: host-side)
def test_emulated_kernels_match_reference_golden(name, rows):
    from emu_driver import run_emu
    g = Golden(name)
    z = g.z
    o = run_emu(g, rows_per_segment=rows)
    assert abs(o["losses"][0] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for i, s in enumerate(g.scales):     # i: slot of the C ABI arrays, s: pyramid level (--scales subsets)
        assert abs(o["losses"][1 + i] - float(z["loss__%d" % s])) <= 1e-5 * abs(float(z["loss__%d" % s]))
        assert rel_l2(o["grad_disp"][i], z["grad_disp__%d" % s]) < 8e-2
    np.testing.assert_allclose(o["depth"][0], z["depth__0"], rtol=2e-6)
    for f in g.frame_ids[1:]:
        np.testing.assert_allclose(o["warped"][(f, 0)], z["color__%s__0" % f], atol=5e-5)
        if f == "s":
            continue
        if g.posecnn:     # kernel variant (md2_problem.posecnn): T per scale from the leaves, gradient on the leaves
            assert rel_l2(o["grad_axisangle"][f].reshape(-1), z["grad_axisangle__%s" % f].reshape(-1)) < 8e-2
            assert rel_l2(o["grad_translation"][f].reshape(-1), z["grad_translation__%s" % f].reshape(-1)) < 8e-2
        else:
            assert rel_l2(o["grad_T"][f], z["grad_cam_T_cam__%s" % f]) < 8e-2
    if g.predictive_mask:   # kernel variant (md2_problem.predictive_mask)
        for i, s in enumerate(g.scales):
            assert rel_l2(o["grad_mask"][i], z["grad_mask__%d" % s]) < 1e-3, s
    if g.n_id > 0:
        for i, s in enumerate(g.scales):
            assert (o["idsel"][i].astype(np.uint8) != z["idsel__%d" % s]).mean() <= 5e-4


def test_emulator_forward_only():
    from emu_driver import run_emu
    g = Golden("mono_iid")
    o = run_emu(g, want_grad=False, side_outputs=False)
    assert abs(o["losses"][0] - float(g.z["loss"])) <= 1e-5 * abs(float(g.z["loss"]))


@pytest.mark.parametrize("name", ["mono_structured", "stereo_iid"])
def test_emulator_pose_leaves(name):
    """The in-call pose construction on the host build of the same source: T built from
    axisangle / translation equals the reference's cam_T_cam, and the pose gradient on the leaves equals the
    matrix-mode gradient pushed through the oracle's transformation_from_parameters."""
    import torch
    from emu_driver import run_emu
    from oracle import view_synthesis as O
    g = Golden(name)
    z = g.z
    o = run_emu(g, pose_leaves=True, rows_per_segment=16)
    m = run_emu(g, rows_per_segment=16)
    assert abs(o["losses"][0] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    for f in g.frame_ids[1:]:
        if f == "s":
            continue
        np.testing.assert_allclose(o["cam_T_cam"][f], z["cam_T_cam__%s" % f], atol=2e-6)
        np.testing.assert_allclose(o["grad_T"][f], m["grad_T"][f], rtol=1e-5, atol=1e-9)
        aa = torch.from_numpy(np.asarray(z["axisangle__%s" % f])).reshape(g.B, 1, 3).requires_grad_(True)
        tr = torch.from_numpy(np.asarray(z["translation__%s" % f])).reshape(g.B, 1, 3).requires_grad_(True)
        O.transformation_from_parameters(aa, tr, f < 0).backward(torch.from_numpy(m["grad_T"][f]))
        assert rel_l2(o["grad_axisangle"][f].reshape(-1), aa.grad.reshape(-1)) < 1e-4
        assert rel_l2(o["grad_translation"][f].reshape(-1), tr.grad.reshape(-1)) < 1e-5
