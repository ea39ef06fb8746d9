"""The oracle (oracle/view_synthesis.py) against every golden vector made by the reference."""
import numpy as np
import pytest
import torch

from helpers import Golden, golden_cases, rel_l2, run_oracle


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(name):
    g = Golden(name)
    r = run_oracle(g)
    z = g.z
    # losses: the oracle calls the same library kernels in the same order -> ~bit-equal
    assert abs(float(r["losses"]["loss"]) - float(z["loss"])) <= 1e-6 * abs(float(z["loss"]))
    for s in g.scales:
        assert abs(float(r["losses"]["loss/%d" % s]) - float(z["loss__%d" % s])) <= 1e-6 * abs(float(z["loss__%d" % s]))
    # side outputs
    np.testing.assert_allclose(r["outs"][("depth", 0, 0)].detach().numpy(), z["depth__0"], rtol=1e-6, atol=0)
    for f in g.frame_ids[1:]:
        np.testing.assert_allclose(r["outs"][("color", f, 0)].detach().numpy(), z["color__%s__0" % f],
                                   rtol=0, atol=2e-6)
    if not g.disable_automasking:
        for s in g.scales:
            m = r["outs"]["identity_selection/%d" % s].numpy().astype(np.uint8)
            assert (m != z["idsel__%d" % s]).mean() <= 1e-4
    # gradients (fp32 vs fp32, same kernels: tight)
    for s in g.scales:
        assert rel_l2(r["leaves"][("disp", s)].grad, z["grad_disp__%d" % s]) < 1e-4
        assert rel_l2(r["outs"][("depth", 0, s)].grad, z["grad_depth__%d" % s]) < 1e-4
    if g.predictive_mask:
        for s in g.scales:
            assert rel_l2(r["leaves"][("mask", s)].grad, z["grad_mask__%d" % s]) < 1e-4
    for f in g.frame_ids[1:]:
        if f == "s":
            continue
        if not g.posecnn:      # under posecnn T is rebuilt per scale from axisangle / translation
            assert rel_l2(r["outs"][("cam_T_cam", 0, f)].grad, z["grad_cam_T_cam__%s" % f]) < 1e-4
        assert rel_l2(r["leaves"][("axisangle", f)].grad, z["grad_axisangle__%s" % f]) < 1e-3
        assert rel_l2(r["leaves"][("translation", f)].grad, z["grad_translation__%s" % f]) < 1e-3


def test_oracle_fp64_close_to_fp32():
    g = Golden("mono_structured")
    r32 = run_oracle(g, torch.float32)
    r64 = run_oracle(g, torch.float64)
    assert abs(float(r32["losses"]["loss"]) - float(r64["losses"]["loss"])) < 1e-6
