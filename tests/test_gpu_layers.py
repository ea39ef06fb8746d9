"""GPU parity of the per-layer drop-in ops (monodepth2_b200.layers) against the oracle's
restatement of layers.py, forward and backward, on seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import view_synthesis as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _leaf(t, dev=None):
    t = t.clone().detach()
    if dev:
        t = t.to(dev)
    return t.requires_grad_(True)


def _inputs(B=2, H=24, W=40, seed=0):
    from monodepth2_b200.synthetic import make_batch
    return make_batch(B, H, W, [0, -1, 1], 4, seed, "structured")


def test_disp_to_depth():
    from monodepth2_b200 import layers as L
    g = torch.Generator().manual_seed(0)
    d = torch.rand(2, 1, 8, 12, generator=g)
    a, b = _leaf(d), _leaf(d, DEV)
    s0, z0 = O.disp_to_depth(a, 0.1, 100.0)
    s1, z1 = L.disp_to_depth(b, 0.1, 100.0)
    np.testing.assert_allclose(z1.detach().cpu(), z0.detach(), rtol=2e-6)
    np.testing.assert_allclose(s1.detach().cpu(), s0.detach(), rtol=2e-6)
    w = torch.rand(2, 1, 8, 12, generator=g)
    (z0 * w + 0.5 * s0).sum().backward()
    (z1 * w.to(DEV) + 0.5 * s1).sum().backward()
    assert rel_l2(b.grad.cpu(), a.grad) < 1e-5


@pytest.mark.parametrize("invert", [False, True])
def test_transformation_from_parameters(invert):
    from monodepth2_b200 import layers as L
    g = torch.Generator().manual_seed(1)
    aa = 0.3 * torch.randn(5, 1, 3, generator=g)
    tr = torch.randn(5, 1, 3, generator=g)
    a0, t0, a1, t1 = _leaf(aa), _leaf(tr), _leaf(aa, DEV), _leaf(tr, DEV)
    T0 = O.transformation_from_parameters(a0, t0, invert)
    T1 = L.transformation_from_parameters(a1, t1, invert)
    np.testing.assert_allclose(T1.detach().cpu(), T0.detach(), atol=2e-6)
    w = torch.randn(5, 4, 4, generator=g)
    (T0 * w).sum().backward()
    (T1 * w.to(DEV)).sum().backward()
    assert rel_l2(a1.grad.cpu(), a0.grad) < 1e-4
    assert rel_l2(t1.grad.cpu(), t0.grad) < 1e-5
    np.testing.assert_allclose(L.rot_from_axisangle(aa.to(DEV)).cpu(), O.rot_from_axisangle(aa), atol=2e-6)
    np.testing.assert_allclose(L.get_translation_matrix(tr.to(DEV)).cpu(), O.translation_matrix(tr), atol=0)


def test_backproject_project_gridsample_chain():
    from monodepth2_b200 import layers as L
    B, H, W = 2, 24, 40
    inputs, outputs, pose, _ = _inputs(B, H, W)
    K, iK = inputs[("K", 0)], inputs[("inv_K", 0)]
    img = inputs[("color", -1, 0)]
    depth = 1.0 / (0.01 + 9.99 * outputs[("disp", 0)])
    T = outputs[("cam_T_cam", 0, -1)]
    d0, T0 = _leaf(depth), _leaf(T)
    d1, T1 = _leaf(depth, DEV), _leaf(T, DEV)
    pts0 = O.backproject(d0, iK)
    grid0 = O.project(pts0, K, T0, H, W)
    out0 = F.grid_sample(img, grid0, mode="bilinear", padding_mode="border", align_corners=False)
    bp, pj = L.BackprojectDepth(B, H, W).to(DEV), L.Project3D(B, H, W).to(DEV)
    pts1 = bp(d1, iK.to(DEV))
    grid1 = pj(pts1, K.to(DEV), T1)
    out1 = L.grid_sample_border(img.to(DEV), grid1)
    np.testing.assert_allclose(pts1.detach().cpu(), pts0.detach(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grid1.detach().cpu(), grid0.detach(), atol=2e-5)
    np.testing.assert_allclose(out1.detach().cpu(), out0.detach(), atol=2e-4)
    w = torch.rand(out0.shape, generator=torch.Generator().manual_seed(3))
    (out0 * w).sum().backward()
    (out1 * w.to(DEV)).sum().backward()
    assert rel_l2(d1.grad.cpu(), d0.grad) < 2e-2      # bilinear-cell flips allowed
    assert rel_l2(T1.grad.cpu(), T0.grad) < 2e-2
    frac = ((d1.grad.cpu() - d0.grad).abs() <= 1e-4 * d0.grad.abs().max()).float().mean()
    assert frac >= 0.995


def test_grid_sample_align_corners_true():
    from monodepth2_b200 import layers as L
    g = torch.Generator().manual_seed(4)
    img = torch.rand(2, 3, 10, 14, generator=g)
    grid = torch.rand(2, 6, 7, 2, generator=g) * 2.4 - 1.2
    for ac in (False, True):
        g0 = _leaf(grid); g1 = _leaf(grid, DEV)
        o0 = F.grid_sample(img, g0, mode="bilinear", padding_mode="border", align_corners=ac)
        o1 = L.grid_sample_border(img.to(DEV), g1, align_corners=ac)
        np.testing.assert_allclose(o1.detach().cpu(), o0.detach(), atol=2e-6)
        o0.sum().backward(); o1.sum().backward()
        np.testing.assert_allclose(g1.grad.cpu(), g0.grad, atol=2e-5)


def test_ssim_forward_backward():
    from monodepth2_b200 import layers as L
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 12, 17, generator=g)
    y = (x + 0.1 * torch.randn(2, 3, 12, 17, generator=g)).clamp(0, 1)
    x0, y0, x1, y1 = _leaf(x), _leaf(y), _leaf(x, DEV), _leaf(y, DEV)
    s0 = O.ssim_dissimilarity(x0, y0)
    s1 = L.SSIM().to(DEV)(x1, y1)
    np.testing.assert_allclose(s1.detach().cpu(), s0.detach(), atol=3e-6)
    w = torch.rand(s0.shape, generator=g)
    (s0 * w).sum().backward()
    (s1 * w.to(DEV)).sum().backward()
    assert rel_l2(x1.grad.cpu(), x0.grad) < 1e-4
    assert rel_l2(y1.grad.cpu(), y0.grad) < 1e-4


def test_get_smooth_loss():
    from monodepth2_b200 import layers as L
    g = torch.Generator().manual_seed(6)
    d = torch.rand(2, 1, 12, 20, generator=g)
    img = torch.rand(2, 3, 12, 20, generator=g)
    d0, d1 = _leaf(d), _leaf(d, DEV)
    l0 = O.smooth_loss(d0, img)
    l1 = L.get_smooth_loss(d1, img.to(DEV))
    assert abs(float(l1) - float(l0)) <= 1e-6 * abs(float(l0))
    (3.0 * l0).backward(); (3.0 * l1).backward()
    assert rel_l2(d1.grad.cpu(), d0.grad) < 1e-5


def test_cpu_tensor_raises():
    from monodepth2_b200 import layers as L
    with pytest.raises(RuntimeError):
        L.SSIM()(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
