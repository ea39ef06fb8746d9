"""GPU-side colour pyramid: Pillow's uint8 LANCZOS resize, byte for byte.
CPU: the oracle restatement (oracle/pillow_resize.py) is pinned against the installed Pillow, which is what
`transforms.Resize(..., Image.ANTIALIAS)` calls on PIL images (mono_dataset.py:57,82-86,98-103).
GPU: md2_resize_lanczos_u8 / ColorPyramid against the oracle and against Pillow itself."""
import numpy as np
import pytest

from oracle.pillow_resize import build_pyramid_u8, resize_lanczos_u8

Image = pytest.importorskip("PIL.Image")

SHAPES = [(375, 1242, 192, 640), (192, 640, 96, 320), (96, 320, 48, 160), (48, 160, 24, 80), (37, 53, 20, 31),
          (64, 64, 64, 32), (50, 40, 25, 40), (30, 30, 45, 50)]


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_is_byte_exact_with_pillow(shape):
    h, w, oh, ow = shape
    a = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(resize_lanczos_u8(a, oh, ow), ref)


def test_oracle_pyramid_follows_the_dataset_chain():
    a = np.random.default_rng(1).integers(0, 256, (375, 1242, 3), dtype=np.uint8)
    levels = build_pyramid_u8(a, 192, 640, 4)
    cur = Image.fromarray(a)
    for i, lv in enumerate(levels):                       # mono_dataset.py:98-103
        cur = cur.resize((640 >> i, 192 >> i), Image.LANCZOS)
        assert np.array_equal(lv, np.asarray(cur)), i


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("hwc", [True, False])
def test_cuda_resize_is_byte_exact(shape, hwc):
    import torch
    from monodepth2_b200.pyramid import ColorPyramid
    h, w, oh, ow = shape
    B = 3
    a = np.random.default_rng(h + w).integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    x = torch.from_numpy(a if hwc else np.ascontiguousarray(a.transpose(0, 3, 1, 2))).cuda()
    y = ColorPyramid(oh, ow, 1).resize(x, oh, ow).cpu().numpy()
    if not hwc:
        y = y.transpose(0, 2, 3, 1)
    for b in range(B):
        assert np.array_equal(y[b], resize_lanczos_u8(a[b], oh, ow)), b
        assert np.array_equal(y[b], np.asarray(Image.fromarray(a[b]).resize((ow, oh), Image.LANCZOS))), b


@pytest.mark.gpu
def test_cuda_pyramid_feeds_the_uint8_entry_of_the_fused_loss():
    """native uint8 frames -> GPU pyramid -> uint8 entry of the fused call == float entry fed with
    ToTensor(Pillow pyramid): the host-side resize + ToTensor of mono_dataset.py:98-109 can be dropped."""
    import torch
    from monodepth2_b200.fused_loss import LossPlan, view_synthesis_loss
    from monodepth2_b200.pyramid import ColorPyramid
    from monodepth2_b200.synthetic import make_batch
    B, H, W, fids = 2, 64, 96, [0, -1, 1]
    rng = np.random.default_rng(3)
    native = {f: rng.integers(0, 256, (B, 130, 200, 3), dtype=np.uint8) for f in fids}
    inputs, outputs, pose, noise = make_batch(B, H, W, fids, 4, 11, "structured")
    dev = "cuda:0"
    pyr = ColorPyramid(H, W, 4)
    ins_u8 = {k: v.to(dev) for k, v in inputs.items() if not (isinstance(k, tuple) and k[0] == "color")}
    ins_f = dict(ins_u8)
    for f in fids:
        levels = pyr(torch.from_numpy(native[f]).to(dev))
        for s, lv in enumerate(levels):
            ins_u8[("color", f, s)] = lv
            ref = np.stack([build_pyramid_u8(native[f][b], H, W, 4)[s] for b in range(B)])
            assert np.array_equal(lv.cpu().numpy(), ref), (f, s)
            ins_f[("color", f, s)] = torch.from_numpy(ref).permute(0, 3, 1, 2).float().div(255).to(dev)   # ToTensor
    plan = LossPlan(B, H, W, fids)
    res = []
    for ins in (ins_u8, ins_f):
        outs = {k: v.to(dev).clone().requires_grad_(True) for k, v in outputs.items()}
        losses = view_synthesis_loss(plan, ins, outs, [n.to(dev) for n in noise])
        losses["loss"].backward()
        res.append((losses, outs))
    assert torch.equal(res[0][0]["loss"], res[1][0]["loss"])
    for s in range(4):
        assert torch.equal(res[0][1][("disp", s)].grad, res[1][1][("disp", s)].grad)
