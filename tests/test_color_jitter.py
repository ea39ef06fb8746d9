"""Colour augmentation (/root/reference/datasets/mono_dataset.py:60-70,136,169-176): the numpy oracle
against the installed Pillow / torchvision (CPU), the CUDA kernels against the oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import color_jitter as CJ

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def _img(seed, h=37, w=53, kind="iid"):
    rng = np.random.default_rng(seed)
    if kind == "iid":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w]
    base = (120 + 100 * np.sin(x / 7.0 + seed) * np.cos(y / 5.0))[..., None] + rng.normal(0, 12, (h, w, 3))
    return np.clip(base + np.array([20, -10, 5]), 0, 255).astype(np.uint8)


def _all_colours():
    a = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(a, a, a, indexing="ij")
    return np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))


def test_hsv_conversions_match_pillow_on_every_colour():
    c = _all_colours()
    assert np.array_equal(CJ.rgb2hsv(c), np.array(Image.fromarray(c, "RGB").convert("HSV")))
    assert np.array_equal(CJ.hsv2rgb(c), np.array(Image.fromarray(c, "HSV").convert("RGB")))


@pytest.mark.parametrize("seed", range(6))
def test_adjustments_match_torchvision_on_pil_images(seed):
    F = pytest.importorskip("torchvision.transforms.functional")
    rng = np.random.default_rng(100 + seed)
    img = _img(seed, kind="iid" if seed % 2 else "structured")
    pil = Image.fromarray(img)
    for fac in list(rng.uniform(0.8, 1.2, 4)) + [0.0, 1.0, 0.5, 1.7]:
        assert np.array_equal(CJ.adjust_brightness(img, fac), np.array(F.adjust_brightness(pil, fac))), ("brightness", fac)
        assert np.array_equal(CJ.adjust_contrast(img, fac), np.array(F.adjust_contrast(pil, fac))), ("contrast", fac)
        assert np.array_equal(CJ.adjust_saturation(img, fac), np.array(F.adjust_saturation(pil, fac))), ("saturation", fac)
    for hf in list(rng.uniform(-0.1, 0.1, 4)) + [0.0, 0.5, -0.5, 0.004]:
        assert np.array_equal(CJ.adjust_hue(img, hf), np.array(F.adjust_hue(pil, hf))), ("hue", hf)


@pytest.mark.parametrize("seed", range(4))
def test_jitter_chain_matches_torchvision_colorjitter(seed):
    T = pytest.importorskip("torchvision.transforms")
    F = pytest.importorskip("torchvision.transforms.functional")
    torch.manual_seed(seed)
    fn_idx, b, c, s, h = T.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))   # mono_dataset.py:60-70
    img = _img(10 + seed, 64, 96, "structured")
    pil = Image.fromarray(img)
    for i in fn_idx:                                   # torchvision ColorJitter.forward
        pil = [F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue][int(i)](pil, [b, c, s, h][int(i)])
    assert np.array_equal(CJ.color_jitter(img, fn_idx.tolist(), b, c, s, h), np.array(pil))


# ------------------------------------------------------------------ GPU: the kernels against the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("hwc", [True, False])
def test_cuda_color_jitter_is_byte_exact(hwc):
    from monodepth2_b200.pyramid import ColorAug
    rng = np.random.default_rng(7)
    H, W = 96, 160
    imgs, params = [], []
    orders = [[0, 1, 2, 3], [3, 2, 1, 0], [2, 1, 3, 0], [1, 0, 3, 2], [3, 0, 2, 1], [0, 2, 3, 1], [1, 3, 0, 2], [2, 0, 1, 3]]
    for n, order in enumerate(orders):
        imgs.append(_img(n, H, W, "iid" if n % 2 else "structured"))
        b, c, s = (float(v) for v in rng.uniform(0.8, 1.2, 3))
        h = float(rng.uniform(-0.1, 0.1))
        if n == 5:
            c = None                                    # a skipped adjustment (torchvision passes None through)
        if n == 6:
            b, s, h = 1.7, 0.0, -0.5                    # extrapolating blend, full desaturation, largest hue shift
        params.append((order, b, c, s, h))
    # the whole colour cube through the hue path once (every RGB value), and a flat image (contrast mean = the level)
    cube = np.ascontiguousarray(np.stack(np.meshgrid(*[np.arange(0, 256, 5, dtype=np.uint8)] * 3, indexing="ij"), -1).reshape(-1, 3))
    pad = np.zeros((H * W, 3), np.uint8); pad[:min(len(cube), H * W)] = cube[:H * W]
    imgs.append(pad.reshape(H, W, 3)); params.append(([3, 1, 0, 2], 1.1, 0.9, 1.15, 0.07))
    imgs.append(np.full((H, W, 3), 77, np.uint8)); params.append(([1, 3, 2, 0], 0.85, 1.2, 0.8, 0.031))
    x = np.stack(imgs)
    want = np.stack([CJ.color_jitter(im, *p) for im, p in zip(imgs, params)])
    xt = torch.from_numpy(x if hwc else np.ascontiguousarray(x.transpose(0, 3, 1, 2))).cuda()
    got = ColorAug()(xt, params).cpu().numpy()
    if not hwc:
        got = got.transpose(0, 2, 3, 1)
    bad = (got != want).any(-1)
    assert not bad.any(), (int(bad.sum()), [int(i) for i in np.unique(np.argwhere(bad)[:, 0])])


@pytest.mark.gpu
def test_cuda_color_jitter_matches_torchvision_colorjitter_draws():
    """End to end as the dataset does it: parameters from ColorJitter.get_params with the reference's ranges
    (mono_dataset.py:60-70), torchvision's PIL path as the checker."""
    T = pytest.importorskip("torchvision.transforms")
    F = pytest.importorskip("torchvision.transforms.functional")
    from monodepth2_b200.pyramid import ColorAug
    imgs, params, want = [], [], []
    for n in range(6):
        torch.manual_seed(40 + n)
        fn_idx, b, c, s, h = T.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
        im = _img(50 + n, 192, 640, "structured")
        pil = Image.fromarray(im)
        for i in fn_idx:
            pil = [F.adjust_brightness, F.adjust_contrast, F.adjust_saturation, F.adjust_hue][int(i)](pil, [b, c, s, h][int(i)])
        imgs.append(im); params.append((fn_idx, b, c, s, h)); want.append(np.array(pil))
    got = ColorAug()(torch.from_numpy(np.stack(imgs)).cuda(), params).cpu().numpy()
    assert np.array_equal(got, np.stack(want))
