"""CPU restatement of Pillow's uint8 LANCZOS resize (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline leg may import this).

The reference builds its colour pyramid on the host with torchvision `transforms.Resize(..., interpolation=
Image.ANTIALIAS)` on PIL images (/root/reference/datasets/mono_dataset.py:57,82-86,98-103), i.e.
`PIL.Image.resize(size, LANCZOS)`: scale 0 from the native frame, scale i from scale i-1.  Pillow is a third-party
dependency that is not vendored in /root/reference (installed here: Pillow 12.2.0); its algorithm
(src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
ImagingResampleVertical_8bpc) is restated below and pinned against the installed Pillow by
tests/test_pyramid.py (byte-exact)."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box.
    Returns (ksize, bounds[out_size, 2] (xmin, count), kk[out_size, ksize] int32)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, out_size, axis):
    """One separable pass over `axis` of a (H, W, C) uint8 image, 8bpc fixed point as in Pillow."""
    in_size = img.shape[axis]
    _, bounds, kk = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_lanczos_u8(img, out_h, out_w):
    """img: (H, W, C) uint8 -> (out_h, out_w, C) uint8, = PIL.Image.resize((out_w, out_h), LANCZOS):
    horizontal pass first (into a uint8 temporary), then vertical (Resample.c ImagingResampleInner)."""
    img = np.asarray(img, np.uint8)
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return img


def build_pyramid_u8(native, height, width, num_scales=4):
    """MonoDataset.preprocess (mono_dataset.py:98-103): scale 0 from the native frame, scale i from scale i-1."""
    levels = []
    cur = np.asarray(native, np.uint8)
    for i in range(num_scales):
        cur = resize_lanczos_u8(cur, height >> i, width >> i)
        levels.append(cur)
    return levels
