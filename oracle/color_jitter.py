"""CPU restatement of the reference's colour augmentation - TEST INFRASTRUCTURE, never imported by the product.

The reference augments PIL images with torchvision's ColorJitter (/root/reference/datasets/mono_dataset.py:60-70,
136,169-176: ``ColorJitter.get_params(brightness, contrast, saturation, hue)`` applied to every frame of an item,
then ``to_tensor``), i.e. torchvision.transforms.functional.adjust_{brightness,contrast,saturation,hue} on uint8 RGB
images in a random order.  Those are Pillow operations (third-party, not vendored in /root/reference; installed here:
Pillow 12.2.0, torchvision 0.26.0):
  * ImageEnhance.{Brightness,Contrast,Color}: ``Image.blend(degenerate, image, factor)`` - libImaging/Blend.c: per byte
    ``in1 + alpha * (in2 - in1)`` in C float, truncated to uint8 (clipped first when alpha is outside [0, 1]); the
    degenerate image is black / the rounded mean of the "L" image / the "L" image;
  * "L" conversion: ``(R * 19595 + G * 38470 + B * 7471 + 0x8000) >> 16`` (libImaging/Convert.c, ITU-R 601-2);
  * hue: RGB -> HSV (libImaging/Convert.c rgb2hsv_row, following colorsys.py, float with double intermediates),
    ``h += uint8(int32(hue_factor * 255))`` with wrap-around (torchvision _functional_pil.adjust_hue), HSV -> RGB.
Pinned: tests/test_color_jitter.py compares every function with the installed Pillow / torchvision - the two HSV
conversions exhaustively over all 2^24 colours - so the CUDA kernels (md2_color_jitter_u8) can be held to this file
byte for byte on the GPU box, where only the wheels are present.
"""
import numpy as np

f32, f64 = np.float32, np.float64


def to_l(rgb):
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(in1, in2, alpha):
    """libImaging/Blend.c ImagingBlend, alpha as C float."""
    a = f32(alpha)
    i1 = in1.astype(np.int32)
    d = (in2.astype(np.int32) - i1).astype(f32)
    t = (i1.astype(f32) + (a * d).astype(f32)).astype(f32)          # float product, then float sum
    if 0.0 <= float(a) <= 1.0:
        return t.astype(np.int32).astype(np.uint8)                   # (UINT8) cast: truncation, value always in range
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(rgb, factor):
    return blend(np.zeros_like(rgb), rgb, factor)


def contrast_mean(rgb):
    """int(ImageStat.Stat(image.convert("L")).mean[0] + 0.5): one gray level per image."""
    l = to_l(rgb)
    return int(float(int(l.astype(np.int64).sum())) / float(l.size) + 0.5)


def adjust_contrast(rgb, factor):
    return blend(np.full_like(rgb, contrast_mean(rgb)), rgb, factor)


def adjust_saturation(rgb, factor):
    l = to_l(rgb)
    return blend(np.stack([l, l, l], -1), rgb, factor)


def rgb2hsv(rgb):
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    with np.errstate(all="ignore"):
        cr = (maxc - minc).astype(f32)
        s = (cr / maxc.astype(f32)).astype(f32)
        rc = ((maxc - r).astype(f32) / cr).astype(f32)
        gc = ((maxc - g).astype(f32) / cr).astype(f32)
        bc = ((maxc - b).astype(f32) / cr).astype(f32)
        h = np.where(r == maxc, (bc - gc).astype(f32),
                     np.where(g == maxc, (f64(2.0) + rc.astype(f64) - bc.astype(f64)).astype(f32),
                              (f64(4.0) + gc.astype(f64) - rc.astype(f64)).astype(f32)))
        h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
        uh = np.clip((h.astype(f64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((s.astype(f64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def _c_round(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def hsv2rgb(hsv):
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    hf = h.astype(f32).astype(f64) * 6.0 / 255.0
    i = np.floor(hf).astype(np.int64)
    f = (hf - i.astype(f32).astype(f64)).astype(f32).astype(f64)
    fs = (s.astype(f32).astype(f64) / 255.0).astype(f32).astype(f64)
    vf = v.astype(f32).astype(f64)
    up = np.clip(_c_round(vf * (1.0 - fs)), 0, 255).astype(np.uint8)
    uq = np.clip(_c_round(vf * (1.0 - fs * f)), 0, 255).astype(np.uint8)
    ut = np.clip(_c_round(vf * (1.0 - fs * (1.0 - f))), 0, 255).astype(np.uint8)
    k = i % 6
    out = np.stack([np.choose(k, [v, uq, up, up, ut, v]), np.choose(k, [ut, v, v, uq, up, up]),
                    np.choose(k, [up, up, ut, v, v, uq])], -1)
    return np.where((s == 0)[..., None], np.stack([v, v, v], -1), out).astype(np.uint8)


def hue_shift(hue_factor):
    """torchvision _functional_pil.adjust_hue: np.int32(hue_factor * 255).astype(np.uint8)"""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


def adjust_hue(rgb, hue_factor):
    hsv = rgb2hsv(rgb)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_shift(hue_factor)).astype(np.uint8)     # wraps modulo 256
    return hsv2rgb(hsv)


def color_jitter(rgb, fn_idx, brightness, contrast, saturation, hue):
    """torchvision ColorJitter.forward on one uint8 (H,W,3) image: the four adjustments in the order fn_idx gives
    (0 brightness, 1 contrast, 2 saturation, 3 hue); a factor that is None is skipped."""
    ops = {0: (adjust_brightness, brightness), 1: (adjust_contrast, contrast), 2: (adjust_saturation, saturation),
           3: (adjust_hue, hue)}
    out = rgb
    for i in fn_idx:
        fn, fac = ops[int(i)]
        if fac is not None:
            out = fn(out, fac)
    return out
