"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU (numpy) restatement of the pixel side of the reference's training data path, SURVEY.md section 8(f) row 2:
`src/datasets/transforms_depth.py` hflip (:206-231), vflip (:234-263), resize (:316-372), crop (:59-60,197-199), ColorJitter
(:551-604), ToTensor / Normalize (:618-660) as `src/datasets/coco.py:74-103` composes them.  The reference runs these on PIL
images; the algorithms below are Pillow's (third-party dependency of the reference, not vendored; checked here against
Pillow 12.2 / torchvision 0.26 as installed in the build container, `tests/test_data_oracle_cpu.py`):

* BILINEAR resize of 8-bit images = `ImagingResample` (Pillow `src/libImaging/Resample.c`): separable, antialiased (the triangle
  filter is stretched by the down-scaling factor), horizontal pass then vertical pass through an 8-bit intermediate, coefficients
  normalised per output pixel and rounded to 22 fractional bits, accumulators start at 2^21, results clipped to [0, 255];
* NEAREST resize of the auxiliary maps = `ImagingScaleAffine` (`Geometry.c`): source index = (int)(0.5 * scale + k * scale) with
  the running sum accumulated in double, as Pillow does;
* `ImageEnhance.Brightness / Contrast / Color` = `ImagingBlend` (`Blend.c`) against black / the rounded mean of the L image / the L
  image, single-precision arithmetic, truncation (0 <= factor <= 1) or clip + truncation (factor > 1);
* L conversion = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 (`Convert.c`);
* hue = RGB -> HSV -> (H + uint8(factor * 255)) mod 256 -> RGB (`Convert.c` rgb2hsv / hsv2rgb, torchvision `_functional_pil.py`).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the triangle filter over the whole axis
    -> (xmin int32 [out], count int32 [out], kk int32 [out, ksize])"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
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
        w = np.zeros(xmax, np.float64)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0 else v
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, kk


def _resample_axis(img, out_size, axis):
    """one 8-bit pass of ImagingResample along `axis` of img [H, W, C] uint8"""
    in_size = img.shape[axis]
    xmin, cnt, kk = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt[xx]):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, oh, ow):
    """PIL `Image.resize((ow, oh), BILINEAR)` of an 8-bit image [H, W, C] (what torchvision F.resize does to the RGB image)"""
    out = img
    if ow != img.shape[1]:
        out = _resample_axis(out, ow, 1)     # horizontal first
    if oh != img.shape[0]:
        out = _resample_axis(out, oh, 0)
    return out


def nearest_index(in_size, out_size):
    """source index of every output index under PIL `Image.resize(..., NEAREST)` (ImagingScaleAffine: running double sum)"""
    a = in_size / out_size
    idx = np.empty(out_size, np.int32)
    xo = 0.0 + a * 0.5
    for x in range(out_size):
        idx[x] = -1 if xo < 0.0 else int(xo)
        xo += a
    return np.minimum(idx, in_size - 1)


def resize_nearest(mat, oh, ow):
    """PIL NEAREST resize of an auxiliary map [H, W] (depth / segmentation, transforms_depth.py:368-370)"""
    iy, ix = nearest_index(mat.shape[0], oh), nearest_index(mat.shape[1], ow)
    return mat[iy][:, ix]


def to_gray(img):
    """PIL convert("L") of an RGB uint8 image [H, W, 3] -> uint8 [H, W]"""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(deg, img, factor):
    """PIL Image.blend(deg, img, factor) on uint8 arrays (single-precision arithmetic, Blend.c)"""
    a = np.float32(factor)
    d, i = deg.astype(np.int32), img.astype(np.int32)
    t = d.astype(np.float32) + a * (i - d).astype(np.float32)
    if 0.0 <= factor <= 1.0:
        return t.astype(np.int32).astype(np.uint8)                  # (UINT8) truncation
    return np.where(t <= 0.0, 0, np.where(t >= 255.0, 255, t.astype(np.int32))).astype(np.uint8)


def adjust_brightness(img, factor):
    return blend(np.zeros_like(img), img, factor)


def contrast_mean(img):
    """the grey level ImageEnhance.Contrast blends with: int(mean(L) + 0.5)"""
    g = to_gray(img)
    return int(g.astype(np.float64).sum() / g.size + 0.5)


def adjust_contrast(img, factor):
    return blend(np.full_like(img, contrast_mean(img)), img, factor)


def adjust_saturation(img, factor):
    g = to_gray(img)
    return blend(np.repeat(g[..., None], 3, axis=2), img, factor)


def rgb_to_hsv(img):
    """Pillow Convert.c rgb2hsv_row on uint8 [H, W, 3] -> uint8 HSV"""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    same = maxc == minc
    cr = np.where(same, 1, maxc - minc).astype(np.float32)
    s = cr / np.where(same, 1, maxc).astype(np.float32)
    rc = (maxc - r).astype(np.float32) / cr
    gc = (maxc - g).astype(np.float32) / cr
    bc = (maxc - b).astype(np.float32) / cr
    # `h = 2.0 + rc - bc` is evaluated in double (the literal is a double) and stored to a float
    f64 = np.float64
    h = np.where(r == maxc, (bc - gc).astype(f64),
                 np.where(g == maxc, 2.0 + rc.astype(f64) - bc.astype(f64), 4.0 + gc.astype(f64) - rc.astype(f64))).astype(np.float32)
    h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
    uh = np.clip((h.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    us = np.clip((s.astype(np.float64) * 255.0).astype(np.int32), 0, 255)
    uh, us = np.where(same, 0, uh), np.where(same, 0, us)
    return np.stack([uh, us, maxc], axis=-1).astype(np.uint8)


def hsv_to_rgb(hsv):
    """Pillow Convert.c hsv2rgb on uint8 HSV [H, W, 3] -> uint8 RGB"""
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    fh = h.astype(np.float32) * np.float32(6.0) / np.float32(255.0)
    i = np.floor(fh).astype(np.int32)
    f = fh - i.astype(np.float32)
    fs = s.astype(np.float32) / np.float32(255.0)
    vf = v.astype(np.float32)

    def rnd(x):     # C round(): half away from zero (single-precision product, exhaustively equal to Pillow over all 2^24 triples)
        return np.clip(np.floor(x.astype(np.float32).astype(np.float64) + 0.5).astype(np.int32), 0, 255)
    p = rnd(vf * (np.float32(1.0) - fs))
    q = rnd(vf * (np.float32(1.0) - fs * f))
    t = rnd(vf * (np.float32(1.0) - fs * (np.float32(1.0) - f)))
    i6 = i % 6
    r = np.choose(i6, [v, q, p, p, t, v])
    g = np.choose(i6, [t, v, v, q, p, p])
    b = np.choose(i6, [p, p, t, v, v, q])
    grey = s == 0
    out = np.stack([np.where(grey, v, r), np.where(grey, v, g), np.where(grey, v, b)], axis=-1)
    return out.astype(np.uint8)


def adjust_hue(img, factor):
    """torchvision F.adjust_hue on a PIL RGB image (_functional_pil.py): H + uint8(factor * 255) with uint8 wrap-around"""
    hsv = rgb_to_hsv(img)
    hsv[..., 0] = (hsv[..., 0].astype(np.int32) + int(np.uint8(np.int64(factor * 255) & 0xFF))) & 0xFF
    return hsv_to_rgb(hsv)


def normalize(img, mean, std):
    """ToTensor + Normalize (transforms_depth.py:618-633): uint8 [H, W, 3] -> float32 [3, H, W]"""
    x = img.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return (x - m) / s


def resize_target_size(w, h, size, max_size=None):
    """get_size_with_aspect_ratio of transforms_depth.py:319-339 -> (oh, ow)"""
    if max_size is not None:
        mn, mx = float(min(w, h)), float(max(w, h))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return h, w
    if w < h:
        return int(size * h / w), size
    return size, int(size * w / h)
