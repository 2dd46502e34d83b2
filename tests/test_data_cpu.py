"""Data path, CPU side: the numpy oracle (oracle/data_oracle.py) against Pillow / torchvision themselves -- the owners of the
algorithms the reference's transforms call (src/datasets/transforms_depth.py) -- and the host-side index tables of the C-ABI library
(gwd_pil_bilinear_coeffs, gwd_pil_nearest_index) against the oracle.  No GPU."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import data_oracle as D  # noqa: E402

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    smooth = np.asarray(Image.fromarray(coarse).resize((w, h), Image.BICUBIC)).astype(int)
    return np.clip(smooth + rng.integers(-25, 25, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("h,w,oh,ow", [(48, 64, 30, 40), (48, 64, 96, 128), (37, 53, 61, 29), (120, 160, 72, 97), (31, 47, 31, 90),
                                       (50, 70, 17, 70), (64, 64, 13, 7), (480, 640, 512, 682)])
def test_resize_bit_exact(h, w, oh, ow):
    img = _image(h, w, h + w)
    assert np.array_equal(D.resize_bilinear_u8(img, oh, ow), np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
    rng = np.random.default_rng(1)
    depth = rng.integers(0, 60000, (h, w)).astype(np.int32)
    assert np.array_equal(D.resize_nearest(depth, oh, ow), np.asarray(Image.fromarray(depth, mode="I").resize((ow, oh), Image.NEAREST)))
    seg = rng.integers(0, 3, (h, w)).astype(np.uint8)
    assert np.array_equal(D.resize_nearest(seg, oh, ow), np.asarray(Image.fromarray(seg, mode="L").resize((ow, oh), Image.NEAREST)))


def test_colour_conversions_exhaustive():
    """grey, RGB -> HSV and HSV -> RGB over all 2^24 triples"""
    a, b, c = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    cube = np.stack([a, b, c], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(D.to_gray(cube), np.asarray(Image.fromarray(cube).convert("L")))
    assert np.array_equal(D.rgb_to_hsv(cube), np.asarray(Image.fromarray(cube).convert("HSV")))
    assert np.array_equal(D.hsv_to_rgb(cube), np.asarray(Image.fromarray(cube, mode="HSV").convert("RGB")))


def test_colour_jitter_ops_bit_exact():
    F = pytest.importorskip("torchvision.transforms.functional")
    img = _image(60, 80, 3)
    pil = Image.fromarray(img)
    for f in (0.0, 0.6, 0.83, 1.0, 1.27, 1.4):
        assert np.array_equal(D.adjust_brightness(img, f), np.asarray(F.adjust_brightness(pil, f)))
        assert np.array_equal(D.adjust_contrast(img, f), np.asarray(F.adjust_contrast(pil, f)))
        assert np.array_equal(D.adjust_saturation(img, f), np.asarray(F.adjust_saturation(pil, f)))
    for f in (-0.5, -0.4, -0.13, 0.0, 0.2, 0.37, 0.5):
        assert np.array_equal(D.adjust_hue(img, f), np.asarray(F.adjust_hue(pil, f)))
    x = D.normalize(img, [0.538, 0.494, 0.453], [0.257, 0.263, 0.273])
    ref = F.normalize(F.to_tensor(pil), [0.538, 0.494, 0.453], [0.257, 0.263, 0.273]).numpy()
    assert np.array_equal(x, ref)


def test_host_tables_match_oracle():
    """the library's HOST functions (no kernel launch): coefficient tables and nearest indices"""
    so = os.path.join(ROOT, "gw-depth_b200", "libgwd_b200.so")
    if not os.path.exists(so):
        pytest.skip("library not built")
    lib = ctypes.CDLL(so)
    for fn in ("gwd_pil_bilinear_ksize", "gwd_pil_bilinear_coeffs", "gwd_pil_nearest_index"):
        getattr(lib, fn).restype = ctypes.c_int
    for n_in, n_out in [(640, 512), (480, 800), (53, 29), (31, 31), (600, 7), (7, 600), (1024, 1023)]:
        xmin, cnt, kk = D.bilinear_coeffs(n_in, n_out)
        ks = lib.gwd_pil_bilinear_ksize(n_in, n_out)
        assert ks == kk.shape[1]
        a, b = np.zeros(n_out, np.int32), np.zeros(n_out, np.int32)
        c = np.zeros((n_out, ks), np.int32)
        p = lambda t: t.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        assert lib.gwd_pil_bilinear_coeffs(n_in, n_out, p(a), p(b), p(c)) == 0
        assert np.array_equal(a, xmin) and np.array_equal(b, cnt) and np.array_equal(c, kk)
        idx = np.zeros(n_out, np.int32)
        assert lib.gwd_pil_nearest_index(n_in, n_out, p(idx)) == 0
        assert np.array_equal(idx, D.nearest_index(n_in, n_out))


def test_resize_target_size_rule():
    """get_size_with_aspect_ratio (transforms_depth.py:319-339): the oracle's restatement against torchvision-free arithmetic"""
    for (w, h, size, mx) in [(640, 480, 800, 1024), (640, 480, 480, 1024), (480, 640, 512, 1024), (1000, 300, 800, 1024), (333, 500, 600, None)]:
        oh, ow = D.resize_target_size(w, h, size, mx)
        assert min(oh, ow) <= size and (mx is None or max(oh, ow) <= mx + 1)
        if mx is None or max(w, h) / min(w, h) * size <= mx:
            assert min(oh, ow) == size


def _reference_transforms():
    import types
    import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference not available")
    pytest.importorskip("torchvision")
    ref_shims.install()
    src = os.path.join(ref_shims.REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    try:
        import shapely.geometry  # noqa: F401
    except ImportError:
        def _no(*a, **k):
            raise ImportError("shapely is not installed")
        sh, g = types.ModuleType("shapely"), types.ModuleType("shapely.geometry")
        g.Polygon, g.mapping, sh.geometry = _no, _no, g
        sys.modules["shapely"], sys.modules["shapely.geometry"] = sh, g
    import datasets.transforms_depth as T
    return T


def _targets(seed, h=480, w=640):
    import torch
    rng = np.random.default_rng(seed)
    lines, centers, ids = [], [], []
    for pid in range(4):
        cx, cy = rng.uniform(0.1 * w, 0.9 * w), rng.uniform(0.1 * h, 0.9 * h)
        rx, ry = rng.uniform(20, 0.3 * w), rng.uniform(20, 0.3 * h)
        pts = [(cx - rx, cy - ry), (cx + rx, cy - ry * 0.8), (cx + rx * 0.9, cy + ry), (cx - rx * 0.7, cy + ry * 0.9), (cx - rx * 1.1, cy)]
        for k in range(5):
            (x0, y0), (x1, y1) = pts[k], pts[(k + 1) % 5]
            if x0 > x1:
                x0, y0, x1, y1 = x1, y1, x0, y0
            lines.append([x0, y0, x1, y1])
            centers.append([cx, cy])
            ids.append(pid)
    n = len(lines)
    return {"lines": torch.tensor(lines, dtype=torch.float32), "poly_centers": torch.tensor(centers, dtype=torch.float32),
            "poly_ids": torch.tensor(ids), "labels": torch.zeros(n, dtype=torch.int64), "area": torch.ones(n), "iscrowd": torch.zeros(n),
            "size": torch.as_tensor([h, w])}


def test_crop_targets_and_size_rule_equal_the_reference():
    """the host side of gw-depth_b200/data.py that needs no GPU: the line clipping / polygon-centre logic of crop() (vectorised here,
    a per-line Python chain in src/datasets/transforms_depth.py:59-202) and the output size of resize (:319-339), against the
    reference functions on PIL images"""
    import importlib
    import torch
    sys.path.insert(0, ROOT)
    T = _reference_transforms()
    data = importlib.import_module("gw-depth_b200.data")
    rng = np.random.default_rng(0)
    img = np.zeros((480, 640, 3), np.uint8)
    checked = 0
    for seed in range(160):
        tgt = _targets(seed)
        w = int(rng.integers(200, 640)); h = int(rng.integers(150, 480))
        i = int(rng.integers(0, 480 - h + 1)); j = int(rng.integers(0, 640 - w + 1))
        try:
            _, ref_t = T.crop(Image.fromarray(img), {k: v.clone() for k, v in tgt.items()}, (i, j, h, w))
        except ImportError:
            continue          # the polygon / window intersection branch (shapely)
        cimg, our_t, _ = data.crop(torch.from_numpy(img), {k: v.clone() for k, v in tgt.items()}, (i, j, h, w))
        assert tuple(cimg.shape) == (h, w, 3)
        assert set(our_t) == set(ref_t)
        for k in ref_t:
            assert torch.equal(torch.as_tensor(our_t[k]), torch.as_tensor(ref_t[k])), (seed, k)
        checked += 1
    assert checked >= 15          # (the other crops hit the shapely branch, which the image does not have)
    for (w, h, size, mx) in [(640, 480, 800, 1024), (640, 480, 480, 1024), (480, 640, 512, 1024), (1000, 300, 800, 1024), (333, 500, 600, None),
                             (500, 375, 608, 1024), (1024, 1024, 800, 1024), (641, 480, 480, None)]:
        ref_img, _ = T.resize(Image.fromarray(np.zeros((h, w, 3), np.uint8)), None, size, mx)
        assert data._short_side_target(w, h, size, mx) == (ref_img.size[1], ref_img.size[0]) == D.resize_target_size(w, h, size, mx)
