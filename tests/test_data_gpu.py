"""Data path on the GPU (gw-depth_b200/data.py + csrc/gwd_data.cu) against (a) the numpy oracle, kernel by kernel, bit for bit, and (b)
the reference's own transforms (src/datasets/transforms_depth.py, staged unmodified under baseline/_ref) run on PIL images with the same
seeds: same decisions, same pixels, same targets."""
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def _data():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import data
    return data


def _image(h, w, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,oh,ow", [(48, 64, 30, 40), (48, 64, 96, 128), (37, 53, 61, 29), (120, 160, 72, 97), (31, 47, 31, 90),
                                       (50, 70, 17, 70), (480, 640, 512, 682), (480, 640, 480, 640)])
@pytest.mark.parametrize("hf,vf", [(False, False), (True, False), (False, True)])
def test_resize_and_flip_bit_exact(h, w, oh, ow, hf, vf):
    import data_oracle as D
    ops = _ops()
    img = _image(h, w, h * 3 + w)
    src = img[:, ::-1] if hf else img
    src = src[::-1] if vf else src
    ref = D.resize_bilinear_u8(np.ascontiguousarray(src), oh, ow)
    got = ops.resize_bilinear_u8(torch.from_numpy(img).cuda(), oh, ow, hflip=hf, vflip=vf)
    assert np.array_equal(got.cpu().numpy(), ref)
    rng = np.random.default_rng(7)
    for mat in (rng.integers(0, 60000, (h, w)).astype(np.int32), rng.integers(0, 3, (h, w)).astype(np.uint8),
                rng.random((h, w)).astype(np.float32), rng.integers(0, 9, (h, w)).astype(np.int64)):
        s = mat[:, ::-1] if hf else mat
        s = s[::-1] if vf else s
        got = ops.gather2d(torch.from_numpy(mat).cuda(), oh, ow, hflip=hf, vflip=vf)
        assert np.array_equal(got.cpu().numpy(), D.resize_nearest(np.ascontiguousarray(s), oh, ow))


def test_resize_of_a_crop_view():
    import data_oracle as D
    ops = _ops()
    img = _image(90, 120, 5)
    t = torch.from_numpy(img).cuda()
    got = ops.resize_bilinear_u8(t[11:71, 23:103], 75, 100)
    assert np.array_equal(got.cpu().numpy(), D.resize_bilinear_u8(np.ascontiguousarray(img[11:71, 23:103]), 75, 100))
    dep = np.random.default_rng(0).integers(0, 5000, (90, 120)).astype(np.int32)
    got = ops.gather2d(torch.from_numpy(dep).cuda()[11:71, 23:103], 75, 100)
    assert np.array_equal(got.cpu().numpy(), D.resize_nearest(np.ascontiguousarray(dep[11:71, 23:103]), 75, 100))


def test_colour_jitter_bit_exact():
    """every op alone over the factor range (incl. the clip branch, factor > 1), then all 24 orders of the four ops"""
    import itertools
    import data_oracle as D
    ops = _ops()
    img = _image(96, 128, 11)
    img[:8] = np.random.default_rng(3).integers(0, 256, (8, 1, 1), dtype=np.uint8)     # grey rows: the s == 0 / max == min branches
    fns = {0: D.adjust_brightness, 1: D.adjust_contrast, 2: D.adjust_saturation, 3: D.adjust_hue}
    for op in range(4):
        for f in ((0.0, 0.31, 0.6, 1.0, 1.27, 1.4) if op < 3 else (-0.5, -0.4, -0.07, 0.0, 0.2, 0.4, 0.5)):
            f = float(np.float32(f))
            got = ops.color_jitter_u8(torch.from_numpy(img).cuda(), [op], [f])
            assert np.array_equal(got.cpu().numpy(), fns[op](img, f)), (op, f)
    fac = {0: float(np.float32(1.23)), 1: float(np.float32(0.71)), 2: float(np.float32(1.37)), 3: float(np.float32(-0.21))}
    for order in itertools.permutations(range(4)):
        ref = img
        for op in order:
            ref = fns[op](ref, fac[op])
        got = ops.color_jitter_u8(torch.from_numpy(img).cuda(), list(order), [fac[o] for o in order])
        assert np.array_equal(got.cpu().numpy(), ref), order


def _reference_transforms():
    """the staged, unmodified reference module (baseline/_ref/src/datasets/transforms_depth.py); shapely is absent from the image and
    only used by one rare branch of crop()"""
    import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference not staged (oracle/stage_ref.sh)")
    pytest.importorskip("PIL")
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
    import datasets.coco as C
    return T, C


def _sample(seed, h=480, w=640):
    """one synthetic GlassRGBD-shaped sample: RGB image, 16-bit depth in mm, glass mask, quadrilateral glass polygons as 4 lines each
    (x0 <= x1 per line, src/datasets/glassrgbd_norhint.py:182-211)"""
    rng = np.random.default_rng(seed)
    img = _image(h, w, seed)
    depth = rng.integers(300, 9000, (h, w)).astype(np.int32)
    seg = rng.integers(0, 3, (h, w)).astype(np.uint8)
    lines, centers, ids = [], [], []
    for pid in range(3):
        cx, cy = rng.uniform(0.25 * w, 0.75 * w), rng.uniform(0.25 * h, 0.75 * h)
        rx, ry = rng.uniform(30, 0.2 * w), rng.uniform(30, 0.2 * h)
        pts = [(cx - rx, cy - ry), (cx + rx, cy - ry * 0.8), (cx + rx * 0.9, cy + ry), (cx - rx * 0.7, cy + ry * 0.9)]
        for k in range(4):
            (x0, y0), (x1, y1) = pts[k], pts[(k + 1) % 4]
            if x0 > x1:
                x0, y0, x1, y1 = x1, y1, x0, y0
            lines.append([x0, y0, x1, y1])
            centers.append([cx, cy])
            ids.append(pid)
    n = len(lines)
    target = {"lines": torch.tensor(lines, dtype=torch.float32), "poly_centers": torch.tensor(centers, dtype=torch.float32),
              "poly_ids": torch.tensor(ids), "labels": torch.zeros(n, dtype=torch.int64), "area": torch.ones(n), "iscrowd": torch.zeros(n),
              "orig_size": torch.as_tensor([h, w]), "size": torch.as_tensor([h, w]), "image_id": torch.tensor([seed])}
    return img, depth, seg, target


def _same_target(a, b):
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(torch.as_tensor(a[k]), torch.as_tensor(b[k])), k


@pytest.mark.parametrize("image_set", ["train", "val"])
def test_pipeline_equals_reference_transforms(image_set):
    """make_coco_transforms(image_set) of the reference on PIL images vs data.make_coco_transforms on device tensors, same seeds: the
    normalised image, the depth / segmentation maps and every target field are EQUAL (train: flips, multi-scale resize, size crop,
    ColorJitter in all its random orders; val: the 1024 resize)"""
    from PIL import Image
    T, C = _reference_transforms()
    data = _data()
    args = types.SimpleNamespace(eval=False)
    ref_tf = C.make_coco_transforms(image_set, args)
    our_tf = data.make_coco_transforms(image_set, args)
    done = 0
    for seed in range(14 if image_set == "train" else 3):
        img, depth, seg, target = _sample(seed, *((480, 640) if seed % 3 else (375, 500)))
        random.seed(100 + seed)
        torch.manual_seed(100 + seed)
        try:
            r_img, r_tgt, (r_dep, r_seg) = ref_tf(Image.fromarray(img), {k: v.clone() for k, v in target.items()},
                                                  aux_mats=[Image.fromarray(depth, mode="I"), Image.fromarray(seg, mode="L")])
        except ImportError:
            continue        # the crop branch that needs shapely
        random.seed(100 + seed)
        torch.manual_seed(100 + seed)
        o_img, o_tgt, (o_dep, o_seg) = our_tf(torch.from_numpy(img).cuda(), {k: v.clone() for k, v in target.items()},
                                              aux_mats=[torch.from_numpy(depth).cuda(), torch.from_numpy(seg).cuda()])
        assert tuple(o_img.shape) == tuple(r_img.shape)
        assert torch.equal(o_img.cpu(), r_img), "seed %d: image differs (max %g)" % (seed, (o_img.cpu() - r_img).abs().max())
        assert torch.equal(o_dep.cpu(), r_dep.to(o_dep.dtype)) and torch.equal(o_seg.cpu(), r_seg.to(o_seg.dtype))
        _same_target(o_tgt, r_tgt)
        done += 1
    assert done >= (8 if image_set == "train" else 3)


def test_batch_from_augmented_samples():
    """augmented samples of different sizes -> the padded batch + mask of nested_tensor_from_tensor_list (src/util/misc.py:291-313)"""
    data, ops = _data(), _ops()
    tf = data.Compose([data.RandomResize([480, 512, 544], max_size=1024), data.ColorJitter()])
    random.seed(3)
    torch.manual_seed(3)
    imgs = []
    for seed in range(3):
        img, depth, seg, target = _sample(seed, 480 - 32 * seed, 640)
        o_img, _, _ = tf(torch.from_numpy(img).cuda(), target, aux_mats=[torch.from_numpy(depth).cuda(), torch.from_numpy(seg).cuda()])
        imgs.append(o_img)
    batch, mask, padded = ops.images_to_batch(imgs)
    H, W = max(t.shape[0] for t in imgs), max(t.shape[1] for t in imgs)
    assert tuple(batch.shape) == (3, 3, H, W) and padded
    for b, t in enumerate(imgs):
        h, w = t.shape[:2]
        ref = (t.cpu().permute(2, 0, 1).float().div(255) - torch.tensor(data.IMAGE_MEAN)[:, None, None]) / torch.tensor(data.IMAGE_STD)[:, None, None]
        assert torch.equal(batch[b, :, :h, :w].cpu(), ref)      # ToTensor + Normalize on the CPU (IEEE division)
        assert not mask[b, :h, :w].any() and mask[b, h:].all() and mask[b, :, w:].all()
