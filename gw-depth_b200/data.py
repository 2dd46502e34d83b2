"""GPU-side training data path: the reference's joint image / line / auxiliary-map augmentation (`src/datasets/transforms_depth.py`,
composed by `src/datasets/coco.py:74-103`, applied by `src/datasets/glassrgbd_norhint.py:269-271`) on decoded uint8 images that already
sit on the device, plus the batch builder.  SURVEY.md section 8(f) row 2.

Same class names, constructor arguments, call signature `(image, target, aux_mats)` and -- call for call -- the same draws from Python's
`random` and torch's CPU generator as the reference, so a seeded run makes the same decisions and (the pixel kernels being bit-exact
restatements of Pillow's, `csrc/gwd_data.cu`) produces the same tensors.  Differences of representation only:

* `image` is a `torch.uint8` CUDA tensor `[H, W, 3]` (RGB, what `np.asarray(PIL image)` holds) instead of a PIL image; after
  `ToTensor` / `Normalize` it is the fp32 `[3, H, W]` tensor of the reference;
* `aux_mats` are CUDA tensors `[H, W]` (depth in mm as stored, segmentation ids); after `ToTensor` they are `[1, H, W]` as in the
  reference;
* targets stay small CPU tensors, transformed with the reference's own arithmetic.

`crop` needs `shapely` for one rare branch (polygon / crop-window intersection when a polygon keeps <= 3 lines), exactly like the
reference; it is imported only there.
"""
import random

import torch

from . import ops

IMAGE_MEAN = [0.538, 0.494, 0.453]     # src/datasets/coco.py:78
IMAGE_STD = [0.257, 0.263, 0.273]


def _size(image):
    """PIL's image.size = (w, h)"""
    if image.dim() == 3 and image.dtype == torch.uint8:
        return image.shape[1], image.shape[0]
    return image.shape[-1], image.shape[-2]


# ------------------------------------------------------------------------------------------------ functional forms
def _map_points(target, scale=None, shift=None, divide=None, swap_ends=False):
    """apply (x, y) -> (x, y) * scale + shift (or / divide) to everything in the target that lives in pixel coordinates: the line end
    points [n,4] (optionally exchanging the two ends), the polygon centres and the reflection points [n,2].  The reference only
    touches the centres when the target has lines (transforms_depth.py:213-224); so does this."""
    target = target.copy()

    def go(pts, reps):
        if scale is not None:
            pts = pts * torch.as_tensor(list(scale) * reps)
        if shift is not None:
            pts = pts + torch.as_tensor(list(shift) * reps)
        if divide is not None:
            pts = pts / torch.tensor(list(divide) * reps, dtype=torch.float32)
        return pts
    has_lines = "lines" in target
    if has_lines:
        lines = target["lines"]
        target["lines"] = go(lines[:, [2, 3, 0, 1]] if swap_ends else lines, 2)
        if "poly_centers" in target:
            target["poly_centers"] = go(target["poly_centers"], 1)
    return target, go, has_lines


def hflip(image, target, aux_mats=None):
    """mirror left-right (transforms_depth.py:206-231): x -> w - x, and the two ends of every line trade places so that the first end
    stays the left one"""
    w, h = _size(image)
    flipped = ops.resize_bilinear_u8(image, h, w, hflip=True)
    target, go, has_lines = _map_points(target, scale=(-1, 1), shift=(w, 0), swap_ends=True)
    if has_lines and "reflection_points" in target:
        target["reflection_points"] = go(target["reflection_points"], 1)
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, hflip=True) for m in aux_mats]
    return flipped, target, aux_mats


def vflip(image, target, aux_mats=None):
    """mirror top-bottom (transforms_depth.py:234-263): y -> h - y; a vertical line then has its lower end first, so its ends are
    exchanged (the data set keeps the upper point first when both x are equal)"""
    w, h = _size(image)
    flipped = ops.resize_bilinear_u8(image, h, w, vflip=True)
    target, go, has_lines = _map_points(target, scale=(1, -1), shift=(0, h))
    if has_lines:
        lines = target["lines"]
        upright = lines[:, 0] == lines[:, 2]
        lines[upright] = lines[upright][:, [2, 3, 0, 1]]
        if "reflection_points" in target:
            target["reflection_points"] = go(target["reflection_points"], 1)
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, vflip=True) for m in aux_mats]
    return flipped, target, aux_mats


def _short_side_target(w, h, size, max_size=None):
    """output (height, width) of a resize that brings the SHORT side to `size` without letting the long side exceed max_size
    (the rule of transforms_depth.py:319-339, truncating divisions included)"""
    short, long_ = (w, h) if w <= h else (h, w)
    if max_size is not None and float(long_) / float(short) * size > max_size:
        size = int(round(max_size * float(short) / float(long_)))
    if short == size:                       # already there: PIL would be asked for the same size
        return h, w
    return (int(size * h / w), size) if w < h else (size, int(size * w / h))


def resize(image, target, size, max_size=None, aux_mats=None):
    """transforms_depth.py:316-372 (size: the short side, or a (w, h) pair)"""
    w0, h0 = _size(image)
    if isinstance(size, (list, tuple)):
        oh, ow = size[::-1]
    else:
        oh, ow = _short_side_target(w0, h0, size, max_size)
    rescaled = ops.resize_bilinear_u8(image, oh, ow)
    if target is None:
        return rescaled, None
    rx, ry = float(ow) / float(w0), float(oh) / float(h0)
    target, go, _ = _map_points(target, scale=(rx, ry))
    if "reflection_points" in target:
        target["reflection_points"] = go(target["reflection_points"], 1)
    target["size"] = torch.tensor([oh, ow])
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, oh, ow) for m in aux_mats]
    return rescaled, target, aux_mats


def _mean_point(points):
    n = len(points)
    return (sum(pt[0] for pt in points) / n, sum(pt[1] for pt in points) / n)


def _polygon_window_overlap(window, polygon):
    """vertices (closed ring) of the intersection of the crop window with a glass polygon, [] when it is not one polygon
    (transforms_depth.py:32-43; shapely, imported here only, exactly like the reference)"""
    import numpy as np
    from shapely.geometry import Polygon, mapping
    common = Polygon(window).intersection(Polygon(polygon))
    desc = mapping(common)
    if desc["type"] != "Polygon" or np.array(desc["coordinates"]).size <= 2:
        return []
    return list(common.exterior.coords)


def _clip_lines(cl, w, h):
    """pull the end points of the (already shifted) lines [n,4] = (x1, y1, x2, y2) onto the crop window along the line: the eight
    sequential rules of transforms_depth.py:97-125, applied to all lines at once (same float32 operations in the same order; a rule
    only changes the rows whose condition holds)"""
    x1, y1, x2, y2 = (cl[:, k].clone() for k in range(4))
    slope = (y2 - y1) / (x2 - x1 + 1e-12)
    zero = torch.zeros_like(x1)
    # end 1 beyond the left / top, end 2 beyond the right / bottom, then the mirrored four cases
    c = x1 < 0
    x1 = torch.where(c, zero, x1)
    y1 = torch.where(c, y2 + (x1 - x2) * slope, y1)
    c = y1 < 0
    y1 = torch.where(c, zero, y1)
    x1 = torch.where(c, x2 - (y2 - y1) / slope, x1)
    c = x2 > w
    x2 = torch.where(c, zero + w, x2)
    y2 = torch.where(c, y1 + (x2 - x1) * slope, y2)
    c = y2 > h
    y2 = torch.where(c, zero + h, y2)
    x2 = torch.where(c, x1 + (y2 - y1) / slope, x2)
    c = x2 < 0
    x2 = torch.where(c, zero, x2)
    y2 = torch.where(c, y1 + (x2 - x1) * slope, y2)
    c = y2 < 0
    y2 = torch.where(c, zero, y2)
    x2 = torch.where(c, x1 - (y1 - y2) / slope, x2)
    c = x1 > w
    x1 = torch.where(c, zero + w, x1)
    y1 = torch.where(c, y2 + (x1 - x2) * slope, y1)
    c = y1 > h
    y1 = torch.where(c, zero + h, y1)
    x1 = torch.where(c, x2 + (y1 - y2) / slope, x1)
    out = torch.stack([x1, y1, x2, y2], dim=1)
    out[:, 0::2].clamp_(min=0, max=w)
    out[:, 1::2].clamp_(min=0, max=h)
    return out


def crop(image, target, region, aux_mats=None):
    """transforms_depth.py:59-202.  Pixels: a view of the device image / maps (no copy until the next kernel reads it).  Targets: lines
    with both ends beyond the same side of the window are dropped, the others are clipped along themselves; a polygon's centre is
    recomputed from what is left of it (or from its overlap with the window when three lines or fewer remain)."""
    top, left, h, w = region
    cropped = image[top:top + h, left:left + w]
    target = target.copy()
    target["size"] = torch.tensor([h, w])
    keep = None
    if "lines" in target:
        lines = target["lines"]
        shifted = lines - torch.as_tensor([left, top, left, top])
        xs, ys = shifted[:, 0::2], shifted[:, 1::2]
        gone = ((xs < 0).all(1) | (xs > w).all(1)) | ((ys < 0).all(1) | (ys > h).all(1))
        keep = ~gone
        target["lines"] = _clip_lines(shifted[keep], w, h)
        ids_before = target["poly_ids"]
        target["poly_ids"] = ids_before[keep]
        if "poly_centers" in target:
            right, bottom = left + w - 1, top + h - 1
            window = [[left, top], [left, bottom], [right, bottom], [right, top]]
            # after a horizontal flip every line runs right to left: the polygon's vertex chain is read from the other end
            mirrored = bool(lines[0, 0] == lines[1, 2] and lines[0, 1] == lines[1, 3])

            def vertex_chain(poly_lines):
                if mirrored:
                    poly_lines = poly_lines.reshape(-1, 2, 2).flip(1).reshape(-1, 4)
                return poly_lines[0].reshape(-1, 2).tolist() + poly_lines[1:, 2:].tolist()
            centers = torch.zeros_like(target["poly_centers"][keep])
            for pid in torch.unique(target["poly_ids"]):
                sel = target["poly_ids"] == pid
                left_over = target["lines"][sel]
                overlap = [] if len(left_over) > 3 else _polygon_window_overlap(window, vertex_chain(lines[ids_before == pid]))
                if overlap:
                    c = torch.tensor(_mean_point(overlap)) - torch.as_tensor([left, top])
                    c[0].clamp_(min=0, max=w)
                    c[1].clamp_(min=0, max=h)
                else:
                    c = torch.tensor(_mean_point(vertex_chain(left_over)))
                centers[sel, :] = c
            target["poly_centers"] = centers
    if "reflection_points" in target:
        pts = target["reflection_points"] - torch.as_tensor([left, top])
        inside = (pts[:, 0] >= 0) & (pts[:, 0] <= w) & (pts[:, 1] >= 0) & (pts[:, 1] <= h)
        target["reflection_points"] = pts[inside]
    if keep is not None:
        for field in ("labels", "area", "iscrowd"):
            if field in target:
                target[field] = target[field][keep]
    if aux_mats is not None:
        aux_mats = [m[top:top + h, left:left + w] for m in aux_mats]
    return cropped, target, aux_mats


def _random_crop_params(image, output_size):
    """torchvision T.RandomCrop.get_params (same torch.randint draws)"""
    w, h = _size(image)
    th, tw = output_size
    if h < th or w < tw:
        raise ValueError("Required crop size %s is larger than input image size %s" % ((th, tw), (h, w)))
    if w == tw and h == th:
        return 0, 0, h, w
    i = torch.randint(0, h - th + 1, size=(1,)).item()
    j = torch.randint(0, w - tw + 1, size=(1,)).item()
    return i, j, th, tw


# ------------------------------------------------------------------------------------------------ transform classes
class _Transform(object):
    """a callable (image, target, aux_mats) -> (image, target, aux_mats); subclasses name the reference's transforms"""

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%r" % kv for kv in sorted(vars(self).items())))


class _CoinFlip(_Transform):
    """applies `self.op` when one draw of random.random() falls below p (the flips of the reference)"""
    op = None

    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, img, target, aux_mats=None):
        hit = random.random() < self.p
        return type(self).op(img, target, aux_mats=aux_mats) if hit else (img, target, aux_mats)


class RandomHorizontalFlip(_CoinFlip):
    op = staticmethod(hflip)


class RandomVerticalFlip(_CoinFlip):
    op = staticmethod(vflip)


class RandomCrop(_Transform):
    def __init__(self, size):
        self.size = size

    def __call__(self, img, target, aux_mats=None):
        return crop(img, target, _random_crop_params(img, self.size), aux_mats=aux_mats)


class RandomSizeCrop(_Transform):
    """a window of random size within [min_size, max_size] (and the image) at a random place: width first, then height, then
    the position draws of RandomCrop -- the reference's order of RNG calls"""

    def __init__(self, min_size, max_size):
        self.min_size, self.max_size = min_size, max_size

    def __call__(self, img, target, aux_mats=None):
        iw, ih = _size(img)
        w = random.randint(self.min_size, min(iw, self.max_size))
        h = random.randint(self.min_size, min(ih, self.max_size))
        return crop(img, target, _random_crop_params(img, [h, w]), aux_mats=aux_mats)


class RandomResize(_Transform):
    def __init__(self, sizes, max_size=None):
        if not isinstance(sizes, (list, tuple)):
            raise TypeError("sizes: a list of short-side lengths")
        self.sizes, self.max_size = sizes, max_size

    def __call__(self, img, target=None, aux_mats=None):
        return resize(img, target, random.choice(self.sizes), self.max_size, aux_mats=aux_mats)


class Resize(_Transform):
    def __init__(self, size):
        self.size = size   # (w, h)

    def __call__(self, img, target=None, aux_mats=None):
        return resize(img, target, self.size, aux_mats=aux_mats)


class ColorJitter(_Transform):
    """transforms_depth.py:551-604: the four adjustments in a random order, each with a factor drawn from its range"""

    def __init__(self, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.4):
        inf = float("inf")
        self.brightness = self._range(brightness, "brightness", 1, 0, inf, True)
        self.contrast = self._range(contrast, "contrast", 1, 0, inf, True)
        self.saturation = self._range(saturation, "saturation", 1, 0, inf, True)
        self.hue = self._range(hue, "hue", 0, -0.5, 0.5, False)

    @staticmethod
    def _range(value, name, center, lowest, highest, floor_at_zero):
        """a number v means [center - v, center + v]; a pair is taken as is; None when the range is the identity"""
        if isinstance(value, (int, float)):
            if value < 0:
                raise ValueError("%s must be non negative" % name)
            lo, hi = center - float(value), center + float(value)
            if floor_at_zero:
                lo = max(lo, 0.0)
        elif isinstance(value, (tuple, list)) and len(value) == 2:
            lo, hi = value
            if not lowest <= lo <= hi <= highest:
                raise ValueError("%s values should be between %s and %s" % (name, lowest, highest))
        else:
            raise TypeError("%s should be a number or a pair" % name)
        return None if lo == hi == center else [lo, hi]

    def __call__(self, img, target, aux_mats=None):
        order, factors = [], []
        for fn_id in torch.randperm(4):
            rng = (self.brightness, self.contrast, self.saturation, self.hue)[int(fn_id)]
            if rng is not None:
                order.append(int(fn_id))
                factors.append(torch.tensor(1.0).uniform_(rng[0], rng[1]).item())
        if order:
            img = ops.color_jitter_u8(img.contiguous().clone() if not img.is_contiguous() else img.clone(), order, factors)
        if aux_mats is not None:
            return img, target, aux_mats
        return img, target


class RandomSelect(_Transform):
    """one draw of random.random(): the first pipeline with probability p, else the second"""

    def __init__(self, transforms1, transforms2, p=0.5):
        self.transforms1, self.transforms2, self.p = transforms1, transforms2, p

    def __call__(self, img, target, aux_mats=None):
        chosen = self.transforms1 if random.random() < self.p else self.transforms2
        return chosen(img, target, aux_mats=aux_mats)


class ToTensor(_Transform):
    """uint8 [H,W,3] -> uint8 [3,H,W] view marked for Normalize (the /255 happens there, in one kernel with the normalisation);
    auxiliary maps -> [1,H,W] (transforms_depth.py:618-628)"""

    def __call__(self, img, target, aux_mats=None):
        if aux_mats is None:
            return img, target
        return img, target, [m.contiguous()[None] for m in aux_mats]


class Normalize(_Transform):
    """ToTensor's /255 + Normalize of the image (one gwd_images_to_batch launch, bit-identical to torchvision) and the division of
    the targets by the image size (transforms_depth.py:631-659)"""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def __call__(self, image, target=None, aux_mats=None):
        w, h = _size(image)
        image = ops.images_to_batch([image.contiguous()], mean=self.mean, std=self.std, want_mask=False)[0][0]
        if target is None:
            return image, None
        target, go, _ = _map_points(target, divide=(w, h))
        if "reflection_points" in target:
            target["reflection_points"] = go(target["reflection_points"], 1)
        if aux_mats is not None:
            return image, target, aux_mats
        return image, target


class Compose(_Transform):
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, image, target, aux_mats):
        state = (image, target, aux_mats)
        for step in self.transforms:
            state = step(state[0], state[1], aux_mats=state[2])
        return state

    def __repr__(self):
        return "Compose(\n    " + "\n    ".join(repr(t) for t in self.transforms) + "\n)"


def make_coco_transforms(image_set, args=None, eval_mode=None):
    """src/datasets/coco.py:74-117"""
    normalize = Compose([ToTensor(), Normalize(IMAGE_MEAN, IMAGE_STD)])
    scales = [480, 512, 544, 576, 608, 640, 672, 680, 690, 704, 736, 768, 788, 800]
    test_size, max_size = 1024, 1024
    is_eval = bool(getattr(args, "eval", False)) if eval_mode is None else eval_mode
    if is_eval or image_set == "val":
        return Compose([RandomResize([test_size], max_size=max_size), normalize])
    if image_set == "train":
        return Compose([
            RandomSelect(RandomHorizontalFlip(), RandomVerticalFlip()),
            RandomSelect(RandomResize(scales, max_size=max_size),
                         Compose([RandomResize([400, 500, 600]), RandomSizeCrop(384, 600), RandomResize(scales, max_size=max_size)])),
            ColorJitter(),
            normalize,
        ])
    raise ValueError("unknown %s" % image_set)


def finish_sample(image, depth_gt, seg_gt, targets, with_center=True):
    """the tail of DataLoadPreprocess.__getitem__ (src/datasets/glassrgbd_norhint.py:273-297): depth mm -> m, glass classes -> one
    class, line + centre coordinates merged"""
    depth_gt = depth_gt / 1000.0
    seg_gt = torch.where(seg_gt > 0, 1, 0).type(torch.long)
    targets = dict(targets)
    if with_center:
        targets["lines"] = torch.cat([targets["lines"], targets["poly_centers"]], dim=1)
    for k in ("poly_centers", "area", "iscrowd"):
        targets.pop(k, None)
    return image, depth_gt, seg_gt, targets
