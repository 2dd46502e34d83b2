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
def hflip(image, target, aux_mats=None):
    """transforms_depth.py:206-231"""
    w, h = _size(image)
    flipped = ops.resize_bilinear_u8(image, h, w, hflip=True)
    target = target.copy()
    if "lines" in target:
        lines = target["lines"]
        target["lines"] = lines[:, [2, 3, 0, 1]] * torch.as_tensor([-1, 1, -1, 1]) + torch.as_tensor([w, 0, w, 0])
        if "poly_centers" in target:
            target["poly_centers"] = target["poly_centers"] * torch.as_tensor([-1, 1]) + torch.as_tensor([w, 0])
        if "reflection_points" in target:
            target["reflection_points"] = target["reflection_points"] * torch.as_tensor([-1, 1]) + torch.as_tensor([w, 0])
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, hflip=True) for m in aux_mats]
    return flipped, target, aux_mats


def vflip(image, target, aux_mats=None):
    """transforms_depth.py:234-263"""
    w, h = _size(image)
    flipped = ops.resize_bilinear_u8(image, h, w, vflip=True)
    target = target.copy()
    if "lines" in target:
        lines = target["lines"] * torch.as_tensor([1, -1, 1, -1]) + torch.as_tensor([0, h, 0, h])
        vertical = lines[:, 0] == lines[:, 2]
        lines[vertical] = torch.index_select(lines[vertical], 1, torch.tensor([2, 3, 0, 1]))
        target["lines"] = lines
        if "poly_centers" in target:
            target["poly_centers"] = target["poly_centers"] * torch.as_tensor([1, -1]) + torch.as_tensor([0, h])
        if "reflection_points" in target:
            target["reflection_points"] = target["reflection_points"] * torch.as_tensor([1, -1]) + torch.as_tensor([0, h])
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, vflip=True) for m in aux_mats]
    return flipped, target, aux_mats


def _get_size_with_aspect_ratio(image_size, size, max_size=None):
    w, h = image_size
    if max_size is not None:
        min_original_size = float(min((w, h)))
        max_original_size = float(max((w, h)))
        if max_original_size / min_original_size * size > max_size:
            size = int(round(max_size * min_original_size / max_original_size))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        ow = size
        oh = int(size * h / w)
    else:
        oh = size
        ow = int(size * w / h)
    return (oh, ow)


def resize(image, target, size, max_size=None, aux_mats=None):
    """transforms_depth.py:316-372 (size: the short side, or a (w, h) pair)"""
    w0, h0 = _size(image)
    if isinstance(size, (list, tuple)):
        oh, ow = size[::-1]
    else:
        oh, ow = _get_size_with_aspect_ratio((w0, h0), size, max_size)
    rescaled = ops.resize_bilinear_u8(image, oh, ow)
    if target is None:
        return rescaled, None
    ratio_width, ratio_height = float(ow) / float(w0), float(oh) / float(h0)
    target = target.copy()
    if "lines" in target:
        target["lines"] = target["lines"] * torch.as_tensor([ratio_width, ratio_height, ratio_width, ratio_height])
        if "poly_centers" in target:
            target["poly_centers"] = target["poly_centers"] * torch.as_tensor([ratio_width, ratio_height])
    if "reflection_points" in target:
        target["reflection_points"] = target["reflection_points"] * torch.as_tensor([ratio_width, ratio_height])
    target["size"] = torch.tensor([oh, ow])
    if aux_mats is not None:
        aux_mats = [ops.gather2d(m, oh, ow) for m in aux_mats]
    return rescaled, target, aux_mats


def _centroid(vertexes):
    xs = [v[0] for v in vertexes]
    ys = [v[1] for v in vertexes]
    return (sum(xs) / len(vertexes), sum(ys) / len(vertexes))


def _intersect_remap(main_coors, poly_coors):
    """transforms_depth.py:32-43 (shapely, as the reference)"""
    import numpy as np
    from shapely.geometry import Polygon, mapping
    inter = Polygon(main_coors).intersection(Polygon(poly_coors))
    m = mapping(inter)
    if m["type"] == "Polygon":
        if np.array(m["coordinates"]).size <= 2:
            return []
        return list(inter.exterior.coords)
    return []


def crop(image, target, region, aux_mats=None):
    """transforms_depth.py:59-202: pixels are a view (no copy until the next kernel reads it); the line clipping is the reference's"""
    i, j, h, w = region
    cropped = image[i:i + h, j:j + w]
    target = target.copy()
    x_lt, y_lt = j, i
    target["size"] = torch.tensor([h, w])
    fields = ["labels", "area", "iscrowd"]
    keep = None
    if "lines" in target:
        lines = target["lines"]
        cl = lines - torch.as_tensor([j, i, j, i])
        eps = 1e-12
        remove_x = torch.logical_or(torch.logical_and(cl[:, 0] < 0, cl[:, 2] < 0), torch.logical_and(cl[:, 0] > w, cl[:, 2] > w))
        remove_y = torch.logical_or(torch.logical_and(cl[:, 1] < 0, cl[:, 3] < 0), torch.logical_and(cl[:, 1] > h, cl[:, 3] > h))
        keep = torch.logical_and(~remove_x, ~remove_y)
        cl = cl[keep]
        clamped = torch.zeros_like(cl)
        for n, line in enumerate(cl):
            x1, y1, x2, y2 = line
            slope = (y2 - y1) / (x2 - x1 + eps)
            if x1 < 0:
                x1 = 0
                y1 = y2 + (x1 - x2) * slope
            if y1 < 0:
                y1 = 0
                x1 = x2 - (y2 - y1) / slope
            if x2 > w:
                x2 = w
                y2 = y1 + (x2 - x1) * slope
            if y2 > h:
                y2 = h
                x2 = x1 + (y2 - y1) / slope
            if x2 < 0:
                x2 = 0
                y2 = y1 + (x2 - x1) * slope
            if y2 < 0:
                y2 = 0
                x2 = x1 - (y1 - y2) / slope
            if x1 > w:
                x1 = w
                y1 = y2 + (x1 - x2) * slope
            if y1 > h:
                y1 = h
                x1 = x2 + (y1 - y2) / slope
            clamped[n, :] = torch.tensor([x1, y1, x2, y2])
        clamped[:, 0::2].clamp_(min=0, max=w)
        clamped[:, 1::2].clamp_(min=0, max=h)
        target["lines"] = clamped
        src_poly_ids = target["poly_ids"]
        target["poly_ids"] = target["poly_ids"][keep]
        if "poly_centers" in target:
            x_rb, y_rb = x_lt + w - 1, y_lt + h - 1
            crp_point = [[x_lt, y_lt], [x_lt, y_rb], [x_rb, y_rb], [x_rb, y_lt]]
            horiz_flipped = bool(lines[0, 0] == lines[1, 2] and lines[0, 1] == lines[1, 3])
            centers = torch.zeros_like(target["poly_centers"][keep])
            for py_id in torch.unique(target["poly_ids"]):
                py_index = target["poly_ids"] == py_id
                py_lines = target["lines"][py_index]

                def points_of(pl):
                    if horiz_flipped:
                        pl = pl.reshape(-1, 2, 2).flip(1).reshape(-1, 4)
                    return pl[0].reshape(-1, 2).tolist() + pl[1:, 2:].tolist()
                if len(py_lines) > 3:
                    centers[py_index, :] = torch.tensor(_centroid(points_of(py_lines)))
                else:
                    joint = _intersect_remap(crp_point, points_of(lines[src_poly_ids == py_id]))
                    if len(joint) > 0:
                        c = torch.tensor(_centroid(joint)) - torch.as_tensor([x_lt, y_lt])
                        c[0].clamp_(min=0, max=w)
                        c[1].clamp_(min=0, max=h)
                        centers[py_index, :] = c
                    else:
                        centers[py_index, :] = torch.tensor(_centroid(points_of(py_lines)))
            target["poly_centers"] = centers
    if "reflection_points" in target:
        pts = target["reflection_points"] - torch.as_tensor([j, i])
        remove = torch.logical_or(torch.logical_or(pts[:, 0] < 0, pts[:, 0] > w), torch.logical_or(pts[:, 1] < 0, pts[:, 1] > h))
        target["reflection_points"] = pts[~remove]
    if keep is not None:
        for field in fields:
            if field in target:
                target[field] = target[field][keep]
    if aux_mats is not None:
        aux_mats = [m[i:i + h, j:j + w] for m in aux_mats]
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
class RandomCrop(object):
    def __init__(self, size):
        self.size = size

    def __call__(self, img, target, aux_mats=None):
        return crop(img, target, _random_crop_params(img, self.size), aux_mats=aux_mats)


class RandomSizeCrop(object):
    def __init__(self, min_size, max_size):
        self.min_size, self.max_size = min_size, max_size

    def __call__(self, img, target, aux_mats=None):
        iw, ih = _size(img)
        w = random.randint(self.min_size, min(iw, self.max_size))
        h = random.randint(self.min_size, min(ih, self.max_size))
        return crop(img, target, _random_crop_params(img, [h, w]), aux_mats=aux_mats)


class RandomHorizontalFlip(object):
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, img, target, aux_mats=None):
        if random.random() < self.p:
            return hflip(img, target, aux_mats=aux_mats)
        return img, target, aux_mats


class RandomVerticalFlip(object):
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, img, target, aux_mats=None):
        if random.random() < self.p:
            return vflip(img, target, aux_mats=aux_mats)
        return img, target, aux_mats


class RandomResize(object):
    def __init__(self, sizes, max_size=None):
        assert isinstance(sizes, (list, tuple))
        self.sizes, self.max_size = sizes, max_size

    def __call__(self, img, target=None, aux_mats=None):
        return resize(img, target, random.choice(self.sizes), self.max_size, aux_mats=aux_mats)


class Resize(object):
    def __init__(self, size):
        self.size = size   # (w, h)

    def __call__(self, img, target=None, aux_mats=None):
        return resize(img, target, self.size, aux_mats=aux_mats)


class ColorJitter(object):
    """transforms_depth.py:551-604: the four adjustments in a random order, each with a factor drawn from its range"""

    def __init__(self, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.4):
        self.brightness = self._check_input(brightness, "brightness")
        self.contrast = self._check_input(contrast, "contrast")
        self.saturation = self._check_input(saturation, "saturation")
        self.hue = self._check_input(hue, "hue", center=0, bound=(-0.5, 0.5), clip_first_on_zero=False)

    @staticmethod
    def _check_input(value, name, center=1, bound=(0, float("inf")), clip_first_on_zero=True):
        if isinstance(value, (int, float)):
            if value < 0:
                raise ValueError("If {} is a single number, it must be non negative.".format(name))
            value = [center - float(value), center + float(value)]
            if clip_first_on_zero:
                value[0] = max(value[0], 0.0)
        elif isinstance(value, (tuple, list)) and len(value) == 2:
            if not bound[0] <= value[0] <= value[1] <= bound[1]:
                raise ValueError("{} values should be between {}".format(name, bound))
        else:
            raise TypeError("{} should be a single number or a list/tuple with lenght 2.".format(name))
        if value[0] == value[1] == center:
            value = None
        return value

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


class RandomSelect(object):
    def __init__(self, transforms1, transforms2, p=0.5):
        self.transforms1, self.transforms2, self.p = transforms1, transforms2, p

    def __call__(self, img, target, aux_mats=None):
        if random.random() < self.p:
            return self.transforms1(img, target, aux_mats=aux_mats)
        return self.transforms2(img, target, aux_mats=aux_mats)


class ToTensor(object):
    """uint8 [H,W,3] -> uint8 [3,H,W] view marked for Normalize (the /255 happens there, in one kernel with the normalisation);
    auxiliary maps -> [1,H,W] (transforms_depth.py:618-628)"""

    def __call__(self, img, target, aux_mats=None):
        if aux_mats is None:
            return img, target
        return img, target, [m.contiguous()[None] for m in aux_mats]


class Normalize(object):
    """ToTensor's /255 + Normalize of the image (one gwd_images_to_batch launch, bit-identical to torchvision) and the division of
    the targets by the image size (transforms_depth.py:631-659)"""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def __call__(self, image, target=None, aux_mats=None):
        w, h = _size(image)
        image = ops.images_to_batch([image.contiguous()], mean=self.mean, std=self.std, want_mask=False)[0][0]
        if target is None:
            return image, None
        target = target.copy()
        if "lines" in target:
            target["lines"] = target["lines"] / torch.tensor([w, h, w, h], dtype=torch.float32)
            if "poly_centers" in target:
                target["poly_centers"] = target["poly_centers"] / torch.tensor([w, h], dtype=torch.float32)
        if "reflection_points" in target:
            target["reflection_points"] = target["reflection_points"] / torch.tensor([w, h], dtype=torch.float32)
        if aux_mats is not None:
            return image, target, aux_mats
        return image, target


class Compose(object):
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, image, target, aux_mats):
        for t in self.transforms:
            image, target, aux_mats = t(image, target, aux_mats=aux_mats)
        return image, target, aux_mats

    def __repr__(self):
        return self.__class__.__name__ + "(" + "".join("\n    {0}".format(t) for t in self.transforms) + "\n)"


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
