"""The whole training stack on one GPU, end to end: decoded uint8 samples -> the reference's training augmentation on the device
(gw-depth_b200/data.py) -> a ragged batch padded to a multiple of 32 with its mask -> Trainer.train_step (forward, 17 losses, Hungarian
matching, backward, clip, AdamW).  The multi-scale resize makes the images 640..1066 pixels wide, i.e. up to 850 tokens at 1/32 scale:
the key-tiled attention forward and the long-sequence attention backward run with key-padding masks."""
import os
import random
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_weights  # noqa: E402

pytestmark = pytest.mark.gpu


def _collate(samples, M):
    """collate_fn_aux of the reference (src/util/misc.py:273-280) with the batch size rounded up to a multiple of 32"""
    images = M.nested_tensor_from_tensor_list([s[0] for s in samples], size_divisibility=32)
    depth = M.nested_tensor_from_tensor_list([s[1].float() for s in samples], size_divisibility=32)
    seg = M.nested_tensor_from_tensor_list([s[2] for s in samples], size_divisibility=32)
    return images, depth.tensors, seg.tensors, [s[3] for s in samples]


def test_augment_collate_train_loss_goes_down():
    import types
    import test_data_gpu as TG
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import data, model as M
    from gwdepth_b200.train_model import Trainer
    tf = data.make_coco_transforms("train", types.SimpleNamespace(eval=False))
    random.seed(11)
    torch.manual_seed(11)
    samples = []
    seed = 0
    while len(samples) < 2:
        img, depth, seg, target = TG._sample(seed)
        seed += 1
        try:
            o_img, o_tgt, (o_dep, o_seg) = tf(torch.from_numpy(img).cuda(), target,
                                              aux_mats=[torch.from_numpy(depth).cuda(), torch.from_numpy(seg).cuda()])
        except ImportError:      # the crop branch that needs shapely
            continue
        if o_tgt["lines"].shape[0] == 0:
            continue
        image, depth_gt, seg_gt, tgt = data.finish_sample(o_img, o_dep, o_seg, o_tgt, with_center=True)
        samples.append((image, depth_gt, seg_gt, {"lines": tgt["lines"].float().cuda(), "labels": tgt["labels"].cuda()}))
    images, depth_gt, seg_gt, targets = _collate(samples, M)
    B, _, H, W = images.tensors.shape
    assert H % 32 == 0 and W % 32 == 0 and (H // 32) * (W // 32) <= 1280
    mask = images.mask if images.padded else None
    _, crit, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
    criterion = crit[0].cuda()
    tr = Trainer(synth_weights(), lr=2e-5, lr_backbone=2e-6, weight_decay=1e-4, max_norm=0.1)
    losses = []
    for _ in range(12):
        total, parts = tr.train_step(images.tensors, targets, depth_gt, seg_gt, criterion, mask=mask)
        losses.append(float(total))
        assert all(torch.isfinite(torch.as_tensor(v)).all() for v in parts.values())
    print("losses:", [round(l, 2) for l in losses])
    assert all(l == l and l < 1e6 for l in losses), losses
    # the same batch twelve times at a fifth of the reference's learning rate: the loss falls (the uncertainty sampling and the
    # matching may change from step to step, so single steps are allowed to go up)
    assert sum(losses[-3:]) / 3 < sum(losses[:3]) / 3, losses
