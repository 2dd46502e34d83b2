"""Dense evaluation on the device: the depth and segmentation bookkeeping of `evaluate` in
src/engine_glassrgbd.py:232-264,309-322 without copying a map to the host.

The reference evaluates at batch 1: per image it moves depth and segmentation maps to the host, scrubs the prediction,
runs util/metrics.compute_depth_errors in numpy and appends the argmax map to a list for compute_mean_ioU at the end.
Here a batch of any size is reduced by two kernels (`gwd_depth_metrics`: nine metrics per image, `gwd_seg_confusion`:
argmax + confusion counts), the running sums stay on the device, and `summary()` is the only synchronisation; under
torch.distributed it all-reduces [9 sums, count] and the confusion matrix, so every rank reports the numbers
of the whole evaluation set.  Per-image metrics inside a batch equal the batch-1 loop (tests/test_kernels_gpu.py)."""
import torch
import torch.distributed as dist

from . import ops, parallel

DEPTH_METRICS = ("silog", "abs_rel", "log10", "rms", "sq_rel", "log_rms", "d1", "d2", "d3")
SEG_LABELS = ("Background", "Glass")       # util/metrics.py labels_name_dicts['GLASS']


class DenseEvaluator:
    def __init__(self, device="cuda", num_classes=2, min_depth_eval=1e-3, max_depth_eval=10.0, ignore_index=255):
        self.dev = torch.device(device)
        self.lo, self.hi, self.ignore = float(min_depth_eval), float(max_depth_eval), ignore_index
        self.depth_sums = torch.zeros(10, dtype=torch.float64, device=self.dev)       # 9 metric sums + image count
        self.confusion = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=self.dev)

    @torch.no_grad()
    def update(self, outputs, depth_gt, seg_gt):
        """outputs: the model's dict; depth_gt fp32 [B,1,H,W] or [B,H,W] metres; seg_gt integer [B,1,H,W] or [B,H,W]"""
        pred = outputs["pred_depth"][-1] if isinstance(outputs["pred_depth"], (list, tuple)) else outputs["pred_depth"]
        seg = outputs["pred_seg"][-1] if isinstance(outputs["pred_seg"], (list, tuple)) else outputs["pred_seg"]
        B = pred.shape[0]
        per_image = ops.depth_metrics(pred.reshape(B, *pred.shape[-2:]).float().contiguous(),
                                      depth_gt.reshape(B, *depth_gt.shape[-2:]).float().contiguous(), self.lo, self.hi)
        self.depth_sums[:9] += per_image.sum(0)
        self.depth_sums[9] += B
        ops.seg_confusion(seg.float(), seg_gt, confusion=self.confusion, ignore_index=self.ignore)
        return per_image

    def summary(self):
        """-> dict with the reference's names: the nine depth metrics averaged over images, IoU per class, 'Pixel accuracy',
        'Mean accuracy', 'Mean IU' (src/util/metrics.py:80-91)"""
        sums, conf = self.depth_sums.clone(), self.confusion.clone()
        if parallel.world_size() > 1:
            dist.all_reduce(sums)
            dist.all_reduce(conf)
        iou, pix, macc, miou = ops.seg_scores(conf)
        vals = torch.cat([sums[:9] / sums[9].clamp_min(1.0), iou, torch.stack([pix, macc, miou])]).cpu().tolist()
        names = list(DEPTH_METRICS) + list(SEG_LABELS[: iou.numel()]) + ["Pixel accuracy", "Mean accuracy", "Mean IU"]
        out = dict(zip(names, vals))
        out["images"] = int(sums[9].item())
        return out
