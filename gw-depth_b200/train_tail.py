"""Training step of the DENSE TAIL on B200: everything behind the last Swin stage of the dense branch -- the 1/4-scale
point-based depth prediction (PointBasedPred + PyramidLayer, K = 80 points), the dense prediction head up to full
resolution and the three losses that sit on them -- as ONE forward / backward / optimizer step.  At 480x640 these are
57 % of the model's forward FLOPs (205 of 359 GFLOP per image: pyramid 188 at both scales, of which this scale 166; head 17).

Reference (under torch.autograd): `ReferTransformer.forward` tail (src/models/multiscale_transformerr.py:1263-1319:
`point_based_pred2` on [x3 | depth token], previous depth, sample points), `DensePrediction.forward`
(src/models/dense_upsample.py:160-182) and the loss loop of the engine (src/engine_glassrgbd.py:65-90): SilogLoss on
depth_pred3 (weight depth_loss_weights[2] = 0.25; the intermediate maps are compared in their own [0,1] units, as the
reference does) and on the full-resolution depth (weight 1), SegLoss x 2.

B200 design: train_points.PointPred and train_dense.DenseHead share the 1/4-scale stage buffer
[x3 | depth token | seg token | depth_pred3, pad] in place (no concat is ever materialised; depth_pred3 is written into its
column between the two forwards); the backward runs head -> (d depth_pred3 column + its own silog gradient) -> point
prediction, and the two gradients of the buffer are summed; the optimizer clips with ONE norm over the four flat buffers
(head, point projections, pyramid).  Inputs that come from the (not yet built) backward of the Swin stages are returned:
d(stage buffer) and d(depth_pred2).
"""
import torch

from . import ops
from .engine import DEFAULT_CFG
from .train_dense import DenseHead
from .train_flat import FlatModule
from .train_points import PointPred


class DenseTail:
    def __init__(self, state_dict, cfg=None, device="cuda", scale3_weight=0.25, **optim):
        self.cfg = c = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        self.C, self.td = c["dense_trans_dim"] >> 3, c["class_token_dim"]
        self.width = self.C + 3 * self.td
        self.K = c["interval_sample_num"][1]
        self.scale3_weight = scale3_weight
        self.point = PointPred(state_dict, "dense_encoder.point_based_pred2.", self.C, self.td, self.K, in_width=self.width,
                               device=device, **optim)
        self.head = DenseHead(state_dict, cfg, device=device, **optim)
        self.loss3 = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)

    def modules(self):
        return [self.head, self.point, self.point.pyramid]

    def state_dict(self):
        sd = self.head.state_dict()
        sd.update(self.point.state_dict())
        return sd

    def grads(self):
        g = self.head.grads()
        g.update(self.point.grads())
        return g

    def forward(self, buf4, depth2, coords, pos, H, W):
        """buf4: bf16 [B, H/4, W/4, width] stage buffer (the depth_pred3 column and the padding are overwritten); depth2:
        fp32 [B, H/8, W/8]; coords: fp32 [B,K,2]; pos: fp32 [H/4*W/4, C] position table.  Returns (depth3 fp32 [B,H/4,W/4] in
        [0,1], depth fp32 [B,1,H,W] metres, seg fp32 [B,2,H,W] logits); the modules keep their tapes."""
        B, H4, W4, width = buf4.shape
        col = self.C + 2 * self.td
        buf2d = buf4.view(B * H4 * W4, width)
        buf2d[:, col:] = 0
        depth3 = self.point.forward(buf2d, depth2, coords, pos, B, H4, W4)
        buf2d[:, col] = depth3.reshape(-1).to(torch.bfloat16)
        depth, seg = self.head.forward(buf4, H, W)
        self._shape = (B, H4, W4)
        return depth3, depth, seg

    def backward(self, g_depth3, g_depth_rows, g_seg_rows):
        """g_depth3: fp32 [B,H/4,W/4] gradient of depth_pred3 from ITS OWN loss (the head's use of it is added here);
        g_depth_rows / g_seg_rows: bf16 [B*H*W, 16] gradients to the outputs of get_depth (ahead of the sigmoid) / get_seg.
        Returns (d(buf4) bf16 [rows, width] with zeros beyond the token columns, d(depth2) fp32 [B, H/8, W/8])."""
        B, H4, W4 = self._shape
        C, td = self.C, self.td
        col = C + 2 * td
        d_buf = self.head.backward(g_depth_rows, g_seg_rows)
        d_depth3 = g_depth3.view(B, H4, W4) + d_buf[:, col].float().view(B, H4, W4)
        d_pp, d_depth2 = self.point.backward(d_depth3)
        d_buf[:, :C + td] += d_pp[:, :C + td]
        d_buf[:, col:] = 0
        return d_buf, d_depth2

    def loss_and_grads(self, buf4, depth2, coords, pos, depth_gt, seg_gt):
        """forward + the three losses that sit on this tail + backward.  depth_gt fp32 [B,1,H,W] metres; seg_gt int64 [B,1,H,W].
        Returns (depth3, depth, seg, losses fp32 [3] = weighted (scale-3 depth, full depth, seg), d(buf4), d(depth2))."""
        B, H4, W4, _ = buf4.shape
        H, W = depth_gt.shape[-2:]
        depth3, depth, seg = self.forward(buf4, depth2, coords, pos, H, W)
        g_depth_rows, g_seg_rows = self.head.loss_grads(depth, seg, depth_gt, seg_gt)
        log_only = bool(self.cfg.get("log_depth_error", False))
        p3 = depth3.view(B, 1, H4, W4)
        sums = ops.silog_sums(p3, depth_gt, log_only=log_only)
        d3 = ops.silog_bwd(p3, depth_gt, sums, weight=self.scale3_weight, log_only=log_only,
                           variance_focus=float(self.cfg.get("variance_focus", 0.85)), loss_out=self.loss3)
        d_buf, d_depth2 = self.backward(d3, g_depth_rows, g_seg_rows)
        return depth3, depth, seg, torch.cat([self.loss3, self.head.losses]), d_buf, d_depth2

    def step(self):
        """one gradient exchange per flat buffer, ONE clip norm over all of them (src/engine_glassrgbd.py:155-159), AdamW"""
        mods = self.modules()
        for m in mods:
            m.allreduce_grads()
        self.sumsq.zero_()
        for m in mods:
            ops.sumsq(m.G, self.sumsq)
        for m in mods:
            FlatModule.step(m, self.sumsq, reduced=True)      # not PointPred.step: that one clips over itself + pyramid only

    def train_step(self, buf4, depth2, coords, pos, depth_gt, seg_gt):
        losses = self.loss_and_grads(buf4, depth2, coords, pos, depth_gt, seg_gt)[3]
        self.step()
        return losses

