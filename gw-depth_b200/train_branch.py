"""Training step of the DENSE BRANCH behind the 1/32 line-window stage on B200: the three class-window Swin stages (1/16, 1/8,
1/4) with their entries, the coarse depth head, both point-based depth predictions, the uncertainty sampling between them,
the dense prediction head and all five losses of the dense branch, as one forward / backward / optimizer step.

Reference (under torch.autograd): `ReferTransformer.forward` from the 1/16 stage on
(src/models/multiscale_transformerr.py:1191-1319), `DensePrediction.forward` (src/models/dense_upsample.py:160-182) and the
loss loop of the engine (src/engine_glassrgbd.py:65-90): SilogLoss on depth_pred1..3 (weights 1/4 each, in their own [0,1]
units) and on the full-resolution depth (weight 1), SegLoss x 2.  The uncertainty sampling (`CertainSample`) is a top-k
selection: the sample coordinates carry no gradient; the ANCHOR depths sampled at them do (into the previous scale's depth).

This module only orchestrates the stage modules (train_entry, train_swin, train_points, train_tail); every kernel runs
through the C ABI.  `backward` returns what the modules in front of it continue with: d(x32) (train_line_stage.LineStage, the
line-window stage at 1/32) and d(C4), d(C3) (train_backbone.BackboneTrain; C2 comes from the frozen layer1 and needs no gradient);
train_model.Trainer joins them into the whole-model step.
"""
import torch

from . import ops
from .engine import DEFAULT_CFG, sine_table, sine_tables_masked
from .train_entry import DepthHead16, StageEntry
from .train_flat import FlatModule
from .train_points import PointPred
from .train_swin import ClassStage
from .train_tail import DenseTail


class DenseBranch:
    def __init__(self, state_dict, cfg=None, device="cuda", scale_weights=(0.25, 0.25, 0.25), cuda_graph=False, **optim):
        self.cfg = c = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        D, td, nh, ws = c["dense_trans_dim"], c["class_token_dim"], c["dense_trans_heads"], c["window"]
        self.D, self.td = D, td
        self.Cs = (D >> 1, D >> 2, D >> 3)
        self.scale_weights = scale_weights
        kw = dict(device=device, **optim)
        self.entries = [StageEntry(state_dict, si, **kw) for si in (1, 2, 3)]
        self.stages = [ClassStage(state_dict, "dense_encoder.class_transformer%d." % si, C, depth, heads=nh, ws=ws, token_dim=td, **kw)
                       for si, C, depth in zip((1, 2, 3), self.Cs, c["class_trans_layers"])]
        self.head16 = DepthHead16(state_dict, **kw)
        self.point1 = PointPred(state_dict, "dense_encoder.point_based_pred1.", self.Cs[1], td, c["interval_sample_num"][0],
                                in_width=self.Cs[1] + td, **kw)
        self.tail = DenseTail(state_dict, cfg, **kw)
        self.edges = [c["min_depth_eval"] / c["max_depth_eval"]] + list(c["depth_interval"]) + [1.0]
        self.loss12 = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._pos = {}
        # forward + losses + backward have no host synchronisation and no data-dependent control flow: with cuda_graph=True
        # the ~720 launches are captured once per input shape and replayed as one graph (the optimizer step stays outside:
        # its bias corrections are host scalars)
        self.use_cuda_graph, self._graphs = cuda_graph, {}

    def modules(self):
        return self.entries + self.stages + [self.head16, self.point1, self.point1.pyramid] + self.tail.modules()

    def state_dict(self):
        sd = {}
        for m in self.entries + self.stages + [self.head16, self.point1, self.tail]:
            sd.update(m.state_dict())
        return sd

    def grads(self):
        g = {}
        for m in self.entries + self.stages + [self.head16, self.point1, self.tail]:
            g.update(m.grads())
        return g

    def _table(self, H, W, C, pad_mask=None):
        if pad_mask is not None:        # ragged batch: per-image position codes from the padding mask of this level
            return sine_tables_masked(pad_mask, C // 2, False)
        key = (H, W, C)
        if key not in self._pos:
            self._pos[key] = sine_table(H, W, C // 2, False, self.dev)
        return self._pos[key]

    def _silog(self, pred, depth_gt, weight, loss_out):
        log_only = bool(self.cfg.get("log_depth_error", False))
        sums = ops.silog_sums(pred, depth_gt, log_only=log_only)
        return ops.silog_bwd(pred, depth_gt, sums, weight=weight, log_only=log_only,
                             variance_focus=float(self.cfg.get("variance_focus", 0.85)), loss_out=loss_out)

    def loss_and_grads(self, x32, depth0, feats, depth_gt, seg_gt, pinned=None):
        """see _loss_and_grads; replays the captured graph when cuda_graph=True (the returned tensors are then the graph's
        static outputs: consume them before the next call with the same shapes)"""
        if not self.use_cuda_graph:
            return self._loss_and_grads(x32, depth0, feats, depth_gt, seg_gt, pinned)
        pinned = pinned or {}
        flat = [x32, depth0, *feats, depth_gt, seg_gt] + [pinned[k] for k in sorted(pinned)]
        key = tuple((tuple(t.shape), t.dtype) for t in flat) + tuple(sorted(pinned))
        st = self._graphs.get(key)
        if st is None:
            static = [t.clone() for t in flat]
            call = lambda: self._loss_and_grads(static[0], static[1], static[2:5], static[5], static[6],  # noqa: E731
                                                dict(zip(sorted(pinned), static[7:])))
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up: lazy kernel attributes, cached tables, transpose tables
                call()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                result = call()
            st = self._graphs[key] = dict(graph=graph, static=static, result=result)
        for dst, src in zip(st["static"], flat):
            dst.copy_(src, non_blocking=True)
        st["graph"].replay()
        return st["result"]

    def forward(self, x32, depth0, feats, H, W, pinned=None, pad_masks=None):
        """x32 bf16 [B,h,w,D] (output of the 1/32 line-window stage); depth0 fp32 [B,h,w] (depth_pred32 of it, only feeds the
        sampling); feats = (C4, C3, C2) bf16 channels-last backbone maps at 1/16, 1/8, 1/4; (H, W): the input size; pinned:
        optional {'sample1', 'sample2'} coordinates overriding the uncertainty sampling.  Returns the outputs dict
        (pred_depth = [depth1, depth2, depth3 fp32 [B,h_i,w_i] in [0,1], depth fp32 [B,1,H,W] metres], pred_seg, sample1/2);
        the stage modules keep their tapes for `backward`."""
        pinned = pinned or {}
        c, td = self.cfg, self.td
        C1, C2, C3 = self.Cs
        B = x32.shape[0]
        (H1, W1), (H2, W2), (H3, W3) = [f.shape[1:3] for f in feats]
        e1, e2, e3 = self.entries
        s1, s2, s3 = self.stages
        # ---- 1/16
        x, d, s = e1.forward(x32, None, None, feats[0])
        x1, d1, t1 = s1.forward(x, d, s, B, H1, W1)
        depth1 = self.head16.forward(x1, d1).view(B, H1, W1)
        coords1 = pinned["sample1"] if "sample1" in pinned else ops.certain_sample(depth0, depth1, c["interval_sample_num"][0], self.edges)[0]
        # ---- 1/8
        x, d, s = e2.forward(x1.view(B, H1, W1, C1), d1, t1, feats[1])
        x2, d2, t2 = s2.forward(x, d, s, B, H2, W2)
        buf2 = torch.cat([x2, d2], dim=1)
        coords1 = coords1.reshape(B, -1, 2).float().contiguous()
        depth2 = self.point1.forward(buf2, depth1, coords1, self._table(H2, W2, C2, None if pad_masks is None else pad_masks[1]), B, H2, W2)
        coords2 = pinned["sample2"] if "sample2" in pinned else ops.certain_sample(depth1, depth2, c["interval_sample_num"][1], self.edges)[0]
        # ---- 1/4 + head
        x, d, s = e3.forward(x2.view(B, H2, W2, C2), d2, t2, feats[2])
        x3, d3, t3 = s3.forward(x, d, s, B, H3, W3)
        buf4 = torch.zeros(B, H3, W3, C3 + 3 * td, dtype=torch.bfloat16, device=self.dev)
        b2d = buf4.view(-1, C3 + 3 * td)
        b2d[:, :C3], b2d[:, C3:C3 + td], b2d[:, C3 + td:C3 + 2 * td] = x3, d3, t3
        coords2 = coords2.reshape(B, -1, 2).float().contiguous()
        depth3, depth, seg = self.tail.forward(buf4, depth2, coords2, self._table(H3, W3, C3, None if pad_masks is None else pad_masks[2]), H, W)
        self._shapes = (B, (H1, W1), (H2, W2), (H3, W3))
        return dict(pred_depth=[depth1, depth2, depth3, depth], pred_seg=seg, sample1=coords1, sample2=coords2)

    def backward(self, g_depth1, g_depth2, g_depth3, g_depth_rows, g_seg_rows):
        """g_depth1..3: fp32 [B,h_i,w_i] gradients of the three coarse maps from THEIR OWN losses (what later stages add through
        the anchor depths is accumulated here); g_depth_rows / g_seg_rows: bf16 [B*H*W, 16] gradients to the outputs of
        get_depth (ahead of the sigmoid) / get_seg.  Fills every flat gradient buffer; returns (d x32 [B,h,w,D], d C4, d C3)."""
        td = self.td
        C1, C2, C3 = self.Cs
        B, (H1, W1), (H2, W2), (H3, W3) = self._shapes
        e1, e2, e3 = self.entries
        s1, s2, s3 = self.stages
        d_buf4, d_depth2 = self.tail.backward(g_depth3, g_depth_rows, g_seg_rows)
        g = s3.backward(d_buf4[:, :C3].contiguous(), d_buf4[:, C3:C3 + td].contiguous(), d_buf4[:, C3 + td:C3 + 2 * td].contiguous())
        d_x2, d_d2, d_t2, _ = e3.backward(*g)
        d_depth2 = d_depth2 + g_depth2.view(B, H2, W2)
        d_buf2, d_depth1 = self.point1.backward(d_depth2)
        g = s2.backward(d_x2.view(-1, C2) + d_buf2[:, :C2], d_d2 + d_buf2[:, C2:C2 + td], d_t2)
        d_x1, d_d1, d_t1, d_c3 = e2.backward(*g, need_dfeat=True)
        d_depth1 = d_depth1 + g_depth1.view(B, H1, W1)
        dh_x, dh_d = self.head16.backward(d_depth1.view(-1))
        g = s1.backward(d_x1.view(-1, C1) + dh_x, d_d1 + dh_d, d_t1)
        d_x32, _, _, d_c4 = e1.backward(*g, need_dfeat=True)
        return d_x32, d_c4, d_c3

    def loss_grads(self, outs, depth_gt, seg_gt):
        """the five dense losses of the engine's loop (src/engine_glassrgbd.py:65-90) on the forward outputs: fills the loss
        buffers and returns the five gradients `backward` takes"""
        B, (H1, W1), (H2, W2), (H3, W3) = self._shapes
        d1, d2, d3, depth = outs["pred_depth"]
        g1 = self._silog(d1.view(B, 1, H1, W1), depth_gt, self.scale_weights[0], self.loss12[0:1])
        g2 = self._silog(d2.view(B, 1, H2, W2), depth_gt, self.scale_weights[1], self.loss12[1:2])
        g3 = self._silog(d3.view(B, 1, H3, W3), depth_gt, self.scale_weights[2], self.tail.loss3)
        g_depth_rows, g_seg_rows = self.tail.head.loss_grads(depth, outs["pred_seg"], depth_gt, seg_gt)
        return g1, g2, g3, g_depth_rows, g_seg_rows

    def losses(self):
        """fp32 [5] on the device: weighted (depth1, depth2, depth3, depth, seg) of the last loss_grads call"""
        return torch.cat([self.loss12, self.tail.loss3, self.tail.head.losses])

    def _loss_and_grads(self, x32, depth0, feats, depth_gt, seg_gt, pinned=None):
        """forward + the five losses + backward.  depth_gt fp32 [B,1,H,W] metres; seg_gt int64 [B,1,H,W].
        Returns (outputs dict, losses fp32 [5] = weighted (depth1, depth2, depth3, depth, seg), d x32, d C4, d C3)."""
        H, W = depth_gt.shape[-2:]
        outs = self.forward(x32, depth0, feats, H, W, pinned)
        d_x32, d_c4, d_c3 = self.backward(*self.loss_grads(outs, depth_gt, seg_gt))
        return outs, self.losses(), d_x32, d_c4, d_c3

    def step(self):
        """one gradient exchange per flat buffer, ONE clip norm over all of them (src/engine_glassrgbd.py:155-159), AdamW"""
        mods = self.modules()
        for m in mods:
            m.allreduce_grads()
        self.sumsq.zero_()
        for m in mods:
            ops.sumsq(m.G, self.sumsq)
        for m in mods:
            FlatModule.step(m, self.sumsq, reduced=True)

    def train_step(self, *args, **kw):
        losses = self.loss_and_grads(*args, **kw)[1]
        self.step()
        return losses
