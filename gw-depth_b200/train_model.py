"""WHOLE-MODEL training step on B200: forward of `GlassRGBD` with the activations kept, the reference's losses, the backward of
every trained parameter, the data-parallel gradient exchange and the clipped AdamW update -- what one iteration of
`train_one_epoch` (src/engine_glassrgbd.py:45-166) does under torch.autograd + DistributedDataParallel
(src/main_glassrgbd.py:46-67):

    outputs = model(samples)                                   src/models/glassrgbd.py:74-123
    SetCriterion (6 Hungarian matchings), SilogLoss x 4 (weights 1/4, 1/4, 1/4, 1), SegLoss x 2
    losses.backward();  clip_grad_norm_(model.parameters(), 0.1);  AdamW (lr 1e-4, backbone 1e-5, weight decay 1e-4)

The step is assembled from the stage modules, each with its parameters in flat fp32 buffers in the layout the kernels read:

    train_backbone.BackboneTrain   stem + layer1 frozen, layer2-4 trained (FrozenBatchNorm folded)          -> C2..C5
    train.LineBranch               input_proj, DETR encoder / decoder, class_embed / lines_embed           -> logits, lines
    train_line_stage.LineStage     dense_input_proj + the 1/32 line-window stage (glass-structure context) -> x32, depth0
    train_branch.DenseBranch       class-window stages at 1/16, 1/8, 1/4, point predictions, dense head    -> 4 depth maps, seg

Order of a step (one CUDA stream + NCCL's): backbone -> line branch -> matching costs to the host (asynchronous) -> line-window
stage -> dense branch forward -> dense losses -> dense / line-window backward (GPU) WHILE the host solves the 6 x B assignments ->
set loss -> line-branch backward -> backbone backward.  Every module's flat gradient buffer is all-reduced (NCCL, asynchronous)
as soon as its backward has been enqueued, so the exchange overlaps the rest of the backward; ONE clip norm is taken over all
buffers after the exchange (the reference clips the whole model) and AdamW runs per buffer (gwd_sumsq, gwd_adamw_step).
The 54 trainable tensors that never receive a gradient in the reference (depth_pred32, layer4 of the pyramids,
proj_seg, ...) are not stored in any flat buffer: torch's AdamW skips parameters whose grad is None, so they stay constant.
"""
import torch
import torch.distributed as dist

from . import ops, parallel
from .engine import level_mask, DEFAULT_CFG
from .train import LineBranch
from .train_backbone import BackboneTrain
from .train_branch import DenseBranch
from .train_flat import FlatModule
from .train_line_stage import LineStage


class Trainer:
    def __init__(self, state_dict, cfg=None, device="cuda", lr=1e-4, lr_backbone=1e-5, weight_decay=1e-4, max_norm=0.1,
                 betas=(0.9, 0.999), eps=1e-8, depth_loss_weights=(0.25, 0.25, 0.25, 1.0), seg_loss_weight=2.0):
        self.cfg = c = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("Trainer runs on libgwd_b200 CUDA kernels only (no CPU fallback)")
        sd = {k: v.detach().to(self.dev) for k, v in state_dict.items()}
        kw = dict(device=self.dev, weight_decay=weight_decay, max_norm=max_norm, betas=betas, eps=eps)
        self.max_norm = max_norm
        self.backbone = BackboneTrain(sd, lr=lr_backbone, **kw)
        self.line = LineBranch(sd, c, lr=lr, **kw)
        self.line.use_cuda_graph = False
        self.stage32 = LineStage(sd, c, lr=lr, **kw)
        self.dense = DenseBranch(sd, c, scale_weights=tuple(depth_loss_weights[:3]), lr=lr, **kw)
        self.dense.tail.scale3_weight = depth_loss_weights[2]
        self.dense.tail.head.depth_weight, self.dense.tail.head.seg_weight = depth_loss_weights[3], seg_loss_weight
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._works = []
        self.exchange_grads = True
        self._defer = None                 # list collecting the modules to exchange while a CUDA graph is captured / replayed
        import os
        # the fused step replays its three kernel sequences as CUDA graphs (GWD_CUDA_GRAPH=0: launched kernel by kernel)
        self.use_cuda_graph = os.environ.get("GWD_CUDA_GRAPH", "1") != "0"
        self._graphs = {}
        self.max_graphs = 6          # captured training steps kept (one memory pool each), least recently used first out
        self._masks = None
        self.last = {}

    # ------------------------------------------------------------------ bookkeeping
    def modules(self):
        """every flat-buffer module, in the order their backward completes"""
        return self.dense.modules() + [self.stage32, self.line, self.backbone]

    def state_dict(self):
        """logical fp32 parameters under the reference's key names (the trained tensors only)"""
        sd = {}
        for m in (self.backbone, self.line, self.stage32, self.dense):
            sd.update(m.state_dict())
        return sd

    def grads(self):
        """gradients of the last backward under the reference's key names and shapes"""
        g = {}
        for m in (self.backbone, self.line, self.stage32, self.dense):
            g.update(m.grads())
        return g

    def load_params(self, state_dict):
        """overwrite every flat master buffer (and its bf16 mirror) from a reference-keyed state dict: the drop-in path calls
        this when a torch optimizer has updated the nn.Parameters; the optimizer moments of the fused step are kept"""
        for m in self.modules():
            if isinstance(m, FlatModule):
                FlatModule.load_params(m, state_dict)
            else:
                m.load_params(state_dict)

    def numel(self):
        return sum(m.numel for m in self.modules())

    def set_lr(self, lr, lr_backbone=None):
        """what `StepLR.step()` does to the reference's two parameter groups (src/main_glassrgbd.py:59-67)"""
        for m in self.modules():
            if m is self.backbone:
                m.lr = lr_backbone if lr_backbone is not None else m.lr
            else:
                m.lr = lr

    def optimizer_state(self):
        """Adam moments and step counts of every flat buffer (physical layout) for a resumable checkpoint: the counterpart of
        `optimizer.state_dict()` in src/main_glassrgbd.py:214-226"""
        return {"format": "gwd_flat_adamw_v1",
                "buffers": [{"numel": m.numel, "t": m.t, "lr": m.lr, "M": m.M.detach().cpu(), "V": m.V.detach().cpu()} for m in self.modules()]}

    def load_optimizer_state(self, state):
        bufs = state["buffers"]
        assert state.get("format") == "gwd_flat_adamw_v1" and len(bufs) == len(self.modules()), "optimizer state of another layout"
        for m, b in zip(self.modules(), bufs):
            assert b["numel"] == m.numel
            m.M.copy_(b["M"])
            m.V.copy_(b["V"])
            m.t, m.lr = b["t"], b["lr"]

    # ------------------------------------------------------------------ forward
    def reference_points(self, logits, lines):
        """top-num_ref lines by RAW line logit -> end points (and centre with --with_dense_center) in [-1,1]
        (src/models/multiscale_transformerr.py:1165-1179); no gradient flows through the selection or the coordinates"""
        c = self.cfg
        return ops.select_lines(logits.contiguous(), lines.contiguous(), c["num_ref"], 3 if c["with_dense_center"] else 2)

    def forward(self, images, pinned=None, after_line=None, mask=None):
        """images fp32 [B,3,H,W] (H and W multiples of 32); mask: None for an equal-size batch, else the bool [B,H,W] padding mask
        of nested_tensor_from_tensor_list (True = padding; the batch must be padded to a multiple of 32).  Returns (logits [S,B,Q,2], lines [S,B,Q,D] fp32 of all
        decoder stages, final stage LAST; dense outputs dict of DenseBranch.forward).  `after_line(logits, lines)` is called as
        soon as the line branch is enqueued (the fused step starts the matching there)."""
        B, _, H, W = images.shape
        if H % 32 or W % 32:
            raise NotImplementedError("the training path is built for input sizes that are multiples of 32 (exact x2 pyramids)")
        pinned = pinned or {}
        bb = self.backbone
        c2 = bb.frozen_front(images.float().contiguous())
        c3, c4, c5 = bb.forward(c2)
        masks = None if mask is None else [level_mask(mask, f.shape[1:3]) for f in (c2, c3, c4, c5)]
        logits, lines = self.line.forward(c5, None if masks is None else masks[3])
        if after_line is not None:
            after_line(logits, lines)
        if "line_ids" in pinned:
            ids = pinned["line_ids"]
            chosen = torch.gather(lines[-1], 1, ids[:, :, None].expand(-1, -1, lines.shape[-1]))
            pts = chosen.reshape(B, self.cfg["num_ref"], -1, 2) * 2 - 1.0
            ref_xy = (pts if self.cfg["with_dense_center"] else pts[:, :, :2]).reshape(B, -1, 2).contiguous().float()
        else:
            ref_xy, ids = self.reference_points(logits[-1], lines[-1])
        x32, depth0 = self.stage32.forward(c5, ref_xy, None if masks is None else masks[3])
        outs = self.dense.forward(x32, depth0, (c4, c3, c2), H, W, {k: v for k, v in pinned.items() if k.startswith("sample")},
                                  pad_masks=None if masks is None else (masks[2], masks[1], masks[0]))
        outs.update(line_ids=ids, depth0=depth0)
        self._c5_shape = c5.shape
        return logits, lines, outs

    # ------------------------------------------------------------------ backward
    def _exchange(self, mods):
        """asynchronous all-reduce of flat gradient buffers whose backward has been enqueued (exchange_grads = False: someone
        else reduces the gradients, e.g. DistributedDataParallel around the drop-in module)"""
        if self._defer is not None:          # inside a captured region: the collective is issued after the replay
            self._defer += mods
            return
        if self.exchange_grads:
            self._works += parallel.allreduce_async([m.G for m in mods])

    def backward_dense(self, g_depth1, g_depth2, g_depth3, g_depth_rows, g_seg_rows):
        """dense branch + 1/32 line-window stage; returns its part of d C5 (and keeps d C4, d C3 for `backward_line`)"""
        d_x32, d_c4, d_c3 = self.dense.backward(g_depth1, g_depth2, g_depth3, g_depth_rows, g_seg_rows)
        self._exchange(self.dense.modules())
        d_c5 = self.stage32.backward(d_x32)
        self._exchange([self.stage32])
        self._dense_grads = (d_c3, d_c4, d_c5)

    def backward_line(self, dlogits, dlines):
        """line branch, then the backbone with the summed gradients of C3, C4, C5"""
        d_c3, d_c4, d_c5 = self._dense_grads
        d_c5b = self.line.backward(dlogits, dlines)
        self._exchange([self.line])
        d_c5 = ops.add_rows(d_c5, d_c5b, d_c5.shape[0]).view(self._c5_shape)
        self.backbone.backward(d_c3, d_c4, d_c5)
        self._exchange([self.backbone])
        self._dense_grads = None

    # ------------------------------------------------------------------ optimizer
    def step(self):
        """wait for the gradient exchange, ONE clip norm over every flat buffer (src/engine_glassrgbd.py:155-159), AdamW"""
        world = parallel.wait_all(self._works)
        self._works = []
        mods = self.modules()
        self.sumsq.zero_()
        for m in mods:
            ops.sumsq(m.G, self.sumsq)
        for m in mods:
            if isinstance(m, FlatModule):
                m._world = world
                # (not m.step: PointPred / DenseTail override it to clip over their own buffers only)
                (BackboneTrain.step if isinstance(m, BackboneTrain) else FlatModule.step)(m, self.sumsq, reduced=True)
            else:       # LineBranch: the same flat-buffer update
                m.t += 1
                ops.adamw_step(m.P, m.G, m.M, m.V, m.Wb, lr=m.lr, betas=m.betas, eps=m.eps, weight_decay=m.weight_decay, step=m.t,
                               max_norm=m.max_norm, grad_scale=1.0 / world, sumsq_buf=self.sumsq)

    def grad_norm(self):
        """the global gradient norm the last `step` clipped with (one host read)"""
        return float(self.sumsq.sqrt().item()) / parallel.world_size()

    # ------------------------------------------------------------------ one fused training step
    def train_step(self, images, targets, depth_gt, seg_gt, criterion, pinned=None, mask=None):
        """images fp32 [B,3,H,W]; targets: list of {'lines' [T,D], 'labels' [T]} on the device; depth_gt fp32 [B,1,H,W] metres;
        seg_gt int64 [B,1,H,W]; criterion: model.SetCriterion.  Returns (total loss tensor [1] on the device, dict of the 17
        un-weighted losses as the engine logs them)."""
        if self.use_cuda_graph and not pinned:
            return self._train_step_graphed(images, targets, depth_gt, seg_gt, criterion, mask)
        pend = {}
        logits, lines, outs = self.forward(images, pinned, mask=mask,
                                           after_line=lambda lo, li: pend.update(h=criterion.matcher.stacked_cost(lo, li, targets)))
        g = self.dense.loss_grads(outs, depth_gt, seg_gt)
        self.backward_dense(*g)                                   # enqueued; the GPU works on it while the host solves the LSAPs
        set_losses, dlogits, dlines = criterion.forward_backward_stacked(logits, lines, targets, pending=pend["h"])
        self.backward_line(dlogits, dlines)
        self.step()
        return self._report(criterion, set_losses, outs, logits, lines)

    def _report(self, criterion, set_losses, outs, logits, lines):
        dl = self.dense.losses()
        total = criterion.last_total + dl.sum()
        losses = dict(set_losses)
        losses.update(loss_depth=dl[:4].sum(), loss_seg=dl[4] / self.dense.tail.head.seg_weight)
        self.last = dict(outs=outs, logits=logits, lines=lines, dense_losses=dl)
        return total, losses

    # ---- the same step as three CUDA-graph replays around the two host-side pieces (matching costs / assignments)
    def _forward_line(self, images, mask=None):
        c2 = self.backbone.frozen_front(images)
        c3, c4, c5 = self.backbone.forward(c2)
        # ragged batch: the level masks are built from the (static) batch mask inside the captured region
        self._masks = None if mask is None else [level_mask(mask, f.shape[1:3]) for f in (c2, c3, c4, c5)]
        logits, lines = self.line.forward(c5, None if mask is None else self._masks[3])
        self._c5_shape = c5.shape
        return (c2, c3, c4, c5), logits, lines

    def _dense_part(self, feats, logits, lines, depth_gt, seg_gt, H, W):
        c2, c3, c4, c5 = feats
        m = self._masks
        ref_xy, ids = self.reference_points(logits[-1], lines[-1])
        x32, depth0 = self.stage32.forward(c5, ref_xy, None if m is None else m[3])
        outs = self.dense.forward(x32, depth0, (c4, c3, c2), H, W, pad_masks=None if m is None else (m[2], m[1], m[0]))
        outs.update(line_ids=ids, depth0=depth0)
        self.backward_dense(*self.dense.loss_grads(outs, depth_gt, seg_gt))
        return outs

    def _capture(self, images, depth_gt, seg_gt, criterion, targets, mask=None):
        B, _, H, W = images.shape
        if H % 32 or W % 32:
            raise NotImplementedError("the training path is built for input sizes that are multiples of 32 (exact x2 pyramids)")
        st = dict(images=images.float().clone(), depth_gt=depth_gt.clone(), seg_gt=seg_gt.clone(),
                  mask=None if mask is None else mask.clone())
        # warm-up (eager, on a side stream): lazy kernel attributes, cached tables, transpose tables
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._defer = []
            feats, lo, li = self._forward_line(st["images"], st["mask"])
            self._dense_part(feats, lo, li, st["depth_gt"], st["seg_gt"], H, W)
            _, dlo, dli = criterion.forward_backward_stacked(lo, li, targets)
            self.backward_line(dlo, dli)
            self._defer = None
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self._defer = []
        from . import capi
        n0 = capi.launch_count()          # library kernels enqueued while capturing = library kernels every replay runs
        st["g1"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st["g1"]):
            st["feats"], st["logits"], st["lines"] = self._forward_line(st["images"], st["mask"])
        st["masks"] = self._masks            # level masks: outputs of graph 1, read by graph 2
        st["g2"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st["g2"], pool=st["g1"].pool()):
            self._masks = st["masks"]
            st["outs"] = self._dense_part(st["feats"], st["logits"], st["lines"], st["depth_gt"], st["seg_gt"], H, W)
        st["ex2"], self._defer = self._defer, []
        st["dlogits"], st["dlines"] = torch.zeros_like(st["logits"]), torch.zeros_like(st["lines"])
        st["g3"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st["g3"], pool=st["g1"].pool()):
            self.backward_line(st["dlogits"], st["dlines"])
        st["ex3"], self._defer = self._defer, None
        st["kernels_per_replay"] = capi.launch_count() - n0
        return st

    def _train_step_graphed(self, images, targets, depth_gt, seg_gt, criterion, mask=None):
        key = (tuple(images.shape), tuple(depth_gt.shape), id(criterion), mask is not None)
        st = self._graphs.pop(key, None)
        if st is None:
            # one memory pool per input shape: a ragged data set cycles through a handful of padded sizes; keep the most recent ones
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            st = self._capture(images, depth_gt, seg_gt, criterion, targets, mask)
        self._graphs[key] = st               # (re-)inserted last = most recently used
        if mask is not None:
            st["mask"].copy_(mask, non_blocking=True)
        self.replayed_kernels = getattr(self, "replayed_kernels", 0) + st["kernels_per_replay"]
        st["images"].copy_(images, non_blocking=True)
        st["depth_gt"].copy_(depth_gt, non_blocking=True)
        st["seg_gt"].copy_(seg_gt, non_blocking=True)
        st["g1"].replay()                                                    # backbone + line branch forward
        pend = criterion.matcher.stacked_cost(st["logits"], st["lines"], targets)
        st["g2"].replay()                                                    # 1/32 stage + dense branch: forward, losses, backward
        self._defer = None
        self._exchange(st["ex2"])
        set_losses, dlogits, dlines = criterion.forward_backward_stacked(st["logits"], st["lines"], targets, pending=pend)
        st["dlogits"].copy_(dlogits, non_blocking=True)
        st["dlines"].copy_(dlines, non_blocking=True)
        st["g3"].replay()                                                    # line branch + backbone backward
        self._exchange(st["ex3"])
        self.step()
        return self._report(criterion, set_losses, st["outs"], st["logits"], st["lines"])
