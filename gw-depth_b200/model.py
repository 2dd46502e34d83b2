"""Host-side mirror of the reference's model interface (src/models/__init__.py:5-6, src/models/glassrgbd.py:44-123,
133-358,360-383,452-506,509-579, src/models/matcher.py:10-87, src/util/misc.py:291-367): same names, argument
meaning and error behaviour, so that src/main_glassrgbd.py and src/engine_glassrgbd.py can import this module in
place of `models`.

    model, [criterion, criterion_depth, criterion_seg, criterion_plane], postprocessors = build_model(args)
    out = model(samples)      # {'pred_logits','pred_lines','aux_outputs','pred_depth': [4 maps],'pred_seg'}

`GlassRGBD` holds the reference's parameters under the reference's state_dict names (see spec.py) and runs its
forward through the sm_100a kernel plan in engine.py.  There is no PyTorch fallback: without libgwd_b200.so or
without a CUDA device the forward raises.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn

from . import engine as _engine
from . import ops
from .spec import model_spec


class NestedTensor(object):
    """(tensors, padding mask) pair, src/util/misc.py:347-367"""

    def __init__(self, tensors, mask, padded=None):
        # padded: host-side knowledge of whether any image was padded (None = unknown); lets the CUDA forward pick the
        # masked path without reading the mask back
        self.tensors, self.mask, self.padded = tensors, mask, padded

    def to(self, device):
        return NestedTensor(self.tensors.to(device), self.mask.to(device) if self.mask is not None else None, self.padded)

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list, size_divisibility=1):
    """pad a list of [C,H,W] images to a common size; mask is True on padding (src/util/misc.py:291-313).  size_divisibility (an
    extension; the reference pads to the largest image only): round the common size up to a multiple -- 32 for TRAINING on ragged
    batches, whose backward kernels want exact x2 pyramids; the extra rows / columns are ordinary padding (zeros, mask True)."""
    if isinstance(tensor_list, torch.Tensor) and tensor_list.dim() == 4:
        tensor_list = list(tensor_list)
    if tensor_list[0].dim() != 3:
        raise ValueError("not supported")
    c = tensor_list[0].shape[0]
    hmax = max(t.shape[1] for t in tensor_list)
    wmax = max(t.shape[2] for t in tensor_list)
    if size_divisibility > 1:
        hmax = -(-hmax // size_divisibility) * size_divisibility
        wmax = -(-wmax // size_divisibility) * size_divisibility
    batch = tensor_list[0].new_zeros((len(tensor_list), c, hmax, wmax))
    mask = torch.ones((len(tensor_list), hmax, wmax), dtype=torch.bool, device=batch.device)
    for i, t in enumerate(tensor_list):
        batch[i, :, : t.shape[1], : t.shape[2]].copy_(t)
        mask[i, : t.shape[1], : t.shape[2]] = False
    padded = any(t.shape[1] != hmax or t.shape[2] != wmax for t in tensor_list)
    return NestedTensor(batch, mask, padded)


class _Node(nn.Module):
    """anonymous container: gives parameters the reference's dotted names"""


def _build_tree(root, entries):
    for name, shape, dtype, kind, trainable in entries:
        *path, leaf = name.split(".")
        node = root
        for part in path:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        if kind == "param":
            node.register_parameter(leaf, nn.Parameter(torch.zeros(shape), requires_grad=trainable))
        else:
            dt = torch.int64 if dtype == "int64" else torch.float32
            node.register_buffer(leaf, torch.zeros(shape, dtype=dt))


def _init_parameters(module):
    """Random initialisation in the spirit of the reference (xavier for the DETR part, transformer.py:42-45;
    truncated normal 0.02 for the dense encoder, multiscale_transformerr.py:1140-1149): training from scratch is
    possible, but the usual entry is load_state_dict() of a reference checkpoint."""
    from .engine import rel_pos_bias  # noqa: F401  (structural buffer helper lives with the tables)
    for name, p in module.named_parameters():
        leaf = name.rsplit(".", 1)[-1]
        if "norm" in name and leaf == "weight":
            nn.init.ones_(p)
        elif leaf in ("bias", "in_proj_bias", "diff_logsigma", "border_logsigma"):
            nn.init.zeros_(p)
        elif p.dim() > 1 and (name.startswith("transformer") or p.dim() == 4):
            nn.init.xavier_uniform_(p)
        elif p.dim() > 1:
            nn.init.trunc_normal_(p, std=0.02)
        else:
            nn.init.normal_(p, std=0.02)
    for name, b in module.named_buffers():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "relative_position_index":
            ws = int(round(b.shape[0] ** 0.5))
            c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
            rel = c[:, :, None] - c[:, None, :]
            b.copy_((rel[0] + ws - 1) * (2 * ws - 1) + rel[1] + ws - 1)
        elif leaf in ("weight", "running_var"):
            b.fill_(1.0)


def _clone_outputs(out):
    if isinstance(out, dict):
        return {k: _clone_outputs(v) for k, v in out.items()}
    if isinstance(out, (list, tuple)):
        return type(out)(_clone_outputs(v) for v in out)
    return out.clone() if isinstance(out, torch.Tensor) else out


class _TrainStep(torch.autograd.Function):
    """autograd edge between the reference's training loop and train_model.Trainer: forward runs the engine's forward with
    the activations kept, backward takes the cotangents of every output the criteria touched, runs the engine's backward and
    returns the gradient of every trained nn.Parameter (so DistributedDataParallel's hooks fire exactly as in the reference;
    the 54 never-used tensors are not inputs, which is what find_unused_parameters=True expects)."""

    @staticmethod
    def forward(ctx, module, images, pinned, mask, *params):
        tr = module.trainer()
        with torch.no_grad():
            logits, lines, outs = tr.forward(images, pinned, mask=mask)
        d1, d2, d3, depth = outs["pred_depth"]
        ctx.module, ctx.depth = module, depth
        ctx.shapes = [tuple(t.shape) for t in (logits, lines, d1, d2, d3, depth, outs["pred_seg"])]
        return logits, lines, d1.unsqueeze(1), d2.unsqueeze(1), d3.unsqueeze(1), depth, outs["pred_seg"]

    @staticmethod
    def backward(ctx, g_logits, g_lines, g1, g2, g3, g_depth, g_seg):
        module = ctx.module
        tr = module.trainer()
        dev = ctx.depth.device
        def dense(g, shp, keep_layout=False):      # outputs no criterion touched arrive as None
            if g is None:
                return torch.zeros(shp, dtype=torch.float32, device=dev)
            return g.float() if keep_layout else g.float().reshape(shp).contiguous()
        shp = ctx.shapes
        g_logits, g_lines = dense(g_logits, shp[0]), dense(g_lines, shp[1])
        g1, g2, g3, g_depth = dense(g1, shp[2]), dense(g2, shp[3]), dense(g3, shp[4]), dense(g_depth, shp[5])
        g_seg = dense(g_seg, shp[6], keep_layout=True)        # [B,2,H,W]; cotangent_rows makes it channels-last
        with torch.no_grad(), torch.cuda.device(dev):
            rows_d, rows_s = tr.dense.tail.head.cotangent_rows(ctx.depth, g_depth, g_seg)
            tr.backward_dense(g1, g2, g3, rows_d, rows_s)
            tr.backward_line(g_logits, g_lines)
            grads = tr.grads()
            out = tuple(grads[n].clone() for n, _ in module.__dict__["_live"])
        return (None, None, None, None) + out


class GlassRGBD(_Node):
    """Drop-in for src/models/glassrgbd.py:44-123 (flag set --with_line --with_center --with_dense)."""

    def __init__(self, args):
        super().__init__()
        if not (getattr(args, "with_line", True) and getattr(args, "with_dense", True)):
            raise NotImplementedError("only the --with_line --with_dense configuration of the reference is a working "
                                      "model (SURVEY.md section 9-F); it is the one built here")
        if getattr(args, "with_line_depth", False):
            raise NotImplementedError("--with_line_depth raises AttributeError in the reference itself (SURVEY.md 9-F)")
        self.args = args
        self.num_queries = args.num_queries
        self.aux_loss = getattr(args, "aux_loss", True)
        self.cfg = dict(
            hidden_dim=args.hidden_dim, nheads=args.nheads, enc_layers=args.enc_layers, dec_layers=args.dec_layers,
            num_queries=args.num_queries, dense_trans_dim=args.dense_trans_dim, dense_trans_heads=args.dense_trans_heads,
            dense_trans_layers=tuple(args.dense_trans_layers), class_trans_layers=tuple(args.class_trans_layers),
            class_token_dim=args.class_token_dim, num_ref=args.num_ref, with_dense_center=bool(args.with_dense_center),
            window=7, interval_sample_num=tuple(args.interval_sample_num)[:2], depth_interval=tuple(args.depth_interval),
            min_depth_eval=args.min_depth_eval, max_depth_eval=args.max_depth_eval, max_depth=float(args.max_depth),
            aux_loss=self.aux_loss)
        hp = dict(hidden_dim=args.hidden_dim, dim_feedforward=args.dim_feedforward, enc_layers=args.enc_layers,
                  dec_layers=args.dec_layers, num_queries=args.num_queries, with_center=bool(args.with_center),
                  dense_trans_dim=args.dense_trans_dim, dense_trans_heads=args.dense_trans_heads,
                  dense_trans_layers=tuple(args.dense_trans_layers), class_trans_layers=tuple(args.class_trans_layers),
                  class_token_dim=args.class_token_dim, window=7, interval_sample_num=tuple(args.interval_sample_num))
        _build_tree(self, model_spec(hp))
        _init_parameters(self)
        self._plan = None
        self._plan_key = None
        # replay the forward as one CUDA graph per input shape (GWD_CUDA_GRAPH=0 launches kernel by kernel)
        import os
        self.use_cuda_graph = os.environ.get("GWD_CUDA_GRAPH", "1") != "0"

    # the kernel plan caches re-laid-out weights; rebuild it whenever a parameter changed or moved
    def _current_key(self):
        vers = tuple(t._version for t in self.parameters()) + tuple(t._version for t in self.buffers())
        dev = next(self.parameters()).device
        return (vers, str(dev))

    def plan(self):
        key = self._current_key()
        if self._plan is None or key != self._plan_key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("GlassRGBD runs on sm_100a CUDA kernels only; move the model to a CUDA device "
                                   "(there is no CPU path)")
            self._plan = _engine.Engine(self.state_dict(), self.cfg, device=dev)
            self._plan_key = key
        return self._plan

    def forward(self, samples, reflc_points=None, reflc_mat=None, img_name=None, _pinned=None, _trace=None, _static=False, _slot=0):
        if isinstance(samples, torch.Tensor) and samples.dim() == 4:
            images, mask = samples, None      # one equal-size batch: no padding mask to build (or to synchronise on)
        else:
            if isinstance(samples, (list, torch.Tensor)):
                samples = nested_tensor_from_tensor_list(samples)
            images, mask = samples.decompose()
            assert mask is not None
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(samples, images, mask, _pinned)
        plan = self.plan()      # raises off-GPU: there is no CPU path
        padded = getattr(samples, "padded", None)      # set on the host by nested_tensor_from_tensor_list (no sync)
        if mask is not None and padded is None:
            padded = bool(mask.any())                  # a hand-made NestedTensor: one host read of the mask
        with torch.cuda.device(images.device):
            x = images.float().contiguous()
            m = mask.to(x.device) if padded else None  # ragged batch: per-image position codes + key-padding masks
            if self.use_cuda_graph and _pinned is None and _trace is None:
                out = plan.forward_graphed(x, mask=m, slot=_slot)
                # the graph's static outputs are overwritten by the next call with the same shape: hand out copies (59 MB at
                # 16 x 480 x 640) unless the caller asked for the static tensors (infer_stream copies them itself)
                return out if _static else _clone_outputs(out)
            return plan.forward(x, pinned=_pinned, trace=_trace, mask=m)

    def invalidate_plan(self):
        """drop the cached kernel plan (re-laid-out weights, CUDA graphs): call after writing parameters through raw pointers
        (in-place writes through torch bump the version counters the cache is keyed on; kernels do not)"""
        self._plan = None
        self._plan_key = None

    # ------------------------------------------------------------------ training (src/engine_glassrgbd.py:45-166)
    def trainer(self, **optim):
        """the whole-model training engine (train_model.Trainer) behind this module, built on first use from the module's
        parameters.  Two ways to train:
          * drop-in: `out = model(samples)` under model.train() returns tensors with an autograd edge into the engine, so the
            reference's loop (criteria on `out`, `losses.backward()`, clip_grad_norm_, optimizer.step()) runs unmodified; the
            engine's flat buffers are re-loaded from the nn.Parameters whenever an optimizer has changed them;
          * fused: `model.trainer(lr=..., ...).train_step(images, targets, depth_gt, seg_gt, criterion)` does forward, losses,
            backward, gradient exchange, clip and AdamW on the flat buffers (call `model.sync_from_trainer()` before
            `state_dict()` / evaluation)."""
        from .train_model import Trainer
        if self.__dict__.get("_trainer") is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("GlassRGBD trains on sm_100a CUDA kernels only; move the model to a CUDA device")
            a = self.args
            kw = dict(lr=getattr(a, "lr", 1e-4), lr_backbone=getattr(a, "lr_backbone", 1e-5), weight_decay=getattr(a, "weight_decay", 1e-4),
                      max_norm=getattr(a, "clip_max_norm", 0.1),
                      depth_loss_weights=tuple(getattr(a, "depth_loss_weights", (0.25, 0.25, 0.25, 1.0))),
                      seg_loss_weight=float(getattr(a, "seg_loss_weight", 2.0)))
            kw.update(optim)
            cfg = dict(self.cfg, log_depth_error=bool(getattr(a, "log_depth_error", False)),
                       variance_focus=float(getattr(a, "variance_focus", 0.85)), dropout=float(getattr(a, "dropout", 0.0) or 0.0))
            self.__dict__["_trainer"] = Trainer(self.state_dict(), cfg, device=dev, **kw)
            self.__dict__["_trainer_versions"] = self._param_versions()
            live = set(self.__dict__["_trainer"].state_dict())
            self.__dict__["_live"] = [(n, p) for n, p in self.named_parameters() if n in live and p.requires_grad]
        return self.__dict__["_trainer"]

    def _param_versions(self):
        return tuple(p._version for p in self.parameters())

    def sync_from_trainer(self):
        """copy the engine's master parameters back into the nn.Parameters (after fused training steps)"""
        tr = self.__dict__.get("_trainer")
        if tr is not None:
            sd = tr.state_dict()
            with torch.no_grad():
                for n, p in self.named_parameters():
                    if n in sd:
                        p.copy_(sd[n])
            self.__dict__["_trainer_versions"] = self._param_versions()
            self.invalidate_plan()

    def _forward_train(self, samples, images, mask, pinned):
        padded = getattr(samples, "padded", None)
        if mask is not None and padded is None:
            padded = bool(mask.any())
        if padded and (images.shape[-2] % 32 or images.shape[-1] % 32):
            raise NotImplementedError("training on a ragged batch needs the batch padded to a multiple of 32: collate with "
                                      "nested_tensor_from_tensor_list(..., size_divisibility=32)")
        tr = self.trainer()
        vers = self._param_versions()
        if vers != self.__dict__["_trainer_versions"]:        # an optimizer (or load_state_dict) changed the nn.Parameters
            tr.load_params(self.state_dict())
            self.__dict__["_trainer_versions"] = vers
        tr.exchange_grads = False          # the caller's DistributedDataParallel (or nobody) reduces the .grad tensors
        live = self.__dict__["_live"]
        with torch.cuda.device(images.device):
            res = _TrainStep.apply(self, images.float().contiguous(), pinned, mask.to(images.device) if padded else None,
                                   *[p for _, p in live])
        logits, lines, d1, d2, d3, depth, seg = res
        out = {"pred_logits": logits[-1], "pred_lines": lines[-1]}
        if self.cfg["aux_loss"]:
            out["aux_outputs"] = [{"pred_logits": a, "pred_lines": b} for a, b in zip(logits[:-1], lines[:-1])]
        out["pred_depth"] = [d1, d2, d3, depth]
        out["pred_seg"] = seg
        return out

    @torch.no_grad()
    def infer_stream(self, host_batches, keys=("pred_logits", "pred_lines", "pred_depth", "pred_seg")):
        """Serving loop over an iterable of PINNED host batches: yields, per batch, a dict of pinned host tensors
        (`pred_depth` = the full-resolution map).  A batch is either fp32 [B,3,H,W] (already normalised, what the
        reference's data loader hands over) or uint8 [B,H,W,3] raw images: those cross PCIe at a quarter of the bytes and
        are normalised on the GPU (gwd_images_to_batch, bit-identical to ToTensor + Normalize of src/datasets/coco.py:77-78).  Uploads and downloads run on
        their own streams with three buffers each, and consecutive batches replay three CUDA-graph instances on three compute
        streams: the copy of batch i+1 and the read-back of batch i-1 overlap the forward of batch i, and the latency-bound
        phases of one forward (DETR chain, coarse Swin stages) share the SMs with the wide phases of the next (+12 % images/s
        at 16 x 480 x 640).  A yielded dict is valid until the next one is requested."""
        plan = self.plan()
        dev = plan.dev
        NBUF = 3        # batches in flight: one uploading / computing, one computing, one being read back
        caller = torch.cuda.current_stream(dev)
        # streams, events and the device / pinned-host buffers live on the module: pinned allocations cost milliseconds
        st = self.__dict__.setdefault("_serve_state", {})
        if st.get("dev") != dev:
            st.clear()
            st.update(dev=dev, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev), s_comp=[torch.cuda.Stream(dev) for _ in range(NBUF)],
                      x_dev=[None] * NBUF, u_dev=[None] * NBUF, u_tab=[None] * NBUF, out_dev=[None] * NBUF, out_host=[None] * NBUF,
                      ev=[[torch.cuda.Event() for _ in range(NBUF)] for _ in range(4)])
        s_in, s_out = st["s_in"], st["s_out"]
        x_dev, out_dev, out_host = st["x_dev"], st["out_dev"], st["out_host"]
        ev_in, ev_used, ev_out, ev_done = st["ev"]
        torch.cuda.synchronize(dev)      # a previous, abandoned iteration may still own the buffers
        pending = []
        two = self.use_cuda_graph and os.environ.get("GWD_SERVE_STREAMS", "2") != "1"

        def pick(out):
            res = {}
            for k in keys:
                v = out[k]
                res[k] = (v[-1] if k == "pred_depth" else v).contiguous()
            return res

        for i, hb in enumerate(host_batches):
            b = i % NBUF
            comp = st["s_comp"][b] if two else caller
            raw = hb.dtype == torch.uint8
            u_dev, u_tab = st["u_dev"], st["u_tab"]
            with torch.cuda.stream(s_in):
                stage = u_dev if raw else x_dev
                if stage[b] is None or stage[b].shape != hb.shape:
                    stage[b] = torch.empty(hb.shape, dtype=hb.dtype if raw else torch.float32, device=dev)
                    u_tab[b] = None
                elif i >= NBUF:
                    s_in.wait_event(ev_used[b])          # the forward of batch i-NBUF has consumed this buffer
                stage[b].copy_(hb, non_blocking=True)
                ev_in[b].record(s_in)
            with torch.cuda.stream(comp):
                comp.wait_event(ev_in[b])
                if raw:     # uint8 HWC -> normalised fp32 NCHW on the compute stream
                    B_, H_, W_ = hb.shape[:3]
                    if x_dev[b] is None or x_dev[b].shape != (B_, 3, H_, W_):
                        x_dev[b] = torch.empty(B_, 3, H_, W_, dtype=torch.float32, device=dev)
                    ops.images_to_batch(u_dev[b], out=x_dev[b], want_mask=False, table=u_tab[b])
                    u_tab[b] = ops.images_to_batch.last_table
                out = pick(self.forward(x_dev[b], _static=True, _slot=b if two else 0))
                ev_used[b].record(comp)
                if out_dev[b] is None or set(out_dev[b]) != set(out) or any(out_dev[b][k].shape != v.shape for k, v in out.items()):
                    out_dev[b] = {k: torch.empty_like(v) for k, v in out.items()}
                    out_host[b] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
                elif i >= NBUF:
                    comp.wait_event(ev_done[b])              # read-back of batch i-NBUF has left this buffer
                for k, v in out.items():
                    out_dev[b][k].copy_(v, non_blocking=True)    # the graph's static outputs are reused by the next replay
                ev_out[b].record(comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_out[b])
                for k, v in out_dev[b].items():
                    out_host[b][k].copy_(v, non_blocking=True)
                ev_done[b].record(s_out)
            pending.append(b)
            if len(pending) == NBUF:
                pb = pending.pop(0)
                ev_done[pb].synchronize()
                yield out_host[pb]
        for pb in pending:
            ev_done[pb].synchronize()
            yield out_host[pb]


# --------------------------------------------------------------------------------------------------
# criteria
# --------------------------------------------------------------------------------------------------
class HungarianMatcher_Line(nn.Module):
    """src/models/matcher.py:10-82.  The block-diagonal cost matrix comes from the gwd_match_cost kernel; the
    assignment itself is solved by scipy.optimize.linear_sum_assignment exactly like the reference (:74)."""

    def __init__(self, cost_class=1, cost_line=1):
        super().__init__()
        assert cost_class != 0 or cost_line != 0, "all costs cant be 0"
        self.cost_class, self.cost_line = cost_class, cost_line

    @torch.no_grad()
    def cost_matrices(self, outputs, targets):
        logits = outputs["pred_logits"].float().contiguous()
        lines = outputs["pred_lines"].float().contiguous()
        B, Q = logits.shape[:2]
        sizes = [len(v["lines"]) for v in targets]
        tgt_lines = torch.cat([v["lines"] for v in targets]).float().contiguous()
        tgt_ids = torch.cat([v["labels"] for v in targets]).to(torch.int64).contiguous()
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        offsets = torch.tensor(offs, dtype=torch.int32, device=logits.device)
        cost, _ = ops.match_cost(logits, lines, tgt_lines, tgt_ids, offsets, float(self.cost_class), float(self.cost_line))
        flat = cost.cpu()
        return [flat[offs[b] * Q: offs[b + 1] * Q].view(Q, sizes[b]) for b in range(B)]

    @torch.no_grad()
    def forward(self, outputs, targets):
        from scipy.optimize import linear_sum_assignment
        result = []
        for c in self.cost_matrices(outputs, targets):
            i, j = linear_sum_assignment(c.numpy())
            result.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        return result

    @torch.no_grad()
    def pairs_from_raw(self):
        """the last stacked assignment in the reference's format: list over stages of per-image (idx_pred, idx_tgt)"""
        qi, ti, cnt, S, B = self.last_raw
        return [[(torch.from_numpy(qi[s * B + b, :cnt[s * B + b]].astype("int64")), torch.from_numpy(ti[s * B + b, :cnt[s * B + b]].astype("int64")))
                 for b in range(B)] for s in range(S)]

    @torch.no_grad()
    def stacked_cost(self, logits, lines, targets):
        """first half of `forward_stacked`: ONE gwd_match_cost launch over the S*B (stage, image) pairs and an ASYNCHRONOUS
        copy of the costs into pinned host memory.  Returns a handle for `stacked_solve`; the caller may enqueue more GPU work
        (e.g. the dense branch, which does not depend on the matching) before it solves."""
        S, B, Q = logits.shape[:3]
        sizes = [len(v["lines"]) for v in targets]
        tgt_lines = torch.cat([v["lines"] for v in targets]).float()
        tgt_ids = torch.cat([v["labels"] for v in targets]).to(torch.int64)
        offs = [0]
        for _ in range(S):
            for n in sizes:
                offs.append(offs[-1] + n)
        offsets = torch.tensor(offs, dtype=torch.int32).to(logits.device, non_blocking=True)
        cost, _ = ops.match_cost(logits.float().reshape(S * B, Q, -1).contiguous(), lines.float().reshape(S * B, Q, -1).contiguous(),
                                 tgt_lines.repeat(S, 1).contiguous(), tgt_ids.repeat(S).contiguous(), offsets,
                                 float(self.cost_class), float(self.cost_line))
        host = getattr(self, "_host_cost", None)
        if host is None or host.numel() < cost.numel():
            host = self._host_cost = torch.empty(max(cost.numel(), 1), dtype=torch.float32).pin_memory()
        host[:cost.numel()].copy_(cost, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return dict(host=host, n=cost.numel(), event=ev, offs=offs, sizes=sizes, S=S, B=B, Q=Q, keep=cost)

    @torch.no_grad()
    def stacked_solve(self, pending, want_pairs=True):
        """second half of `forward_stacked`: wait for the cost copy, solve the S*B assignments on the host"""
        from scipy.optimize import linear_sum_assignment
        import numpy as np
        pending["event"].synchronize()
        S, B, Q, offs, sizes = pending["S"], pending["B"], pending["Q"], pending["offs"], pending["sizes"]
        flat = pending["host"][:pending["n"]].numpy()
        total = sum(sizes)
        starts = np.concatenate([[0], np.cumsum(sizes)])
        if os.environ.get("GWD_LSAP", "native") == "scipy":     # the reference's solver, problem by problem
            pairs = [linear_sum_assignment(flat[offs[p] * Q:(offs[p] + sizes[p % B]) * Q].reshape(Q, sizes[p % B])) for p in range(S * B)]
            qi = np.zeros((S * B, max(Q, 1)), dtype=np.int32)
            ti = np.zeros_like(qi)
            cnt = np.array([len(q) for q, _ in pairs], dtype=np.int32)
            for p, (q, t) in enumerate(pairs):
                qi[p, :len(q)], ti[p, :len(q)] = q, t
        else:       # same algorithm and tie rules, all S*B problems on the host cores at once (tests/test_lsap_cpu.py)
            qi, ti, cnt = ops.lsap_batch(flat, [o * Q for o in offs[:-1]], sizes * S, Q, n_threads=_lsap_threads(), raw=True)
        # the assignment as int32 columns (stage, image, query, row of the concatenated targets) for gwd_set_loss, built
        # without a Python loop over the S*B problems
        p_idx, k_idx = np.nonzero(np.arange(qi.shape[1])[None, :] < cnt[:, None])
        b_idx = p_idx % B
        match = np.stack([p_idx // B, b_idx, qi[p_idx, k_idx], ti[p_idx, k_idx] + starts[b_idx]]).astype(np.int32)
        per_stage = cnt.reshape(S, B).sum(1)
        self.last_match = (np.ascontiguousarray(match), np.concatenate([[0], np.cumsum(per_stage)]).astype(np.int32))
        self.last_raw = (qi, ti, cnt, S, B)
        if not want_pairs:
            return None
        result = self.pairs_from_raw()
        assert offs[-1] == S * total
        return result

    @torch.no_grad()
    def set_pairs(self, pairs, targets):
        """install a given stacked assignment (list over stages of per-image (idx_pred, idx_tgt)) as if `stacked_solve` had
        found it: parity tests pin the discrete matching to the oracle's, exactly like the line / sample selections"""
        import numpy as np
        S, B = len(pairs), len(pairs[0])
        sizes = [len(v["lines"]) for v in targets]
        starts = np.concatenate([[0], np.cumsum(sizes)])
        cols, per_stage = [], []
        for s_, stage in enumerate(pairs):
            n = 0
            for b, (qi, ti) in enumerate(stage):
                qi, ti = np.asarray(qi, dtype=np.int64), np.asarray(ti, dtype=np.int64)
                cols.append(np.stack([np.full_like(qi, s_), np.full_like(qi, b), qi, ti + starts[b]]))
                n += len(qi)
            per_stage.append(n)
        match = np.concatenate(cols, axis=1).astype(np.int32) if cols else np.zeros((4, 0), np.int32)
        self.last_match = (np.ascontiguousarray(match), np.concatenate([[0], np.cumsum(per_stage)]).astype(np.int32))
        self.last_raw = None

    @torch.no_grad()
    def forward_stacked(self, logits, lines, targets, want_pairs=True):
        """All S decoder stages at once (the reference calls the matcher once per stage, src/models/glassrgbd.py:318,344):
        logits [S,B,Q,C], lines [S,B,Q,D] -> list over stages of the per-image (idx_pred, idx_tgt) lists.  ONE
        gwd_match_cost launch over S*B (stage, image) pairs, ONE device-to-host copy, then the S*B assignments."""
        return self.stacked_solve(self.stacked_cost(logits, lines, targets), want_pairs)


def _lsap_threads():
    """host threads for gwd_lsap_batch: the box's cores divided by the ranks that share them (8 ranks solving on
    hardware_concurrency() threads each oversubscribe the cores: 1.4 -> 2.7 ms per step in SCALE_r01); GWD_LSAP_THREADS overrides"""
    env = os.environ.get("GWD_LSAP_THREADS")
    if env:
        return max(1, int(env))
    local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, min(16, (os.cpu_count() or 1) // max(local, 1)))


def build_matcher(args, type=None):
    return HungarianMatcher_Line(cost_class=args.set_cost_class, cost_line=args.set_cost_line)


def _world_size():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class SetCriterion(nn.Module):
    """line set loss, src/models/glassrgbd.py:133-358 (losses 'lines_labels' and 'lines', repeated on the aux outputs)"""

    def __init__(self, num_classes, weight_dict, eos_coef, losses, args, matcher=None):
        super().__init__()
        self.num_classes, self.matcher, self.weight_dict = num_classes, matcher, weight_dict
        self.eos_coef, self.losses, self.args = eos_coef, losses, args
        w = torch.ones(num_classes + 1)
        w[-1] = eos_coef
        self.register_buffer("empty_weight", w)
        if getattr(args, "label_loss_func", "cross_entropy") != "cross_entropy":
            raise NotImplementedError("only label_loss_func='cross_entropy' (the default) is built")

    def _pairs(self, indices):
        batch_idx = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        return batch_idx, torch.cat([src for src, _ in indices])

    def loss_lines_labels(self, outputs, targets, num_items, origin_indices):
        logits = outputs["pred_logits"]
        idx = self._pairs(origin_indices)
        matched = torch.cat([t["labels"][j.to(t["labels"].device)] for t, (_, j) in zip(targets, origin_indices)])
        classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        classes[idx[0].to(logits.device), idx[1].to(logits.device)] = matched.to(logits.device)
        return {"loss_ce": F.cross_entropy(logits.transpose(1, 2), classes, self.empty_weight.to(logits.device))}

    def loss_lines(self, outputs, targets, num_items, origin_indices):
        idx = self._pairs(origin_indices)
        dev = outputs["pred_lines"].device
        src = outputs["pred_lines"][idx[0].to(dev), idx[1].to(dev)]
        tgt = torch.cat([t["lines"][j.to(t["lines"].device)] for t, (_, j) in zip(targets, origin_indices)], dim=0).to(dev)
        return {"loss_line": F.l1_loss(src, tgt, reduction="none").sum() / num_items}

    def get_loss(self, loss, outputs, targets, num_items, **kw):
        table = {"lines_labels": self.loss_lines_labels, "lines": self.loss_lines,
                 "POST_lines_labels": self.loss_lines_labels, "POST_lines": self.loss_lines}
        assert loss in table, f"do you really want to compute {loss} loss?"
        return table[loss](outputs, targets, num_items, **kw)

    def forward_stacked(self, logits, lines, targets):
        """SetCriterion.forward for all decoder stages at once: logits [S,B,Q,2], lines [S,B,Q,D] with the FINAL stage
        last (the layout the line heads produce).  Same losses and key names as `forward` (stage S-1 -> "loss_ce",
        "loss_line"; stage i < S-1 -> "_i"), computed with one stacked matching (HungarianMatcher_Line.forward_stacked),
        one host-to-device copy of the assignment indices and a dozen batched torch ops instead of ~60 per stage."""
        S, B, Q = logits.shape[:3]
        dev = logits.device
        indices = self.matcher.forward_stacked(logits, lines, targets)
        n = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float, device=dev)
        if _world_size() > 1:
            torch.distributed.all_reduce(n)
        num_items = torch.clamp(n / _world_size(), min=1)
        sizes = [len(t["labels"]) for t in targets]
        starts = [0]
        for m in sizes:
            starts.append(starts[-1] + m)
        s_idx = torch.cat([torch.full_like(i, s) for s, stage in enumerate(indices) for i, _ in stage])
        b_idx = torch.cat([torch.full_like(i, b) for stage in indices for b, (i, _) in enumerate(stage)])
        q_idx = torch.cat([i for stage in indices for i, _ in stage])
        t_idx = torch.cat([j + starts[b] for stage in indices for b, (_, j) in enumerate(stage)])
        idx = torch.stack([s_idx, b_idx, q_idx, t_idx]).to(dev, non_blocking=True)
        labels = torch.cat([t["labels"] for t in targets]).to(dev)
        tgt_lines = torch.cat([t["lines"] for t in targets]).to(dev)
        classes = torch.full((S, B, Q), self.num_classes, dtype=torch.int64, device=dev)
        classes[idx[0], idx[1], idx[2]] = labels[idx[3]]
        w = self.empty_weight.to(dev)[classes]
        nll = -torch.log_softmax(logits, dim=-1).gather(-1, classes.unsqueeze(-1)).squeeze(-1)
        loss_ce = (w * nll).sum((1, 2)) / w.sum((1, 2))
        l1 = (lines[idx[0], idx[1], idx[2]] - tgt_lines[idx[3]]).abs().sum(-1)
        loss_line = torch.zeros(S, dtype=l1.dtype, device=dev).index_add_(0, idx[0], l1) / num_items
        losses = {}
        for s in range(S):
            suffix = "" if s == S - 1 else "_%d" % s
            if "lines_labels" in self.losses:
                losses["loss_ce" + suffix] = loss_ce[s]
            if "lines" in self.losses:
                losses["loss_line" + suffix] = loss_line[s]
        self.last_indices = indices
        return losses

    @torch.no_grad()
    def forward_backward_stacked(self, logits, lines, targets, pending=None, pinned_pairs=None):
        """forward_stacked AND its gradient in one gwd_set_loss launch: -> (losses dict, dlogits, dlines) where the
        gradients are those of sum_k weight_dict[k] * losses[k].  No autograd graph, no per-loss torch kernels: the host
        only solves the assignments and uploads them (one int32 [4, M] copy)."""
        S, B, Q = logits.shape[:3]
        dev = logits.device
        # `pending`: a handle of matcher.stacked_cost issued earlier (the host solve then overlaps whatever GPU work the caller
        # enqueued in between)
        if pinned_pairs is not None:
            self.matcher.set_pairs(pinned_pairs, targets)
        else:
            self.matcher.stacked_solve(pending if pending is not None else self.matcher.stacked_cost(logits, lines, targets), want_pairs=False)
        n = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float, device=dev)
        if _world_size() > 1:
            torch.distributed.all_reduce(n)
        num_items = torch.clamp(n / _world_size(), min=1)
        match_np, soff_np = self.matcher.last_match
        match = torch.from_numpy(match_np).to(dev, non_blocking=True)
        soff = torch.from_numpy(soff_np).to(dev, non_blocking=True)
        key = [("loss_ce" + ("" if s_ == S - 1 else "_%d" % s_), "loss_line" + ("" if s_ == S - 1 else "_%d" % s_)) for s_ in range(S)]
        wd = self.weight_dict
        cache = getattr(self, "_stage_weights", None)
        if cache is None or cache[0] != (S, str(dev)):
            w_ce = torch.tensor([float(wd.get(a, 0.0)) if "lines_labels" in self.losses else 0.0 for a, _ in key]).to(dev)
            w_line = torch.tensor([float(wd.get(b, 0.0)) if "lines" in self.losses else 0.0 for _, b in key]).to(dev)
            cache = self._stage_weights = ((S, str(dev)), w_ce, w_line, self.empty_weight.to(dev).float().contiguous())
        _, w_ce, w_line, class_w = cache
        tgt_lines = torch.cat([t["lines"] for t in targets]).float().contiguous()
        tgt_labels = torch.cat([t["labels"] for t in targets]).to(torch.int64).contiguous()
        vals, dlogits, dlines = ops.set_loss(logits.float().contiguous(), lines.float().contiguous(), tgt_lines, tgt_labels, match,
                                             soff, class_w, w_ce, w_line, num_items)
        self.last_total = (vals[:, 0] * w_ce + vals[:, 1] * w_line).sum()      # sum_k weight_dict[k] * losses[k]
        losses = {}
        for s_, (a, b) in enumerate(key):
            if "lines_labels" in self.losses:
                losses[a] = vals[s_, 0]
            if "lines" in self.losses:
                losses[b] = vals[s_, 1]
        self.last_indices = None        # built on demand from the matcher's raw arrays (`indices_of_last_call`)
        return losses, dlogits, dlines

    def indices_of_last_call(self):
        """assignments of the last stacked call, list over stages of per-image (idx_pred, idx_tgt)"""
        return self.last_indices if self.last_indices is not None else self.matcher.pairs_from_raw()

    def forward(self, outputs, targets, origin_indices=None, depth_gt=None):
        plain = {k: v for k, v in outputs.items() if k != "aux_outputs"}
        origin_indices = self.matcher(plain, targets)
        n = torch.as_tensor([sum(len(t["labels"]) for t in targets)], dtype=torch.float,
                            device=next(iter(outputs.values())).device)
        if _world_size() > 1:
            torch.distributed.all_reduce(n)
        num_items = torch.clamp(n / _world_size(), min=1).item()
        losses = {}
        for loss in self.losses:
            losses.update(self.get_loss(loss, outputs, targets, num_items, origin_indices=origin_indices))
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            idx = self.matcher(aux, targets)
            for loss in self.losses:
                d = self.get_loss(loss, aux, targets, num_items, origin_indices=idx)
                losses.update({k + f"_{i}": v for k, v in d.items()})
        return losses


class SilogLoss(nn.Module):
    """src/models/glassrgbd.py:360-374, without the boolean-index compaction (masked sums instead: no host sync)"""

    def __init__(self, variance_focus=0.85, log_depth_error=True):
        super().__init__()
        self.variance_focus, self.log_depth_error = variance_focus, log_depth_error

    def forward(self, depth_est, depth_gt, mask):
        m = mask.to(depth_est.dtype)
        n = m.sum()
        est = torch.where(mask, depth_est, torch.ones_like(depth_est))
        gt = torch.where(mask, depth_gt, torch.ones_like(depth_gt))
        d = (torch.log(est) - torch.log(gt)) if self.log_depth_error else ((est + torch.log(est)) - (gt + torch.log(gt)))
        d = d * m
        return torch.sqrt((d ** 2).sum() / n - self.variance_focus * (d.sum() / n) ** 2) * 10.0


class SegLoss(nn.Module):
    """src/models/glassrgbd.py:376-383"""

    def forward(self, seg_pred, seg_gt):
        return F.cross_entropy(seg_pred, seg_gt)


class PostProcess_Line(nn.Module):
    """src/models/glassrgbd.py:452-506"""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, output_type):
        if output_type in ("prediction", "prediction_POST"):
            logits = outputs["pred_logits"]
            lines = outputs["pred_lines" if output_type == "prediction" else "POST_pred_lines"]
            assert len(logits) == len(target_sizes) and target_sizes.shape[1] == 2
            scores, labels = F.softmax(logits, -1)[..., :-1].max(-1)
            h, w = target_sizes.unbind(1)
            scale = torch.stack([w, h, w, h], dim=1)[:, None, :]
            return [{"scores": s, "labels": l, "lines": b} for s, l, b in zip(scores, labels, lines * scale)]
        if output_type == "ground_truth":
            h, w = target_sizes.unbind(1)
            scale = torch.stack([w, h, w, h], dim=1)
            return [{"labels": d["labels"], "lines": d["lines"] * scale, "image_id": d["image_id"]} for d in outputs]
        raise AssertionError(output_type)


def build(args):
    """src/models/glassrgbd.py:509-579 -> (model, [criterion, criterion_depth, criterion_seg, criterion_plane], postprocessors)"""
    model = GlassRGBD(args)
    weight_dict = {"loss_ce": 1, "loss_line": args.line_loss_coef}
    if args.aux_loss:
        for i in range(args.dec_layers - 1):
            weight_dict.update({k + f"_{i}": v for k, v in list(weight_dict.items())[:2]})
    device = torch.device(args.device)
    criterion = SetCriterion(1, weight_dict=weight_dict, eos_coef=args.eos_coef, losses=["lines_labels", "lines"], args=args,
                             matcher=build_matcher(args, type="origin_line")).to(device)
    criterion_depth = SilogLoss(variance_focus=args.variance_focus, log_depth_error=args.log_depth_error).to(device)
    criterion_seg = SegLoss().to(device)
    if getattr(args, "with_plane_norm_loss", False):
        raise NotImplementedError("--with_plane_norm_loss needs matplotlib polygons and B == 1 in the reference; not built")
    return model, [criterion, criterion_depth, criterion_seg, None], {"line": PostProcess_Line()}


def build_model(args):
    return build(args)


def default_args(**overrides):
    """the reference's argparse defaults for the flags that shape the model (src/args.py), with the working flag set
    --with_line --with_center --with_dense --num_queries 100"""
    import argparse
    d = dict(device="cuda", hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6, dim_feedforward=2048, dropout=0.1,
             num_queries=100, aux_loss=True, with_line=True, with_dense=True, with_center=True, with_dense_center=False,
             with_line_depth=False, with_plane_norm_loss=False, set_cost_class=1.0, set_cost_line=5.0, line_loss_coef=5.0,
             eos_coef=0.1, label_loss_func="cross_entropy", variance_focus=0.85, log_depth_error=False, max_depth=10,
             min_depth_eval=1e-3, max_depth_eval=10.0, dense_trans_dim=512, dense_trans_heads=16, dense_trans_layers=[4],
             class_trans_layers=[2, 2, 1], class_token_dim=64, num_ref=20, interval_sample_num=[30, 80, 160],
             depth_interval=[0.1, 0.3, 0.5, 0.7, 0.9], depth_loss_weights=[0.25, 0.25, 0.25, 1], seg_loss_weight=2.0,
             lr_backbone=1e-5, backbone="resnet50", layer1_num=3)
    d.update(overrides)
    return argparse.Namespace(**d)
