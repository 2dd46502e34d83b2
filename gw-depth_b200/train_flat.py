"""Flat-buffer parameter storage shared by the dense-branch training modules (train_dense.DenseHead, train_pyramid.Pyramid).

Every trainable tensor of a module lives in ONE flat fp32 master buffer in the PHYSICAL layout the tcgen05 GEMM reads
(Linear / 1x1 conv [n_pad, cin_pad], 3x3 conv [tap = dx*3+dy][n_pad][cin_pad], channel counts padded to 16 with exact
zeros, optional input-channel maps for inputs that are padded concats) with flat gradient / Adam-moment twins and a bf16
mirror.  Forward weights are views of the mirror, the weight-gradient kernels accumulate into views of the gradient
buffer, so the optimizer (reference: AdamW + clip_grad_norm_, src/main_glassrgbd.py:59-67, src/engine_glassrgbd.py:155-159)
is two launches (gwd_sumsq, gwd_adamw_step) and data-parallel training one all-reduce per module.  Padding rows / columns
have zero gradient and zero value, so they stay zero under AdamW.
"""
import os

import torch

from . import ops, parallel
from .ops import RES_AFTER, RES_NONE, PackedWeight, conv_gemm, round_up


# Weight gradients are off the critical path of a backward pass (nothing reads them before the optimizer), so they run on a SIDE
# stream: in the captured training step they become parallel graph branches next to the data-gradient chain, which matters for
# the many launches that do not fill 148 SMs (1/32 ... 1/8 scale maps at batch 8).  Operands are kept alive until the join.
_SIDE = {}
_KEEP = []
FORK_WGRAD = True


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=key)
    return _SIDE[key]


def fork_wgrad(fn, *operands):
    """run fn() (a weight-gradient launch reading `operands`) on the side stream, ordered after everything enqueued so far"""
    if not FORK_WGRAD:
        return fn()
    main = torch.cuda.current_stream()
    side = _side_stream(operands[0].device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fn()
    _KEEP.extend(operands)


def join_wgrads(dev=None):
    """the main stream waits for every forked weight gradient (call at the end of a module's backward, before the flat
    gradient buffer is read or the operands' memory is reused)"""
    if _KEEP:
        torch.cuda.current_stream().wait_stream(_side_stream(_KEEP[0].device))
        del _KEEP[:]


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class Conv3x3:
    """one 3x3 convolution without bias: mirror / gradient views [9, n_pad, cin_pad] + the flipped, transposed mirror whose
    gwd_conv_gemm on dY is the data gradient"""

    def __init__(self, owner, name):
        self.name = name
        self.wb, self.gw = owner.view(owner.Wb, name), owner.view(owner.G, name)
        _, n_pad, c_pad = self.wb.shape
        n, c = owner.index[name][2][:2]
        self.n, self.c, self.n_pad, self.c_pad = n, c, n_pad, c_pad
        self.pw = PackedWeight(self.wb, None, 9, n, c_pad)
        self.wT = torch.empty(9, c_pad, n_pad, dtype=torch.bfloat16, device=self.wb.device)
        self.pwT = PackedWeight(self.wT, None, 9, c_pad, n_pad)

    def transposes(self):
        # data gradient of a stride-1 pad-1 conv = the conv of dY with the filter flipped in (dy, dx): tap 8 - t
        return [(self.wb[8 - t], self.wT[t]) for t in range(9)]


class Linear:
    """Linear / 1x1 conv (bias optional)"""

    def __init__(self, owner, wname, bname=None):
        self.wb, self.gw = owner.view(owner.Wb, wname), owner.view(owner.G, wname)
        self.gb = owner.view(owner.G, bname) if bname else None
        n_pad, k = self.wb.shape
        self.n, self.n_pad, self.k = owner.index[wname][2][0], n_pad, k
        self.pw = PackedWeight(self.wb.view(1, n_pad, k), owner.view(owner.P, bname) if bname else None, 1, self.n, k)
        self.wT = torch.empty(k, n_pad, dtype=torch.bfloat16, device=self.wb.device)
        self.pwT = PackedWeight(self.wT.view(1, k, n_pad), None, 1, k, n_pad)

    def transposes(self):
        return [(self.wb, self.wT)]


# GWD_FUSE_ACT_GRAD=1: the activation backward runs in the epilogue of the data-gradient GEMM that produces its operand
# (gwd_conv_gemm, res_mode = GWD_RES_MUL_ACTGRAD) instead of as separate gwd_act_bwd launches.  Correct (tests/test_gemm_gpu.py and the
# whole training suite pass with it), but the step time does not move (33.9 ms either way: the removed passes were overlapped, the
# epilogues get longer), so the separate launches stay the default.
FUSE_ACT_GRAD = os.environ.get("GWD_FUSE_ACT_GRAD", "0") == "1"


class FlatModule:
    def __init__(self, tensors, layout=None, device="cuda", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                 max_norm=0.1):
        """tensors: {short name: logical fp32 tensor}; layout: {short name: dict(cin_pad=..., col_map=LongTensor)} overrides"""
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("%s runs on libgwd_b200 CUDA kernels only (no CPU fallback)" % type(self).__name__)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.t = 0
        self.layout = layout or {}
        self.index, off = {}, 0
        for short, v in tensors.items():
            lay = self.layout.get(short, {})
            if v.dim() == 4 and v.shape[-1] == 3:
                phys = (9, round_up(v.shape[0], 16), lay.get("cin_pad") or round_up(v.shape[1], 16))
            elif v.dim() in (2, 4):
                phys = (round_up(v.shape[0], 16), lay.get("cin_pad") or round_up(v.shape[1], 16))
            else:
                phys = (round_up(v.shape[0], 16),)
            self.index[short] = (off, phys, tuple(v.shape))
            off += _numel(phys)
        self.numel = off
        self.P = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.G, self.M, self.V = torch.zeros_like(self.P), torch.zeros_like(self.P), torch.zeros_like(self.P)
        for short, v in tensors.items():
            self.view(self.P, short).copy_(self._to_physical(short, v.detach().to(self.dev, torch.float32)))
        self.Wb = self.P.to(torch.bfloat16)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._wt_tables = None
        # physical input columns a weight reads from a SHARED buffer but does not map (they belong to other consumers)
        self._unmapped = {}
        for short, lay in self.layout.items():
            if lay.get("shared_input") and lay.get("col_map") is not None:
                free = torch.ones(self.index[short][1][-1], dtype=torch.bool)
                free[lay["col_map"]] = False
                if bool(free.any()):
                    self._unmapped[short] = free.nonzero().flatten().to(self.dev)

    # ------------------------------------------------------------------ views, logical <-> physical layout
    def view(self, flat, short):
        o, phys, _ = self.index[short]
        return flat[o:o + _numel(phys)].view(phys)

    def _cols(self, short, c, device):
        cm = self.layout.get(short, {}).get("col_map")
        return (cm if cm is not None else torch.arange(c)).to(device)

    def _to_physical(self, short, v):
        _, phys, logical = self.index[short]
        out = torch.zeros(phys, dtype=torch.float32, device=v.device)
        if len(phys) == 3:          # [N, C, 3(dy), 3(dx)] -> [dx*3+dy][N][C]
            n, c = logical[:2]
            out.view(3, 3, phys[1], phys[2])[:, :, :n, self._cols(short, c, v.device)] = v.permute(3, 2, 0, 1)
        elif len(phys) == 2:
            n, c = logical[:2]
            out[:n, self._cols(short, c, v.device)] = v.reshape(n, c)
        else:
            out[: logical[0]] = v
        return out

    def _to_logical(self, short, p):
        _, phys, logical = self.index[short]
        if len(phys) == 3:
            n, c = logical[:2]
            return p.view(3, 3, phys[1], phys[2])[:, :, :n, self._cols(short, c, p.device)].permute(2, 3, 1, 0).contiguous()
        if len(phys) == 2:
            n, c = logical[:2]
            return p[:n, self._cols(short, c, p.device)].reshape(logical).clone()
        return p[: logical[0]].clone()

    def state_dict(self, prefix=""):
        """logical fp32 parameters under the reference's key names"""
        return {prefix + s: self._to_logical(s, self.view(self.P, s)) for s in self.index}

    def grads(self, prefix=""):
        return {prefix + s: self._to_logical(s, self.view(self.G, s)) for s in self.index}

    # ------------------------------------------------------------------ re-loading parameters (drop-in path: torch optimizers
    # update the nn.Parameters, the flat buffers follow)
    def _key_map(self):
        """{short name: reference state_dict key}; state_dict() walks self.index in order, so the two lists align"""
        if getattr(self, "_keys", None) is None:
            full = list(self.state_dict().keys())[:len(self.index)]
            self._keys = dict(zip(self.index.keys(), full))
        return self._keys

    def _adapt(self, short, v):
        """reference-shaped tensor -> this module's logical shape"""
        return v.reshape(self.index[short][2])

    def refresh_mirror(self):
        self.Wb.copy_(self.P)

    def load_params(self, state_dict):
        """overwrite the master parameters (and the bf16 mirror) from a reference-keyed state dict; optimizer moments are kept"""
        for short, full in self._key_map().items():
            v = state_dict[full].detach().to(self.dev, torch.float32)
            self.view(self.P, short).copy_(self._to_physical(short, self._adapt(short, v)))
        self.refresh_mirror()

    def ln(self, name):
        """(gamma, beta, dgamma, dbeta) views of a LayerNorm"""
        return tuple(self.view(f, "%s.%s" % (name, wb)) for f in (self.P, self.G) for wb in ("weight", "bias"))

    # ------------------------------------------------------------------ backward helpers
    def _weights(self):
        raise NotImplementedError

    def refresh_transposes(self):
        """flipped / transposed mirrors for the data-gradient GEMMs: every tap of every weight in ONE launch"""
        if self._wt_tables is None:
            pairs = []
            for w in self._weights():
                pairs += w.transposes()
            self._wt_tables = ops.transpose_batch_tables(pairs)
        ops.transpose_batch(self._wt_tables)

    def mask_grads(self):
        """Weights that read a shared stage buffer in place (layout flag `shared_input`) have physical columns for channels
        of OTHER consumers (e.g. the seg token under `pre_proj`): their value is zero, but the buffer holds data there, so
        the weight-gradient kernel fills them.  Zero them, or AdamW would grow weights the reference does not have."""
        for short, idx in self._unmapped.items():
            self.view(self.G, short).index_fill_(-1, idx, 0.0)

    @staticmethod
    def conv_bwd(cv, dY, X, need_dx=True, res=None, act_grad=None):
        """dY [B,H,W,n_pad], X [B,H,W,cin_pad] bf16: dW accumulated into the flat gradient view; returns dX (+ res), or dX times the
        derivative of the activation that produced X (act_grad = (saved, act, from_input, y_mul, scale), fused in the GEMM epilogue)"""
        fork_wgrad(lambda: ops.conv3x3_wgrad(dY, X, cv.gw), dY, X)
        if not need_dx:
            return None
        return conv_gemm(dY, cv.pwT, bias=False, res=res, res_mode=RES_AFTER if res is not None else RES_NONE, act_grad=act_grad)

    @staticmethod
    def lin_bwd(lin, dY, X, need_dx=True, res=None, x_coff=0, out=None, y_coff=0, accumulate=False, act_grad=None):
        """weight / bias gradient (forked) and the data gradient dX = dY W.  out / y_coff: write dX into a channel slice of a wider
        buffer; accumulate: add to what `out` already holds (in place: the epilogue reads the old tile and stores the sum)"""
        fork_wgrad(lambda: ops.linear_wgrad(dY, X, lin.gw, lin.gb, x_coff=x_coff), dY, X)
        if not need_dx:
            return None
        if accumulate:
            assert out is not None and res is None
            return conv_gemm(dY, lin.pwT, bias=False, res=out, res_coff=y_coff, res_mode=RES_AFTER, out=out, y_coff=y_coff)
        return conv_gemm(dY, lin.pwT, bias=False, res=res, res_mode=RES_AFTER if res is not None else RES_NONE, out=out, y_coff=y_coff,
                         act_grad=act_grad)

    # ------------------------------------------------------------------ optimizer
    def allreduce_grads(self):
        """the data-parallel exchange of this module: ONE sum all-reduce of the flat gradient buffer"""
        self._world = parallel.allreduce_sum_(self.G)
        return self._world

    def step(self, sumsq=None, reduced=False):
        """gradient all-reduce + clip + AdamW + mirror refresh.  `sumsq` (fp64 [1] on the device): the squared norm of the
        (all-reduced) gradients of ALL modules of the model when the clip is global (the reference clips the whole model,
        engine_glassrgbd.py:155-159); None = this module's own norm.  `reduced`: allreduce_grads() was already called (a
        shared norm has to be taken after the exchange)."""
        world = self._world if reduced else self.allreduce_grads()
        self.t += 1
        if sumsq is None:
            self.sumsq.zero_()
            ops.sumsq(self.G, self.sumsq)
            sumsq = self.sumsq
        ops.adamw_step(self.P, self.G, self.M, self.V, self.Wb, lr=self.lr, betas=self.betas, eps=self.eps,
                       weight_decay=self.weight_decay, step=self.t, max_norm=self.max_norm, grad_scale=1.0 / world,
                       sumsq_buf=sumsq)
