"""Training step of the DENSE PREDICTION HEAD on B200: DensePrediction forward with the activations kept, SilogLoss on
the full-resolution depth + SegLoss on the segmentation logits, the backward through hand-written kernels and the fused
clip + AdamW update of train.LineBranch.

Reference: src/models/dense_upsample.py:74-111,160-182 (`DensePrediction`, `upconv`, `Mlp`) under torch.autograd;
`SilogLoss` / `SegLoss` (src/models/glassrgbd.py:360-383) as the engine applies them to the last prediction scale
(src/engine_glassrgbd.py:65-90: depth weight `depth_loss_weights[-1]` = 1, `seg_loss_weight` = 2).

B200 design
* the 22 tensors of `depth_decoder.*` live in ONE flat fp32 master buffer in the PHYSICAL layout the tcgen05 GEMM reads
  (Linear [n_pad, cin_pad] with the channel map of the 1/4-scale stage buffer, 3x3 conv [tap][n_pad][cin_pad]) with flat
  gradient / Adam-moment twins and a bf16 mirror: forward weights are views of the mirror, weight-gradient kernels
  accumulate straight into views of the flat gradient buffer, so clip + AdamW + mirror refresh are two launches and
  data-parallel training is one all-reduce;
* forward = the inference kernel sequence of engine.Engine.dense_head with the values the backward needs kept by the
  GEMM epilogue (y_raw: the GELU input, the ELU output ahead of the LayerNorm); `upconv` runs as gwd_upsample_nearest +
  a plain 3x3 conv here, because the up-sampled map is the weight-gradient operand anyway;
* backward: loss gradients are produced as the padded bf16 dY rows of the last convolutions (gwd_silog_bwd with the
  sigmoid * max_depth derivative folded, gwd_seg_ce), dW on gwd_conv3x3_wgrad / gwd_linear_wgrad, dX on gwd_conv_gemm
  with the flipped / transposed mirror (all 110 tap transposes of a step in one gwd_transpose_batch launch), ELU / GELU
  by gwd_act_bwd, LayerNorm by gwd_layernorm_bwd, nearest x2 up-sampling backward = 2x2 sum = gwd_avgpool with the factor 4
  folded into the next gwd_act_bwd.

`backward` returns the gradient of the 1/4-scale stage buffer [x3 | depth token | seg token | depth_pred3] so that the
backward of the dense encoder (not built) can be attached.
"""
import torch

from . import ops, parallel
from .engine import DEFAULT_CFG
from .ops import ACT_ELU, ACT_GELU, ACT_NONE, ACT_SIGMOID, RES_AFTER, RES_NONE, PackedWeight, conv_gemm, round_up

PREFIX = "depth_decoder."
KINDS = ("depth", "seg")


class _Conv:
    """one bias-free 3x3 convolution: mirror / gradient views [9, n_pad, cin_pad] + the flipped, transposed mirror"""

    def __init__(self, owner, name):
        self.name = name
        self.wb, self.gw = owner.view(owner.Wb, name), owner.view(owner.G, name)
        _, n_pad, c_pad = self.wb.shape
        n, c = owner.index[name][2][:2]
        self.pw = PackedWeight(self.wb, None, 9, n, c_pad)
        self.wT = torch.empty(9, c_pad, n_pad, dtype=torch.bfloat16, device=self.wb.device)
        self.pwT = PackedWeight(self.wT, None, 9, c, n_pad)

    def transposes(self):
        # data gradient of a stride-1 pad-1 conv = the conv of dY with the filter flipped in (dy, dx) = tap 8 - t
        return [(self.wb[8 - t], self.wT[t]) for t in range(9)]


class _Lin:
    def __init__(self, owner, name):
        self.wb, self.gw = owner.view(owner.Wb, name + ".weight"), owner.view(owner.G, name + ".weight")
        self.gb = owner.view(owner.G, name + ".bias")
        n_pad, k = self.wb.shape
        n = owner.index[name + ".weight"][2][0]
        self.n, self.n_pad, self.k = n, n_pad, k
        self.pw = PackedWeight(self.wb.view(1, n_pad, k), owner.view(owner.P, name + ".bias"), 1, n, k)
        self.wT = torch.empty(k, n_pad, dtype=torch.bfloat16, device=self.wb.device)
        self.pwT = PackedWeight(self.wT.view(1, k, n_pad), None, 1, k, n_pad)

    def transposes(self):
        return [(self.wb, self.wT)]


class DenseHead:
    def __init__(self, state_dict, cfg=None, device="cuda", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                 max_norm=0.1, depth_weight=1.0, seg_weight=2.0):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DenseHead runs on libgwd_b200 CUDA kernels only (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.depth_weight, self.seg_weight = depth_weight, seg_weight
        self.t = 0
        c = self.cfg
        C, td = c["dense_trans_dim"] >> 3, c["class_token_dim"]
        self.C, self.td, self.width = C, td, C + 3 * td     # stage buffer: [x3 | depth token | seg token | depth_pred3 (+ pad)]
        feat, dtok, stok = torch.arange(C), torch.arange(td) + C, torch.arange(td) + C + td
        self.col_maps = {"depth_token_fuse.fc1.weight": torch.cat([feat, torch.tensor([C + 2 * td]), dtok]),
                         "seg_token_fuse.fc1.weight": torch.cat([feat, stok])}
        # ---- flat layout (physical shapes)
        self.index, off = {}, 0
        for name, v in state_dict.items():
            if not (name.startswith(PREFIX) and v.is_floating_point()):
                continue
            short = name[len(PREFIX):]
            if v.dim() == 4:
                phys = (9, round_up(v.shape[0], 16), round_up(v.shape[1], 16))
            elif v.dim() == 2:
                phys = (round_up(v.shape[0], 16), self.width if short in self.col_maps else round_up(v.shape[1], 16))
            else:
                phys = (round_up(v.shape[0], 16),)
            size = 1
            for s in phys:
                size *= s
            self.index[short] = (off, phys, tuple(v.shape))
            off += size
        self.numel = off
        self.P = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.G, self.M, self.V = torch.zeros_like(self.P), torch.zeros_like(self.P), torch.zeros_like(self.P)
        for short in self.index:
            self.view(self.P, short).copy_(self._to_physical(short, state_dict[PREFIX + short].detach().to(self.dev, torch.float32)))
        self.Wb = self.P.to(torch.bfloat16)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.losses = torch.zeros(2, dtype=torch.float32, device=self.dev)     # weighted (depth, seg) loss of the last step
        self._build_views()
        self.tape, self._wt_tables = None, None

    # ------------------------------------------------------------------ flat-buffer views, logical <-> physical layout
    def view(self, flat, short):
        o, phys, _ = self.index[short]
        n = 1
        for s in phys:
            n *= s
        return flat[o:o + n].view(phys)

    def _to_physical(self, short, v):
        _, phys, logical = self.index[short]
        out = torch.zeros(phys, dtype=torch.float32, device=v.device)
        if len(logical) == 4:       # [N, C, 3(dy), 3(dx)] -> [dx*3+dy][N][C]
            n, c = logical[:2]
            out.view(3, 3, phys[1], phys[2])[:, :, :n, :c] = v.permute(3, 2, 0, 1)
        elif len(logical) == 2:
            cm = self.col_maps.get(short)
            if cm is None:
                out[: logical[0], : logical[1]] = v
            else:
                out[: logical[0], cm.to(v.device)] = v
        else:
            out[: logical[0]] = v
        return out

    def _to_logical(self, short, p):
        _, phys, logical = self.index[short]
        if len(logical) == 4:
            return ops.unpack_conv3x3_grad(p, logical[0], logical[1])
        if len(logical) == 2:
            cm = self.col_maps.get(short)
            return (p[: logical[0], : logical[1]] if cm is None else p[: logical[0], cm.to(p.device)]).clone()
        return p[: logical[0]].clone()

    def state_dict(self):
        """logical fp32 parameters under the reference's key names"""
        return {PREFIX + s: self._to_logical(s, self.view(self.P, s)) for s in self.index}

    def grads(self):
        return {PREFIX + s: self._to_logical(s, self.view(self.G, s)) for s in self.index}

    def _build_views(self):
        self.mods = {}
        for kind in KINDS:
            m = {"fc1": _Lin(self, kind + "_token_fuse.fc1"), "fc2": _Lin(self, kind + "_token_fuse.fc2")}
            for cv in ("upconv1_%s.conv", "conv1_%s.0", "upconv2_%s.conv", "conv2_%s.0"):
                m[cv.split("_")[0]] = _Conv(self, (cv % kind) + ".weight")
            m["norm"] = tuple(self.view(f, "norm_%s.%s" % (kind, wb)) for f in (self.P, self.G) for wb in ("weight", "bias"))
            self.mods[kind] = m
        self.mods["depth"]["last"] = _Conv(self, "get_depth.0.weight")
        self.mods["seg"]["last"] = _Conv(self, "get_seg.weight")

    def refresh_transposes(self):
        if self._wt_tables is None:
            pairs = []
            for m in self.mods.values():
                for v in m.values():
                    if isinstance(v, (_Conv, _Lin)):
                        pairs += v.transposes()
            self._wt_tables = ops.transpose_batch_tables(pairs)
        ops.transpose_batch(self._wt_tables)

    # ------------------------------------------------------------------ forward (activations kept for the backward)
    def forward(self, buf4, H, W):
        """buf4: bf16 channels-last [B, H/4, W/4, width] = [x3 | depth token | seg token | depth_pred3, zero pad].
        Returns depth fp32 [B,1,H,W] (metres) and seg logits fp32 [B,2,H,W] (a channels-last view)."""
        B, H4, W4, width = buf4.shape
        assert width == self.width and buf4.dtype == torch.bfloat16 and buf4.is_contiguous()
        if (2 * H4, 2 * W4) != (H // 2, W // 2) or (H % 4 or W % 4):
            raise NotImplementedError("the dense head needs input sizes that are multiples of 4")
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        rows = B * H4 * W4
        x2d = buf4.view(rows, width)
        tp = self.tape = {"B": B, "H4": H4, "W4": W4, "H": H, "W": W, "x": x2d}
        outs = {}
        for kind in KINDS:
            m = self.mods[kind]
            hid = m["fc2"].k
            h_raw = torch.empty(rows, hid, **bf)
            t = conv_gemm(x2d, m["fc1"].pw, post_act=ACT_GELU, out_channels=hid, y_raw=h_raw)
            f = conv_gemm(t, m["fc2"].pw).view(B, H4, W4, -1)
            fu = ops.upsample_nearest(f, 2 * H4, 2 * W4)
            e1 = torch.empty(B, 2 * H4, 2 * W4, m["upconv1"].pw.n_pad, **bf)
            u1 = conv_gemm(fu, m["upconv1"].pw, bias=False, pre_act=ACT_ELU, ln=(m["norm"][0], m["norm"][1]), y_raw=e1)
            c1 = conv_gemm(u1, m["conv1"].pw, bias=False, post_act=ACT_ELU)
            c1u = ops.upsample_nearest(c1, H, W)
            u2 = conv_gemm(c1u, m["upconv2"].pw, bias=False, post_act=ACT_ELU)
            c2 = conv_gemm(u2, m["conv2"].pw, bias=False, post_act=ACT_ELU)
            tp[kind] = dict(h_raw=h_raw, t=t, fu=fu, e1=e1, u1=u1, c1=c1, c1u=c1u, u2=u2, c2=c2)
            outs[kind] = c2
        depth = conv_gemm(outs["depth"], self.mods["depth"]["last"].pw, bias=False, post_act=ACT_SIGMOID,
                          out_scale=float(self.cfg["max_depth"]), out_f32=True).view(B, 1, H, W)
        seg = conv_gemm(outs["seg"], self.mods["seg"]["last"].pw, bias=False, out_f32=True).permute(0, 3, 1, 2)
        return depth, seg

    # ------------------------------------------------------------------ backward
    def _conv_bwd(self, cv, dY, X, need_dx=True):
        ops.conv3x3_wgrad(dY, X, cv.gw)
        return conv_gemm(dY, cv.pwT, bias=False) if need_dx else None

    def _lin_bwd(self, lin, dY, X, res=None):
        ops.linear_wgrad(dY, X, lin.gw, lin.gb)
        return conv_gemm(dY, lin.pwT, bias=False, res=res, res_mode=RES_AFTER if res is not None else RES_NONE)

    def backward(self, g_depth, g_seg, keep_tape=False):
        """g_depth / g_seg: bf16 [B*H*W, 16] rows = gradients to the outputs of get_depth (ahead of the sigmoid) / get_seg in
        columns 0 / 0-1.  Fills the flat gradient buffer; returns d(buf4) [B*H4*W4, width] bf16."""
        tp = self.tape
        B, H4, W4, H, W = tp["B"], tp["H4"], tp["W4"], tp["H"], tp["W"]
        self.refresh_transposes()
        self.G.zero_()
        d_x = None
        for kind, g in (("depth", g_depth), ("seg", g_seg)):
            m, s = self.mods[kind], tp[kind]
            img = lambda d, h, w: d.view(B, h, w, d.shape[-1])
            d = self._conv_bwd(m["last"], img(g, H, W), s["c2"])
            d = self._conv_bwd(m["conv2"], img(ops.act_bwd(d, s["c2"], ACT_ELU), H, W), s["u2"])
            d = self._conv_bwd(m["upconv2"], img(ops.act_bwd(d, s["u2"], ACT_ELU), H, W), s["c1u"])
            # nearest x2 up-sampling backward: every low-resolution pixel collects its 2x2 block (mean * 4)
            d = ops.act_bwd(ops.avgpool(d, 2), s["c1"], ACT_ELU, scale=4.0)
            d = self._conv_bwd(m["conv1"], img(d, H // 2, W // 2), s["u1"])
            gam, _, dgam, dbet = m["norm"]
            d = ops.layernorm_bwd(d, s["e1"], gam, dgam, dbet)
            d = self._conv_bwd(m["upconv1"], img(ops.act_bwd(d, s["e1"], ACT_ELU), H // 2, W // 2), s["fu"])
            d = ops.act_bwd(ops.avgpool(d, 2), None, ACT_NONE, scale=4.0)
            d = self._lin_bwd(m["fc2"], d, s["t"])
            d = ops.act_bwd(d, s["h_raw"], ACT_GELU, from_input=True)
            d_x = self._lin_bwd(m["fc1"], d, tp["x"], res=d_x)
        if not keep_tape:
            self.tape = None
        return d_x

    # ------------------------------------------------------------------ losses
    def loss_and_grads(self, buf4, depth_gt, seg_gt, H=None, W=None):
        """forward + SilogLoss (last scale) + SegLoss + backward.  depth_gt fp32 [B,1,H,W] metres, seg_gt int64 [B,1,H,W].
        Returns (depth, seg, losses fp32 [2] on the device = weighted (loss_depth, loss_seg), d(buf4))."""
        H, W = H or depth_gt.shape[-2], W or depth_gt.shape[-1]
        depth, seg = self.forward(buf4, H, W)
        log_only = bool(self.cfg.get("log_depth_error", False))
        sums = ops.silog_sums(depth, depth_gt, log_only=log_only)
        g_depth = ops.silog_bwd(depth, depth_gt, sums, weight=self.depth_weight, log_only=log_only,
                                variance_focus=float(self.cfg.get("variance_focus", 0.85)), sig_scale=float(self.cfg["max_depth"]),
                                out_cols=16, loss_out=self.losses[0:1])
        _, g_seg = ops.seg_ce(seg, seg_gt.contiguous(), weight=self.seg_weight, out_cols=16, loss_out=self.losses[1:2])
        d_buf4 = self.backward(g_depth, g_seg)
        return depth, seg, self.losses, d_buf4

    # ------------------------------------------------------------------ optimizer
    def step(self, sumsq=None):
        """gradient all-reduce + clip + AdamW + mirror refresh.  `sumsq` (fp64 [1] on the device): the squared gradient norm
        of ALL modules of the model when the clip is global (the reference clips the whole model, engine_glassrgbd.py:155-159);
        None = this module's own norm."""
        world = parallel.allreduce_sum_(self.G)
        self.t += 1
        if sumsq is None:
            self.sumsq.zero_()
            ops.sumsq(self.G, self.sumsq)
            sumsq = self.sumsq
        ops.adamw_step(self.P, self.G, self.M, self.V, self.Wb, lr=self.lr, betas=self.betas, eps=self.eps,
                       weight_decay=self.weight_decay, step=self.t, max_norm=self.max_norm, grad_scale=1.0 / world,
                       sumsq_buf=sumsq)

    def train_step(self, buf4, depth_gt, seg_gt):
        _, _, losses, _ = self.loss_and_grads(buf4, depth_gt, seg_gt)
        self.step()
        return losses
