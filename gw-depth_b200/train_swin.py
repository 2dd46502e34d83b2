"""Training path of a CLASS-WINDOW SWIN STAGE on B200: the BasicLayer of
WindowClassAttention blocks that runs at 1/16, 1/8 and 1/4 scale on [features | depth token | seg token], forward with the
tape kept and the backward on hand-written kernels.

Reference (under torch.autograd): `BasicLayer.forward` (src/models/multiscale_transformerr.py:926-979),
`SwinTransformerBlock.forward` (:646-788: norm1 -> pad -> cyclic shift -> window partition -> attention -> reverse -> crop
-> residual -> norm2 -> Mlp, the same for both tokens), `WindowClassAttention.forward` (:455-580, group_attention=False:
biased window MSA on the features + class-token CHANNEL attention whose keys / values come from cat[x_attn, depth, seg]; the
seg token is projected with `proj_dth` too, :578 -- `proj_seg`, `diff_*`, `border_*` never run and have no gradient).

B200 design
* parameters of all blocks of the stage in ONE flat fp32 buffer (train_flat.FlatModule); `global_k` / `global_v` are adjacent
  in it, so one GEMM produces k | v and one weight-gradient launch fills both; the relative-position bias is re-gathered
  from the fp32 table each step and its gradient scattered back with one index_add_;
* forward = the inference kernel sequence of engine.Engine.class_stage with the softmax scale as a kernel argument (the
  inference path folds it into the weights) and the three streams kept as separate contiguous tensors, so that every
  LayerNorm input survives for its backward;
* backward per block: Mlp (Linear dgrad on tcgen05 / wgrad, GELU from the kept pre-activation) -> gwd_layernorm_bwd with
  the residual gradient added -> window partition of the gradient (gwd_window_gather without LayerNorm = the adjoint of the
  merge) -> `proj_dth` -> gwd_token_attention_bwd -> global_k | global_v, cls_*_q -> `proj` -> gwd_window_attention_bwd (dq |
  dk | dv + bias gradient, persistent per head) -> `qkv` -> window reverse of the gradient (gwd_window_merge on a zero
  shortcut = the adjoint of the gather) -> gwd_layernorm_bwd of norm1 / norm_depth1 / norm_seg1 with the shortcut gradient.
"""
import torch

from . import ops
from .engine import shift_mask
from .ops import ACT_GELU, RES_AFTER, PackedWeight, conv_gemm
from .train_flat import FUSE_ACT_GRAD, join_wgrads, FlatModule, Linear

_DEAD = ("attn.diff_mu", "attn.diff_logsigma", "attn.border_mu", "attn.border_logsigma", "attn.proj_seg.weight",
         "attn.proj_seg.bias")


class FusedLinear:
    """several Linears with the same input, adjacent in the flat buffer, as one [sum n, k] weight (k | v of the token attention)"""

    def __init__(self, owner, wnames, bnames):
        offs = [owner.index[n] for n in wnames]
        k = offs[0][1][1]
        rows = sum(o[1][0] for o in offs)
        pos = offs[0][0]
        for o, phys, logical in offs:
            assert o == pos and phys[1] == k and phys == tuple(logical), "fused Linears must be adjacent and un-padded"
            pos += phys[0] * k
        boffs = [owner.index[n] for n in bnames]
        bpos = boffs[0][0]
        for o, phys, _ in boffs:
            assert o == bpos
            bpos += phys[0]
        w0, b0 = offs[0][0], boffs[0][0]
        self.wb, self.gw = owner.Wb[w0:w0 + rows * k].view(rows, k), owner.G[w0:w0 + rows * k].view(rows, k)
        self.gb = owner.G[b0:b0 + rows]
        self.n = self.n_pad = rows
        self.k = k
        self.pw = PackedWeight(self.wb.view(1, rows, k), owner.P[b0:b0 + rows], 1, rows, k)
        self.wT = torch.empty(k, rows, dtype=torch.bfloat16, device=self.wb.device)
        self.pwT = PackedWeight(self.wT.view(1, k, rows), None, 1, k, rows)

    def transposes(self):
        return [(self.wb, self.wT)]


class ClassStage(FlatModule):
    def __init__(self, state_dict, prefix, C, depth, heads=16, ws=7, token_dim=64, device="cuda", **optim):
        """prefix: e.g. 'dense_encoder.class_transformer3.'; C: feature channels of the stage; depth: blocks"""
        self.prefix, self.C, self.depth, self.heads, self.ws, self.td = prefix, C, depth, heads, ws, token_dim
        self.hd, self.tC, self.N = C // heads, C + 2 * token_dim, ws * ws
        assert C % 16 == 0 and token_dim % 16 == 0 and self.hd <= 32 and self.N <= 64
        tensors = {}
        for i in range(depth):
            bp = "blocks.%d." % i
            names = [k[len(prefix):] for k in state_dict if k.startswith(prefix + bp) and state_dict[k].is_floating_point()]
            names = [n for n in names if n[len(bp):] not in _DEAD]
            # global_k then global_v (weights, then biases) adjacent: one fused GEMM
            front = [bp + "attn.global_k.weight", bp + "attn.global_v.weight", bp + "attn.global_k.bias", bp + "attn.global_v.bias"]
            for n in front + [n for n in names if n not in front]:
                tensors[n] = state_dict[prefix + n]
        super().__init__(tensors, None, device=device, **optim)
        self.scale = self.hd ** -0.5
        rel = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = rel[:, :, None] - rel[:, None, :]
        self.rel_index = ((rel[0] + ws - 1) * (2 * ws - 1) + rel[1] + ws - 1).reshape(-1).to(self.dev)
        self.blocks = []
        for i in range(depth):
            bp = "blocks.%d." % i
            lin = lambda n: Linear(self, bp + n + ".weight", bp + n + ".bias")
            mlp = lambda n: (lin(n + ".fc1"), lin(n + ".fc2"))
            self.blocks.append(dict(
                n1=self.ln(bp + "norm1"), nd1=self.ln(bp + "norm_depth1"), ns1=self.ln(bp + "norm_seg1"),
                n2=self.ln(bp + "norm2"), nd2=self.ln(bp + "norm_depth2"), ns2=self.ln(bp + "norm_seg2"),
                qkv=lin("attn.qkv"), proj=lin("attn.proj"), dq=lin("attn.cls_dth_q"), sq=lin("attn.cls_seg_q"),
                pdth=lin("attn.proj_dth"),
                gkv=FusedLinear(self, [bp + "attn.global_k.weight", bp + "attn.global_v.weight"],
                                [bp + "attn.global_k.bias", bp + "attn.global_v.bias"]),
                mlp=mlp("mlp"), mlp_d=mlp("mlp_depth"), mlp_s=mlp("mlp_seg"),
                table=bp + "attn.relative_position_bias_table",
                dbias=torch.zeros(heads, self.N, self.N, dtype=torch.float32, device=self.dev)))
        self.tape = None
        self._masks = {}

    def _weights(self):
        out = []
        for b in self.blocks:
            out += [b[k] for k in ("qkv", "proj", "dq", "sq", "pdth", "gkv")]
            out += [l for k in ("mlp", "mlp_d", "mlp_s") for l in b[k]]
        return out

    def state_dict(self):
        return super().state_dict(self.prefix)

    def grads(self):
        return super().grads(self.prefix)

    def _mask(self, H, W):
        key = (H, W)
        if key not in self._masks:
            self._masks[key] = shift_mask(H, W, self.ws, self.ws // 2, self.dev)
        return self._masks[key]

    # ------------------------------------------------------------------ forward
    def _mlp_fwd(self, x_ln, x_res, mlp):
        fc1, fc2 = mlp
        h_raw = torch.empty(x_ln.shape[0], fc1.n_pad, dtype=torch.bfloat16, device=self.dev)
        hmid = conv_gemm(x_ln, fc1.pw, post_act=ACT_GELU, y_raw=h_raw)
        return conv_gemm(hmid, fc2.pw, res=x_res, res_mode=RES_AFTER), h_raw, hmid

    def forward(self, x, d, s, B, H, W):
        """x bf16 [B*H*W, C], d / s bf16 [B*H*W, token_dim] (contiguous) -> the three streams after the stage"""
        C, td, tC, nh, hd, ws, N = self.C, self.td, self.tC, self.heads, self.hd, self.ws, self.N
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        nW = (Hp // ws) * (Wp // ws)
        rows_w = B * nW * N
        assert x.shape == (B * H * W, C) and d.shape == s.shape == (B * H * W, td)
        self.tape = dict(B=B, H=H, W=W, blocks=[])
        for i, blk in enumerate(self.blocks):
            shift = 0 if i % 2 == 0 else ws // 2
            mask = self._mask(H, W) if shift else None
            # relative-position bias [heads, N, N] gathered from the fp32 table (device index: no host copy, graph-capturable)
            bias = self.view(self.P, blk["table"])[self.rel_index].view(N, N, nh).permute(2, 0, 1).contiguous()
            xw = ops.window_gather(x, B, H, W, ws, shift, blk["n1"][0], blk["n1"][1], C=C)
            tx = torch.empty(rows_w, tC, dtype=torch.bfloat16, device=self.dev)          # cat[x_attn, depth tok, seg tok]
            ops.window_gather(d, B, H, W, ws, shift, blk["nd1"][0], blk["nd1"][1], C=td, out=tx, y_coff=C)
            ops.window_gather(s, B, H, W, ws, shift, blk["ns1"][0], blk["ns1"][1], C=td, out=tx, y_coff=C + td)
            qkv = conv_gemm(xw, blk["qkv"].pw)
            o = torch.empty(rows_w, C, dtype=torch.bfloat16, device=self.dev)
            ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, items=B * nW, heads=nh, Lq=N, Lk=N, hd=hd,
                          q_strides=(N * 3 * C, 3 * C), k_strides=(N * 3 * C, 3 * C), v_strides=(N * 3 * C, 3 * C),
                          o_strides=(N * C, C), bias=bias, mask=mask, scale=self.scale)
            conv_gemm(o, blk["proj"].pw, out=tx, y_coff=0)
            dq = conv_gemm(tx, blk["dq"].pw, x_coff=C)
            sq = conv_gemm(tx, blk["sq"].pw, x_coff=C + td)
            gkv = conv_gemm(tx, blk["gkv"].pw)                                           # [rows_w, 2 tC] = global_k | global_v
            dout = torch.empty(rows_w, td, dtype=torch.bfloat16, device=self.dev)
            sout = torch.empty_like(dout)
            ops.token_attention(dq, sq, gkv, gkv[:, tC:], dout, sout, items=B * nW, N=N, heads=nh, td=td // nh, tc=tC // nh,
                                q_rs=td, k_rs=2 * tC, v_rs=2 * tC, o_rs=td, scale=self.scale)
            dpr = conv_gemm(dout, blk["pdth"].pw)
            spr = conv_gemm(sout, blk["pdth"].pw)
            x_new, x_ln = ops.window_merge(tx, x, B, H, W, ws, shift, blk["n2"][0], blk["n2"][1], want_ln=True, C=C)
            d_new, d_ln = ops.window_merge(dpr, d, B, H, W, ws, shift, blk["nd2"][0], blk["nd2"][1], want_ln=True, C=td)
            s_new, s_ln = ops.window_merge(spr, s, B, H, W, ws, shift, blk["ns2"][0], blk["ns2"][1], want_ln=True, C=td)
            x_out, xh_raw, xh = self._mlp_fwd(x_ln, x_new, blk["mlp"])
            d_out, dh_raw, dh = self._mlp_fwd(d_ln, d_new, blk["mlp_d"])
            s_out, sh_raw, sh = self._mlp_fwd(s_ln, s_new, blk["mlp_s"])
            self.tape["blocks"].append(dict(
                shift=shift, mask=mask, bias=bias, x=x, d=d, s=s, xw=xw, tx=tx, qkv=qkv, o=o, dq=dq, sq=sq, gkv=gkv, dout=dout,
                sout=sout, new=(x_new, d_new, s_new), ln=(x_ln, d_ln, s_ln), h_raw=(xh_raw, dh_raw, sh_raw), h=(xh, dh, sh)))
            x, d, s = x_out, d_out, s_out
        return x, d, s

    # ------------------------------------------------------------------ backward
    def _mlp_ln_bwd(self, g, mlp, ln, x_new, x_ln, h_raw, hmid):
        """y = x_new + fc2(gelu(fc1(LN(x_new)))): returns d(x_new)"""
        fc1, fc2 = mlp
        if FUSE_ACT_GRAD:        # d(fc1 output) = (g W_fc2) * gelu'(h_raw) in the epilogue of the data-gradient GEMM
            d_hraw = self.lin_bwd(fc2, g, hmid, act_grad=(h_raw, ACT_GELU, True, 1.0, 1.0))
        else:
            d_hraw = ops.act_bwd(self.lin_bwd(fc2, g, hmid), h_raw, ACT_GELU, from_input=True)
        d_ln = self.lin_bwd(fc1, d_hraw, x_ln)
        return ops.layernorm_bwd(d_ln, x_new, ln[0], ln[2], ln[3], add=g)

    def backward(self, g_x, g_d, g_s, keep_tape=False):
        """gradients of the three output streams (bf16, contiguous) -> gradients of the three input streams; fills the flat
        gradient buffer"""
        tp = self.tape
        B, H, W = tp["B"], tp["H"], tp["W"]
        C, td, tC, nh, hd, ws, N = self.C, self.td, self.tC, self.heads, self.hd, self.ws, self.N
        rows = B * H * W
        self.refresh_transposes()
        self.G.zero_()
        zeros = torch.zeros(rows, max(C, td), dtype=torch.bfloat16, device=self.dev)
        for blk, t in zip(reversed(self.blocks), reversed(tp["blocks"])):
            shift = t["shift"]
            g_new = [self._mlp_ln_bwd(g, blk[m], blk[n], t["new"][j], t["ln"][j], t["h_raw"][j], t["h"][j])
                     for j, (g, m, n) in enumerate(((g_x, "mlp", "n2"), (g_d, "mlp_d", "nd2"), (g_s, "mlp_s", "ns2")))]
            # adjoint of the window merge: partition the gradient (zero rows in the padding)
            # d tx [rows_w, tC] is assembled in place: the residual-branch gradient of the features lands in its first C columns, the
            # two token-query data gradients in the token columns, and the gkv data gradient is accumulated over the whole width
            Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
            d_tx = torch.empty(B * Hp * Wp, tC, dtype=torch.bfloat16, device=self.dev)
            ops.window_gather(g_new[0], B, H, W, ws, shift, C=C, out=d_tx, y_coff=0)
            d_dpr = ops.window_gather(g_new[1], B, H, W, ws, shift, C=td)
            d_spr = ops.window_gather(g_new[2], B, H, W, ws, shift, C=td)
            d_dout = self.lin_bwd(blk["pdth"], d_dpr, t["dout"])
            d_sout = self.lin_bwd(blk["pdth"], d_spr, t["sout"])
            g_dq, g_sq, g_gkv = ops.token_attention_bwd(t["dq"], t["sq"], t["gkv"], d_dout, d_sout, items=d_tx.shape[0] // N, N=N,
                                                        heads=nh, td=td // nh, tc=tC // nh, scale=self.scale)
            self.lin_bwd(blk["dq"], g_dq, t["tx"], x_coff=C, out=d_tx, y_coff=C)
            self.lin_bwd(blk["sq"], g_sq, t["tx"], x_coff=C + td, out=d_tx, y_coff=C + td)
            self.lin_bwd(blk["gkv"], g_gkv, t["tx"], out=d_tx, accumulate=True)            # [rows_w, tC]
            d_o = self.lin_bwd(blk["proj"], d_tx, t["o"])                                  # reads the first C columns of d_tx
            blk["dbias"].zero_()
            dqkv = ops.window_attention_bwd(t["qkv"], d_o, items=d_o.shape[0] // N, heads=nh, N=N, hd=hd, scale=self.scale,
                                            bias=t["bias"], mask=t["mask"], dbias=blk["dbias"])
            self.view(self.G, blk["table"]).index_add_(0, self.rel_index, blk["dbias"].permute(1, 2, 0).reshape(N * N, nh))
            d_xw = self.lin_bwd(blk["qkv"], dqkv, t["xw"])
            # adjoint of LayerNorm + window gather: reverse the windows, then the LayerNorm backward with the shortcut gradient
            outs = []
            for win, coff, width, ln, src, g_sc in ((d_xw, 0, C, "n1", t["x"], g_new[0]), (d_tx, C, td, "nd1", t["d"], g_new[1]),
                                                    (d_tx, C + td, td, "ns1", t["s"], g_new[2])):
                d_ln, _ = ops.window_merge(win, zeros, B, H, W, ws, shift, C=width, win_coff=coff)
                outs.append(ops.layernorm_bwd(d_ln, src, blk[ln][0], blk[ln][2], blk[ln][3], add=g_sc))
            g_x, g_d, g_s = outs
        join_wgrads()
        if not keep_tape:
            self.tape = None
        return g_x, g_d, g_s
