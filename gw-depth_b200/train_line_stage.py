"""Training path of the 1/32 LINE-WINDOW STAGE on B200: `dense_input_proj` and the BasicLayer of
four Swin blocks whose window attention is re-queried through the end points of the top-20 predicted lines (the
"glass-structure context"), forward with the tape kept and the backward on hand-written kernels.

Reference (under torch.autograd): `GlassRGBD.dense_input_proj` (src/models/glassrgbd.py:70,101), `BasicLayer.forward`
(src/models/multiscale_transformerr.py:926-979), `SwinTransformerBlock.forward` (:646-755: norm1 -> pad -> cyclic shift ->
nearest grid_sample of the shifted features + position code at the reference points (:676-701) -> window partition ->
attention -> reverse -> crop -> residual -> norm2 -> Mlp) and `WindowAttention.forward` (:267-332):

    q, k, v   = qkv(x_win);  ref = ref_qk(x_ref);  ref_k = diff_mu + exp(diff_logsigma) * ref[:, :D];  ref_v = ref[:, D:]
    a0        = (q * scale) @ ref_k^T                                     per (image, head): a [window tokens, R] plane
    a_{i+1}   = a_i + gelu(layer_norm_{plane}(conv3x3_{heads->heads}(a_i)))            three rounds, shared filter
    q_new     = (softmax_R(a3) @ ref_v) * scale                           (q is scaled twice)
    out       = proj(softmax(q_new @ k^T + relative position bias (+ shift mask)) @ v)

`grid_sample(mode='nearest')` has no gradient w.r.t. the sample coordinates, so nothing flows back into the predicted lines
(the line branch); `depth_pred32` only feeds the uncertainty sampling and never gets a gradient.

B200 design
* parameters of the four blocks + `dense_input_proj` in ONE flat fp32 buffer (train_flat.FlatModule); the 16x16x3x3 diffusion
  filter is re-laid-out on the device once per step (gwd_diffuse_filter_pack) into the layout the convolution kernels read and
  into its adjoint, so the optimizer's device-side update is seen without a host copy (the inference plan passes the
  filter by value);
* forward = the inference kernel sequence of engine.Engine.line_stage with q kept (it feeds the score backward) and q_new
  written into the q columns of a second fused [q_new | k | v] buffer, which is what the window-attention forward / backward
  kernels read;
* backward per block: Mlp -> LayerNorm -> window partition of the gradient -> `proj` -> gwd_window_attention_bwd (d q_new |
  dk | dv + bias gradient) -> gwd_ref_requery_bwd (soft-max backward, d ref_v) -> 3 x gwd_ref_diffuse_bwd (GELU, plane
  normalisation, adjoint convolution, filter / bias gradient) -> gwd_ref_scores_bwd (dq into the q columns, d ref_k) ->
  gwd_ref_affine_bwd (d mu, d logsigma, d ref) -> `ref_qk` -> `qkv` -> gwd_line_ref_scatter (reference-token gradients back
  onto their window tokens) -> window reverse -> LayerNorm backward with the shortcut gradient.
"""
import torch

from . import ops
from .engine import shift_mask, sine_table, sine_tables_masked, _compose
from .ops import ACT_GELU, ACT_SIGMOID, RES_AFTER, PackedWeight, conv_gemm, pack_linear
from .train_flat import FUSE_ACT_GRAD, join_wgrads, FlatModule, Linear

PREFIX = "dense_encoder.dense_transformer."
DIP = "dense_input_proj."


class LineStage(FlatModule):
    def __init__(self, state_dict, cfg, device="cuda", **optim):
        c = self.cfg = cfg
        self.D, self.heads, self.ws = c["dense_trans_dim"], c["dense_trans_heads"], c["window"]
        self.depth = c["dense_trans_layers"][0]
        self.hd, self.N = self.D // self.heads, self.ws * self.ws
        self.scale = self.hd ** -0.5
        assert self.heads == 16 and self.hd <= 32, "the diffusion kernels are built for 16 heads of <= 32 channels"
        tensors = {}
        for k, v in state_dict.items():
            if not v.is_floating_point():
                continue
            if k.startswith(PREFIX):
                short = k[len(PREFIX):]
                tensors[short] = v.reshape(-1) if short.endswith(("diff_mu", "diff_logsigma")) else v
            elif k.startswith(DIP):
                tensors[k] = v
        super().__init__(tensors, None, device=device, **optim)
        D = self.D
        rel = torch.stack(torch.meshgrid(torch.arange(self.ws), torch.arange(self.ws), indexing="ij")).flatten(1)
        rel = rel[:, :, None] - rel[:, None, :]
        self.rel_index = ((rel[0] + self.ws - 1) * (2 * self.ws - 1) + rel[1] + self.ws - 1).reshape(-1).to(self.dev)
        self.dip = Linear(self, DIP + "weight", DIP + "bias")
        self.blocks = []
        for i in range(self.depth):
            bp = "blocks.%d." % i
            lin = lambda n: Linear(self, bp + n + ".weight", bp + n + ".bias")  # noqa: E731
            qkv = lin("attn.qkv")
            bias = self.view(self.P, bp + "attn.qkv.bias")
            blk = dict(
                n1=self.ln(bp + "norm1"), n2=self.ln(bp + "norm2"), qkv=qkv, proj=lin("attn.proj"), ref=lin("attn.ref_qk"),
                mlp=(lin("mlp.fc1"), lin("mlp.fc2")),
                # forward views of the fused projection: q alone (kept for the score backward), k | v into the fused buffer
                q_pw=PackedWeight(qkv.wb[:D].view(1, D, D), bias[:D], 1, D, D),
                kv_pw=PackedWeight(qkv.wb[D:].view(1, 2 * D, D), bias[D:3 * D], 1, 2 * D, D),
                mu=self.view(self.P, bp + "attn.diff_mu"), ls=self.view(self.P, bp + "attn.diff_logsigma"),
                gmu=self.view(self.G, bp + "attn.diff_mu"), gls=self.view(self.G, bp + "attn.diff_logsigma"),
                fw=self.view(self.P, bp + "attn.ref_attn_diffusion.weight"), fb=self.view(self.P, bp + "attn.ref_attn_diffusion.bias"),
                gfw=self.view(self.G, bp + "attn.ref_attn_diffusion.weight"), gfb=self.view(self.G, bp + "attn.ref_attn_diffusion.bias"),
                filt=(torch.empty(2320, dtype=torch.float32, device=self.dev), torch.empty(2320, dtype=torch.float32, device=self.dev)),
                table=bp + "attn.relative_position_bias_table",
                dbias=torch.zeros(self.heads, self.N, self.N, dtype=torch.float32, device=self.dev))
            self.blocks.append(blk)
        # depth_pred32 = Linear o Linear o Sigmoid: never trained (no gradient reaches it), packed once
        p32 = "dense_encoder.depth_pred32."
        w, b = _compose(state_dict[p32 + "0.weight"].to(self.dev), state_dict[p32 + "0.bias"].to(self.dev),
                        state_dict[p32 + "1.weight"].to(self.dev), state_dict[p32 + "1.bias"].to(self.dev))
        self.depth32 = pack_linear(w.float(), b.float())
        self.tape, self._tables = None, {}

    def _weights(self):
        out = [self.dip]
        for b in self.blocks:
            out += [b["qkv"], b["proj"], b["ref"], b["mlp"][0], b["mlp"][1]]
        return out

    def _names(self, d):
        out = {}
        for k, v in d.items():
            if k.startswith(DIP):
                out[k] = v
            else:
                out[PREFIX + k] = v.view(1, 1, -1) if k.endswith(("diff_mu", "diff_logsigma")) else v
        return out

    def state_dict(self):
        return self._names(super().state_dict(""))

    def grads(self):
        return self._names(super().grads(""))

    def _table(self, key, fn):
        if key not in self._tables:
            self._tables[key] = fn()
        return self._tables[key]

    # ------------------------------------------------------------------ forward
    def forward(self, c5, ref_xy, pad_mask=None):
        """c5 bf16 [B,h,w,2048] (backbone C5, channels-last); ref_xy fp32 [B,R,2] reference points in [-1,1] (no gradient).
        Returns x32 bf16 [B,h,w,D] and depth0 fp32 [B,h,w] (depth_pred32: feeds the uncertainty sampling only)."""
        B, H, W, _ = c5.shape
        D, nh, hd, ws, N = self.D, self.heads, self.hd, self.ws, self.N
        R = ref_xy.shape[1]
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        nW = (Hp // ws) * (Wp // ws)
        P = nW * N
        # pad_mask: bool [B,H,W] padding mask of a ragged batch at this level -> per-image position codes (multiscale_transformerr.py:1035)
        pos = (self._table(("pos", H, W), lambda: sine_table(H, W, D // 2, False, self.dev)) if pad_mask is None
               else sine_tables_masked(pad_mask, D // 2, False))
        mask = self._table(("mask", H, W), lambda: shift_mask(H, W, ws, ws // 2, self.dev))
        c5tok = c5.reshape(B * H * W, c5.shape[-1])
        x = conv_gemm(c5tok, self.dip.pw)
        tp = self.tape = dict(B=B, H=H, W=W, R=R, nW=nW, c5=c5tok, ref_xy=ref_xy, blocks=[])
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        for i, blk in enumerate(self.blocks):
            shift = 0 if i % 2 == 0 else ws // 2
            ops.diffuse_filter_pack(blk["fw"], blk["fb"], *blk["filt"])
            bias = self.view(self.P, blk["table"])[self.rel_index].view(N, N, nh).permute(2, 0, 1).contiguous()
            xw = ops.window_gather(x, B, H, W, ws, shift, blk["n1"][0], blk["n1"][1])
            xref = ops.line_ref_gather(xw, pos, ref_xy, R, B, H, W, ws, shift, D).view(B * R, D)
            q = conv_gemm(xw, blk["q_pw"])
            qkv3 = torch.empty(B * P, 3 * D, **bf)
            conv_gemm(xw, blk["kv_pw"], out=qkv3, y_coff=D)
            ref = conv_gemm(xref, blk["ref"].pw, out_f32=True)                      # fp32 [B*R, 2D] = ref_qk | ref_v
            refk = ops.ref_affine(ref, blk["mu"], blk["ls"], D)
            a0 = torch.empty(B, nh, P, R, dtype=torch.float32, device=self.dev)
            ops.ref_scores(q, D, refk, D, a0, B, nW, N, nh, hd, R, scale=self.scale)
            a, rounds = a0, []
            for _ in range(3):
                a_next, raw, stats = ops.ref_diffuse_dev(a, blk["filt"][0], B, nh, P, R)
                rounds.append((a, raw, stats))
                a = a_next
            ops.ref_requery(a, ref[:, D:], 2 * D, qkv3, 3 * D, B, nW, N, nh, hd, R, self.scale)      # q_new -> columns [0, D)
            o = torch.empty(B * P, D, **bf)
            ops.attention(qkv3, qkv3[:, D:], qkv3[:, 2 * D:], o, items=B * nW, heads=nh, Lq=N, Lk=N, hd=hd,
                          q_strides=(N * 3 * D, 3 * D), k_strides=(N * 3 * D, 3 * D), v_strides=(N * 3 * D, 3 * D),
                          o_strides=(N * D, D), bias=bias, mask=mask if shift else None, scale=1.0)
            pr = conv_gemm(o, blk["proj"].pw)
            x_new, x_ln = ops.window_merge(pr, x, B, H, W, ws, shift, blk["n2"][0], blk["n2"][1], want_ln=True)
            fc1, fc2 = blk["mlp"]
            h_raw = torch.empty(B * H * W, fc1.n_pad, **bf)
            hmid = conv_gemm(x_ln, fc1.pw, post_act=ACT_GELU, y_raw=h_raw)
            x_out = conv_gemm(hmid, fc2.pw, res=x_new, res_mode=RES_AFTER)
            tp["blocks"].append(dict(shift=shift, mask=mask if shift else None, bias=bias, x=x, xw=xw, xref=xref, q=q, qkv3=qkv3,
                                     ref=ref, refk=refk, rounds=rounds, a3=a, o=o, x_new=x_new, x_ln=x_ln, h_raw=h_raw, hmid=hmid))
            x = x_out
        depth0 = conv_gemm(x, self.depth32, post_act=ACT_SIGMOID, out_f32=True).view(B, H, W)
        return x.view(B, H, W, D), depth0

    # ------------------------------------------------------------------ backward
    def backward(self, g_x32, keep_tape=False):
        """g_x32 bf16 [B*h*w, D] (or [B,h,w,D]) -> d C5 bf16 [B*h*w, 2048]; fills the flat gradient buffer"""
        tp = self.tape
        B, H, W, R, nW = tp["B"], tp["H"], tp["W"], tp["R"], tp["nW"]
        D, nh, hd, ws, N = self.D, self.heads, self.hd, self.ws, self.N
        P = nW * N
        rows = B * H * W
        g = g_x32.reshape(rows, D)
        self.refresh_transposes()
        self.G.zero_()
        zeros = torch.zeros(rows, D, dtype=torch.bfloat16, device=self.dev)
        ws_draw = torch.empty(B, nh, P, R, dtype=torch.float32, device=self.dev)
        ws_stats = torch.empty(4 * B * nh, dtype=torch.float64, device=self.dev)
        for blk, t in zip(reversed(self.blocks), reversed(tp["blocks"])):
            shift = t["shift"]
            fc1, fc2 = blk["mlp"]
            # y = x_new + fc2(gelu(fc1(LN2(x_new))))
            if FUSE_ACT_GRAD:
                d_hraw = self.lin_bwd(fc2, g, t["hmid"], act_grad=(t["h_raw"], ACT_GELU, True, 1.0, 1.0))
            else:
                d_hraw = ops.act_bwd(self.lin_bwd(fc2, g, t["hmid"]), t["h_raw"], ACT_GELU, from_input=True)
            d_ln = self.lin_bwd(fc1, d_hraw, t["x_ln"])
            g_new = ops.layernorm_bwd(d_ln, t["x_new"], blk["n2"][0], blk["n2"][2], blk["n2"][3], add=g)
            # x_new = x + unwindow(proj(attention))
            d_pr = ops.window_gather(g_new, B, H, W, ws, shift)
            d_o = self.lin_bwd(blk["proj"], d_pr, t["o"])
            blk["dbias"].zero_()
            dqkv3 = ops.window_attention_bwd(t["qkv3"], d_o, items=B * nW, heads=nh, N=N, hd=hd, scale=1.0, bias=t["bias"],
                                             mask=t["mask"], dbias=blk["dbias"])
            self.view(self.G, blk["table"]).index_add_(0, self.rel_index, blk["dbias"].permute(1, 2, 0).reshape(N * N, nh))
            # q_new = scale * softmax_R(a3) @ ref_v
            d_kv = torch.zeros(B * R, 2 * D, dtype=torch.float32, device=self.dev)         # d ref_k | d ref_v (accumulated)
            d_a = ops.ref_requery_bwd(t["a3"], t["ref"][:, D:], 2 * D, dqkv3, 3 * D, d_kv[:, D:], 2 * D, B, P, nh, hd, R, self.scale)
            for a_in, raw, stats in reversed(t["rounds"]):
                d_a = ops.ref_diffuse_bwd(d_a, raw, stats, a_in, blk["filt"][1], blk["gfw"], blk["gfb"], B, nh, P, R,
                                          ws=(ws_draw, ws_stats))
            # a0 = scale * q @ ref_k^T: dq overwrites the (consumed) d q_new columns of the fused gradient buffer
            ops.ref_scores_bwd(d_a, t["refk"], D, t["q"], D, dqkv3, 3 * D, d_kv, 2 * D, B, P, nh, hd, R, self.scale)
            d_ref = ops.ref_affine_bwd(d_kv, t["ref"], blk["ls"], blk["gmu"], blk["gls"], D)
            d_xref = self.lin_bwd(blk["ref"], d_ref, t["xref"])
            d_xw = self.lin_bwd(blk["qkv"], dqkv3, t["xw"])
            ops.line_ref_scatter(d_xref, tp["ref_xy"], R, d_xw, B, H, W, ws, shift, D)
            d_n1, _ = ops.window_merge(d_xw, zeros, B, H, W, ws, shift)
            g = ops.layernorm_bwd(d_n1, t["x"], blk["n1"][0], blk["n1"][2], blk["n1"][3], add=g_new)
        d_c5 = self.lin_bwd(self.dip, g, tp["c5"])
        join_wgrads()
        if not keep_tape:
            self.tape = None
        return d_c5
