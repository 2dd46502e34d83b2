"""Training step of the LINE BRANCH on B200: input_proj -> DETR encoder / decoder -> class_embed / lines_embed, the set
criterion (6 Hungarian matchings), the backward through hand-written kernels and a fused clip + AdamW update.

Reference: the modules of src/models/transformer.py:47-233, src/models/multi_head_attention.py:188-380 and
src/models/glassrgbd.py:87-90 under torch.autograd; SetCriterion (src/models/glassrgbd.py:308-358); the optimizer of
src/main_glassrgbd.py:59-67 (AdamW, lr 1e-4, weight decay 1e-4) and the clip of src/engine_glassrgbd.py:155-159
(max_norm 0.1).  Train-mode dropout (`--dropout`, default 0.1) is built with regenerated masks (see __init__); `--dropout 0.0`
is the gradient-parity configuration (masks of another generator cannot be compared element by element).

B200 design
* every trainable tensor of the branch lives in ONE flat fp32 master buffer with flat gradient / Adam-moment twins and
  a flat bf16 mirror; the packed weights the tcgen05 GEMM reads are VIEWS of the mirror, gradients are written by the
  kernels straight into views of the flat gradient buffer.  Data-parallel training is therefore one NCCL all-reduce of
  one buffer, and clip + AdamW + mirror refresh are two launches (gwd_sumsq, gwd_adamw_step) with the clip coefficient
  read on the device;
* forward = the inference kernel sequence of engine.Engine.detr with the pre-LayerNorm values kept (y_raw);
* backward: dX = dY W on gwd_conv_gemm (tcgen05) with the transposed weight mirror (gwd_transpose), dW = dY^T X and
  db on gwd_linear_wgrad (split-K mma.sync, both operands read as stored), gwd_layernorm_bwd, gwd_act_bwd,
  gwd_attention_bwd (soft-max recomputed, tensor cores).

Scope: gradients stop at the C5 feature map (the backbone / dense-branch backward is not built); `backward` returns
dC5 so that a backbone backward can be attached.
"""
import torch

from . import ops, parallel
from .train_flat import FUSE_ACT_GRAD
from .engine import DEFAULT_CFG, sine_table, sine_tables_masked
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID, RES_AFTER, RES_BEFORE_NORM, RES_NONE, PackedWeight, conv_gemm, round_up

PREFIXES = ("input_proj.", "query_embed.", "transformer.", "class_embed.", "lines_embed.")


class _Lin:
    """one Linear (or a row slice of a packed in-projection): views into the flat buffers + its transposed mirror"""

    def __init__(self, owner, wname, bname, rows=None):
        wb, gw, p_b, g_b = owner.view(owner.Wb, wname), owner.view(owner.G, wname), owner.view(owner.P, bname), owner.view(owner.G, bname)
        n = owner.index[wname][2][0]
        if rows is not None:
            r0, r1 = rows
            wb, gw, p_b, g_b, n = wb[r0:r1], gw[r0:r1], p_b[r0:r1], g_b[r0:r1], r1 - r0
        self.n, self.n_pad, self.k = n, wb.shape[0], wb.shape[1]
        self.w2d, self.gw, self.gb = wb, gw, g_b
        self.pw = PackedWeight(wb.view(1, self.n_pad, self.k), p_b, 1, n, self.k)
        self.wT = torch.empty(self.k, self.n_pad, dtype=torch.bfloat16, device=wb.device)
        self.pwT = PackedWeight(self.wT.view(1, self.k, self.n_pad), None, 1, self.k, self.n_pad)


class LineBranch:
    def __init__(self, state_dict, cfg=None, device="cuda", lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                 max_norm=0.1):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("LineBranch runs on libgwd_b200 CUDA kernels only (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.t = 0
        # soft-max scale hd^-0.5 (multi_head_attention.py:236): the q and k projections are stored pre-scaled by its square
        # root each (fused q|k GEMM, out_scale) or q alone by the whole factor (cross attention), so the attention kernels
        # run with scale 1 -- the tcgen05 forward kernel's fast path -- and the backward multiplies dQ / dK back
        self.rs = float((self.cfg["hidden_dim"] // self.cfg["nheads"]) ** -0.25)
        # train-mode dropout of the DETR layers (src/args.py:51 default 0.1; transformer.py:149-162,212-233: dropout1/2/3 on the
        # sub-layer outputs, `dropout` on the FFN hidden layer; multi_head_attention.py:368: on the attention probabilities).
        # Masks are never stored: every site regenerates its mask from (step seed in device memory, site id, element index)
        # in the backward; the seed is incremented by the forward itself, so a replayed CUDA graph draws new masks every step.
        self.p_drop = float(self.cfg.get("dropout", 0.0) or 0.0)
        import torch.distributed as _dist
        rank = _dist.get_rank() if _dist.is_available() and _dist.is_initialized() else 0
        self.drop_seed = torch.full((1,), (torch.initial_seed() * 2654435761 + rank * 7919) & 0x3FFFFFFF, dtype=torch.int32, device=self.dev)
        # ---- flat layout: 2-D weights [N, K] are stored with N padded to 16 (zero rows), vectors padded to 16
        self.index, self.shapes, off = {}, {}, 0
        for name, v in state_dict.items():
            if not (name.startswith(PREFIXES) and v.is_floating_point()):
                continue
            logical = tuple(v.shape[:2]) if v.dim() == 4 else tuple(v.shape)
            padded = (round_up(logical[0], 16),) + logical[1:]
            assert len(logical) <= 2 and (len(logical) == 1 or logical[1] % 16 == 0), name
            size = padded[0] * (padded[1] if len(padded) == 2 else 1)
            self.index[name] = (off, padded, logical)
            self.shapes[name] = tuple(v.shape)           # e.g. input_proj.weight is [256, 2048, 1, 1] in the reference
            off += size
        self.numel = off
        self.P = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.G, self.M, self.V = torch.zeros_like(self.P), torch.zeros_like(self.P), torch.zeros_like(self.P)
        for name, (o, padded, logical) in self.index.items():
            self.view(self.P, name)[: logical[0]] = state_dict[name].detach().to(self.dev, torch.float32).reshape(logical)
        self.Wb = self.P.to(torch.bfloat16)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self._build_views()
        self._tables, self.tape, self._graphs, self._wt_tables = {}, None, {}, None
        import os
        self.use_cuda_graph = os.environ.get("GWD_CUDA_GRAPH", "1") != "0"
        self._side, self._side2, self._keep = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev), []

    # ------------------------------------------------------------------ flat-buffer views
    def view(self, flat, name):
        o, padded, _ = self.index[name]
        n = padded[0] * (padded[1] if len(padded) == 2 else 1)
        return flat[o:o + n].view(padded)

    def state_dict(self):
        """logical fp32 parameters under the reference's key names"""
        return {name: self.view(self.P, name)[: lg[0]].clone().view(self.shapes[name]) for name, (_, _, lg) in self.index.items()}

    def grads(self):
        """logical gradients in the reference's shapes (views of the flat gradient buffer)"""
        return {name: self.view(self.G, name)[: lg[0]].view(self.shapes[name]) for name, (_, _, lg) in self.index.items()}

    def load_params(self, state_dict):
        """overwrite the master parameters (and the bf16 mirror) from a reference-keyed state dict; optimizer moments are kept"""
        for name, (_, _, logical) in self.index.items():
            self.view(self.P, name)[: logical[0]] = state_dict[name].detach().to(self.dev, torch.float32).reshape(logical)
        self.Wb.copy_(self.P)

    def _mha(self, p, fused):
        E = self.cfg["hidden_dim"]
        w, b = p + "in_proj_weight", p + "in_proj_bias"
        d = {"v": _Lin(self, w, b, (2 * E, 3 * E)), "o": _Lin(self, p + "out_proj.weight", p + "out_proj.bias")}
        if fused:
            d["qk"] = _Lin(self, w, b, (0, 2 * E))
        else:
            d["q"], d["k"] = _Lin(self, w, b, (0, E)), _Lin(self, w, b, (E, 2 * E))
        return d

    def _ln(self, name):
        return (self.view(self.P, name + ".weight"), self.view(self.P, name + ".bias"),
                self.view(self.G, name + ".weight"), self.view(self.G, name + ".bias"))

    def _build_views(self):
        c = self.cfg
        L = lambda n: _Lin(self, n + ".weight", n + ".bias")
        self.input_proj = L("input_proj")
        self.enc = []
        for i in range(c["enc_layers"]):
            p = "transformer.encoder.layers.%d." % i
            self.enc.append({"attn": self._mha(p + "self_attn.", True), "l1": L(p + "linear1"), "l2": L(p + "linear2"),
                             "n1": self._ln(p + "norm1"), "n2": self._ln(p + "norm2")})
        self.dec = []
        for i in range(c["dec_layers"]):
            p = "transformer.decoder.layers.%d." % i
            self.dec.append({"self": self._mha(p + "self_attn.", True), "cross": self._mha(p + "multihead_attn.", False),
                             "l1": L(p + "linear1"), "l2": L(p + "linear2"), "n1": self._ln(p + "norm1"),
                             "n2": self._ln(p + "norm2"), "n3": self._ln(p + "norm3")})
        self.dec_norm = self._ln("transformer.decoder.norm")
        self.class_embed = L("class_embed")
        self.lines_embed = [L("lines_embed.layers.%d" % i) for i in range(3)]
        Q = c["num_queries"]
        self.query_pos = self.view(self.Wb, "query_embed.weight")[:Q]
        self.g_query = self.view(self.G, "query_embed.weight")[:Q]
        self.lins = [self.input_proj, self.class_embed] + self.lines_embed
        for ly in self.enc:
            self.lins += [ly["attn"]["qk"], ly["attn"]["v"], ly["attn"]["o"], ly["l1"], ly["l2"]]
        for ly in self.dec:
            self.lins += [ly["self"]["qk"], ly["self"]["v"], ly["self"]["o"], ly["cross"]["q"], ly["cross"]["k"],
                          ly["cross"]["v"], ly["cross"]["o"], ly["l1"], ly["l2"]]

    def refresh_transposes(self):
        """W^T mirrors for the data-gradient GEMMs (first thing in every backward: the mirror changes with every step)"""
        if self._wt_tables is None:
            self._wt_tables = ops.transpose_batch_tables([(lin.w2d, lin.wT) for lin in self.lins])
        ops.transpose_batch(self._wt_tables)

    # ------------------------------------------------------------------ forward (activations kept for the backward)
    def _drop(self, site):
        return (self.drop_seed, site, self.p_drop) if self.p_drop > 0 else None

    def _attend(self, q, k, v, B, Lq, Lk, q_rs, k_rs, site=0, kpm=None):
        E, nh = self.cfg["hidden_dim"], self.cfg["nheads"]
        o = torch.empty(B * Lq, E, dtype=torch.bfloat16, device=self.dev)
        ops.attention(q, k, v, o, items=B, heads=nh, Lq=Lq, Lk=Lk, hd=E // nh, q_strides=(Lq * q_rs, q_rs),
                      k_strides=(Lk * k_rs, k_rs), v_strides=(Lk * E, E), o_strides=(Lq * E, E), dropout=self._drop(site),
                      key_padding=kpm)
        return o

    def _fork_gemm(self, x, lin):
        """y = Linear(x) on the side stream (a parallel branch of the captured graph); the caller joins before using y"""
        y = torch.empty(x.shape[0], lin.n_pad, dtype=torch.bfloat16, device=self.dev)
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            conv_gemm(x, lin.pw, out=y)
        return y

    def _ln_gemm(self, x, lin, res, ln, site=0):
        """LN(res + dropout(Linear(x))) -> (output, pre-LayerNorm sum kept for the backward)"""
        if self.p_drop > 0:      # the dropout sits between the GEMM and the residual add: un-fused epilogue
            z = ops.dropout(conv_gemm(x, lin.pw), self.drop_seed, site, self.p_drop, res=res)
            return ops.layernorm(z, ln[0], ln[1]), z
        z = torch.empty(x.shape[0], lin.n_pad, dtype=torch.bfloat16, device=self.dev)
        y = conv_gemm(x, lin.pw, res=res, res_mode=RES_BEFORE_NORM, ln=(ln[0], ln[1]), y_raw=z)
        return y, z

    def _ffn_hidden(self, x, lin, site):
        hm = conv_gemm(x, lin.pw, post_act=ACT_RELU)
        if self.p_drop > 0:
            ops.dropout(hm, self.drop_seed, site, self.p_drop, out=hm)
        return hm

    def forward(self, c5, mask5=None):
        """c5: bf16 channels-last [B, h, w, 2048] (no gradient flows further back); mask5: None, or the bool [B,h,w] padding mask of a
        ragged batch at this level (True = padding): per-image position codes and a key-padding mask on the encoder self-attention and
        the decoder cross-attention (src/models/transformer.py:52-57).  Returns fp32 (logits [6,B,Q,2], lines [6,B,Q,6]) and keeps the
        tape for `backward`."""
        c = self.cfg
        B, h, w, _ = c5.shape
        E, L, Q = c["hidden_dim"], h * w, c["num_queries"]
        if mask5 is None:
            key = ("pos5", h, w)
            if key not in self._tables:
                self._tables[key] = sine_table(h, w, E // 2, True, self.dev).to(torch.bfloat16)
            pos, period, kpm = self._tables[key], L, None
        else:
            pos, period = sine_tables_masked(mask5, E // 2, True).view(B * L, E).to(torch.bfloat16), B * L
            kpm = mask5.reshape(B, L).to(torch.uint8).contiguous()
        tp = self.tape = {"B": B, "L": L, "enc": [], "dec": [], "kpm": kpm}
        if self.p_drop > 0:
            self.drop_seed.add_(1)          # new masks every step (captured: every graph replay increments it too)
        tp["c5"] = c5.reshape(B * L, c5.shape[-1])
        x = conv_gemm(tp["c5"], self.input_proj.pw)
        for li, ly in enumerate(self.enc):
            a = ly["attn"]
            st = 10 * li
            xp = ops.add_rows(x, pos, period)
            v = self._fork_gemm(x, a["v"])
            qk = conv_gemm(xp, a["qk"].pw, out_scale=self.rs)
            torch.cuda.current_stream().wait_stream(self._side)
            o = self._attend(qk, qk[:, E:], v, B, L, L, 2 * E, 2 * E, site=st + 1, kpm=kpm)
            x1, z1 = self._ln_gemm(o, a["o"], x, ly["n1"], site=st + 2)
            hm = self._ffn_hidden(x1, ly["l1"], st + 3)
            x2, z2 = self._ln_gemm(hm, ly["l2"], x1, ly["n2"], site=st + 4)
            tp["enc"].append(dict(x=x, xp=xp, qk=qk, v=v, o=o, z1=z1, x1=x1, h=hm, z2=z2))
            x = x2
        memory = x
        mem_pos = ops.add_rows(memory, pos, period)
        tp["memory"], tp["mem_pos"] = memory, mem_pos
        tgt = torch.zeros(B * Q, E, dtype=torch.bfloat16, device=self.dev)
        hs = torch.empty(len(self.dec), B * Q, E, dtype=torch.bfloat16, device=self.dev)
        # K / V projections of the memory for every decoder layer: a parallel branch next to the decoder chain
        ckv = [(torch.empty(B * L, E, dtype=torch.bfloat16, device=self.dev), torch.empty(B * L, E, dtype=torch.bfloat16, device=self.dev))
               for _ in self.dec]
        self._side2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side2):
            for ly, (ck, cv) in zip(self.dec, ckv):
                conv_gemm(mem_pos, ly["cross"]["k"].pw, out=ck)
                conv_gemm(memory, ly["cross"]["v"].pw, out=cv)
        for i, ly in enumerate(self.dec):
            s, cr = ly["self"], ly["cross"]
            st = 100 + 10 * i
            tq1 = ops.add_rows(tgt, self.query_pos, Q)
            v = self._fork_gemm(tgt, s["v"])
            qk = conv_gemm(tq1, s["qk"].pw, out_scale=self.rs)
            torch.cuda.current_stream().wait_stream(self._side)
            o1 = self._attend(qk, qk[:, E:], v, B, Q, Q, 2 * E, 2 * E, site=st + 1)
            x1, z1 = self._ln_gemm(o1, s["o"], tgt, ly["n1"], site=st + 2)
            tq2 = ops.add_rows(x1, self.query_pos, Q)
            cq = conv_gemm(tq2, cr["q"].pw, out_scale=self.rs * self.rs)
            ck, cv = ckv[i]
            if i == 0:
                torch.cuda.current_stream().wait_stream(self._side2)
            o2 = self._attend(cq, ck, cv, B, Q, L, E, E, site=st + 3, kpm=kpm)
            x2, z2 = self._ln_gemm(o2, cr["o"], x1, ly["n2"], site=st + 4)
            hm = self._ffn_hidden(x2, ly["l1"], st + 5)
            x3, z3 = self._ln_gemm(hm, ly["l2"], x2, ly["n3"], site=st + 6)
            ops.layernorm(x3, self.dec_norm[0], self.dec_norm[1], out=hs[i])
            tp["dec"].append(dict(x0=tgt, tq1=tq1, qk=qk, v=v, o1=o1, z1=z1, x1=x1, tq2=tq2, cq=cq, ck=ck, cv=cv, o2=o2,
                                  z2=z2, x2=x2, h=hm, z3=z3, x3=x3))
            tgt = x3
        flat = hs.view(-1, E)
        logits = conv_gemm(flat, self.class_embed.pw, out_f32=True)
        t1 = conv_gemm(flat, self.lines_embed[0].pw, post_act=ACT_RELU)
        t2 = conv_gemm(t1, self.lines_embed[1].pw, post_act=ACT_RELU)
        lines = conv_gemm(t2, self.lines_embed[2].pw, post_act=ACT_SIGMOID, out_f32=True)
        tp.update(flat=flat, t1=t1, t2=t2, lines=lines)
        nl = len(self.dec)
        return logits.view(nl, B, Q, -1), lines.view(nl, B, Q, -1)

    # ------------------------------------------------------------------ backward
    def _lin_bwd(self, lin, dY, X, need_dx=True, res=None, act_grad=None):
        """dY [R, n_pad] bf16, X [R, K] bf16: accumulates dW / db into the flat gradient views; returns dX (+ res) or None.
        The weight gradient is off the critical path (nothing downstream reads it before the optimizer), so it runs on a
        side stream: in the captured graph it becomes a parallel branch next to the dX chain."""
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)                 # dY (and, first time round, the zeroed gradient buffer) is ready
        with torch.cuda.stream(self._side):
            ops.linear_wgrad(dY, X, lin.gw, lin.gb)
        self._keep.append(dY)                        # the side stream still reads it: no reuse before the join
        if not need_dx:
            return None
        return conv_gemm(dY, lin.pwT, bias=False, res=res, res_mode=RES_AFTER if res is not None else RES_NONE, act_grad=act_grad)

    def _ln_bwd(self, dy, z, ln, add=None):
        return ops.layernorm_bwd(dy, z, ln[0], ln[2], ln[3], add=add)

    def _sub_grad(self, dz, site):
        """gradient of the sub-layer output from the gradient of the pre-LayerNorm sum z = res + dropout(y)"""
        return ops.dropout(dz, self.drop_seed, site, self.p_drop) if self.p_drop > 0 else dz

    def _attend_bwd(self, q, k, v, o, d_o, B, Lq, Lk, q_rs, k_rs, dq, dk, dv, fused=True, site=0, kpm=None):
        E, nh = self.cfg["hidden_dim"], self.cfg["nheads"]
        ops.attention_bwd(q, k, v, d_o, dq, dk, dv, items=B, heads=nh, Lq=Lq, Lk=Lk, hd=E // nh,
                          q_strides=(Lq * q_rs, q_rs), k_strides=(Lk * k_rs, k_rs), v_strides=(Lk * E, E),
                          do_strides=(Lq * E, E), dq_strides=(Lq * q_rs, q_rs), dk_strides=(Lk * k_rs, k_rs),
                          dv_strides=(Lk * E, E), scale=1.0, o=o, o_strides=(Lq * E, E),
                          dq_mul=self.rs if fused else self.rs * self.rs, dk_mul=self.rs if fused else 1.0, dropout=self._drop(site),
                          key_padding=kpm)

    def _ffn_bwd(self, ly, d_out, z, hm, x_in, ln, site_hidden=0, site_out=0):
        dz = self._ln_bwd(d_out, z, ln)
        # hm is the hidden layer AFTER its dropout: hm > 0 <=> active and kept, and kept units carry the factor 1 / (1 - p)
        keep_scale = 1.0 / (1.0 - self.p_drop) if self.p_drop > 0 else 1.0
        if FUSE_ACT_GRAD:        # relu' (and the dropout factor) in the epilogue of linear2's data-gradient GEMM
            dpre = self._lin_bwd(ly["l2"], self._sub_grad(dz, site_out), hm, act_grad=(hm, ACT_RELU, False, 1.0, keep_scale))
        else:
            dpre = ops.act_bwd(self._lin_bwd(ly["l2"], self._sub_grad(dz, site_out), hm), hm, ACT_RELU, scale=keep_scale)
        return self._lin_bwd(ly["l1"], dpre, x_in, res=dz)

    def _add(self, a, b):
        return ops.add_rows(a, b, a.shape[0])

    def backward(self, dlogits, dlines, keep_tape=False):
        """dlogits [6,B,Q,2], dlines [6,B,Q,6] fp32 -> fills the flat gradient buffer; returns dC5 [B*L, 2048] bf16"""
        tp, c = self.tape, self.cfg
        B, L, E, Q = tp["B"], tp["L"], c["hidden_dim"], c["num_queries"]
        bf = dict(dtype=torch.bfloat16, device=self.dev)
        self.refresh_transposes()
        self.G.zero_()
        self._keep = []
        gq = torch.zeros(Q, E, dtype=torch.float32, device=self.dev)
        # ---- heads
        dl = ops.act_bwd(dlogits.reshape(-1, dlogits.shape[-1]).contiguous(), None, ACT_NONE, out_cols=self.class_embed.n_pad)
        dflat = self._lin_bwd(self.class_embed, dl, tp["flat"])
        d3 = ops.act_bwd(dlines.reshape(-1, dlines.shape[-1]).contiguous(), tp["lines"], ACT_SIGMOID,
                         out_cols=self.lines_embed[2].n_pad)
        dt2 = ops.act_bwd(self._lin_bwd(self.lines_embed[2], d3, tp["t2"]), tp["t2"], ACT_RELU)
        dt1 = ops.act_bwd(self._lin_bwd(self.lines_embed[1], dt2, tp["t1"]), tp["t1"], ACT_RELU)
        dhs = self._lin_bwd(self.lines_embed[0], dt1, tp["flat"], res=dflat).view(len(self.dec), B * Q, E)
        # ---- decoder
        dmem, d_next = None, None
        for i in reversed(range(len(self.dec))):
            ly, s = self.dec[i], tp["dec"][i]
            st = 100 + 10 * i
            d_x3 = self._ln_bwd(dhs[i], s["x3"], self.dec_norm, add=d_next)
            d_x2 = self._ffn_bwd(ly, d_x3, s["z3"], s["h"], s["x2"], ly["n3"], st + 5, st + 6)
            # cross attention
            cr = ly["cross"]
            dz2 = self._ln_bwd(d_x2, s["z2"], ly["n2"])
            d_o = self._lin_bwd(cr["o"], self._sub_grad(dz2, st + 4), s["o2"])
            dq, dk, dv = torch.empty(B * Q, E, **bf), torch.empty(B * L, E, **bf), torch.empty(B * L, E, **bf)
            self._attend_bwd(s["cq"], s["ck"], s["cv"], s["o2"], d_o, B, Q, L, E, E, dq, dk, dv, fused=False, site=st + 3, kpm=tp["kpm"])
            dtq = self._lin_bwd(cr["q"], dq, s["tq2"])
            gq += dtq.view(B, Q, E).sum(0, dtype=torch.float32)
            d_x1 = self._add(dtq, dz2)
            dmem = self._lin_bwd(cr["k"], dk, tp["mem_pos"], res=dmem)
            dmem = self._lin_bwd(cr["v"], dv, tp["memory"], res=dmem)
            # self attention
            sa = ly["self"]
            dz1 = self._ln_bwd(d_x1, s["z1"], ly["n1"])
            d_o = self._lin_bwd(sa["o"], self._sub_grad(dz1, st + 2), s["o1"])
            dqk, dv = torch.empty(B * Q, 2 * E, **bf), torch.empty(B * Q, E, **bf)
            self._attend_bwd(s["qk"], s["qk"][:, E:], s["v"], s["o1"], d_o, B, Q, Q, 2 * E, 2 * E, dqk, dqk[:, E:], dv, site=st + 1)
            dtq = self._lin_bwd(sa["qk"], dqk, s["tq1"])
            gq += dtq.view(B, Q, E).sum(0, dtype=torch.float32)
            d_next = self._lin_bwd(sa["v"], dv, s["x0"], res=self._add(dtq, dz1))
        self.g_query += gq
        # ---- encoder
        d_x = dmem
        for i in reversed(range(len(self.enc))):
            ly, s = self.enc[i], tp["enc"][i]
            st = 10 * i
            d_x1 = self._ffn_bwd(ly, d_x, s["z2"], s["h"], s["x1"], ly["n2"], st + 3, st + 4)
            a = ly["attn"]
            dz1 = self._ln_bwd(d_x1, s["z1"], ly["n1"])
            d_o = self._lin_bwd(a["o"], self._sub_grad(dz1, st + 2), s["o"])
            dqk, dv = torch.empty(B * L, 2 * E, **bf), torch.empty(B * L, E, **bf)
            self._attend_bwd(s["qk"], s["qk"][:, E:], s["v"], s["o"], d_o, B, L, L, 2 * E, 2 * E, dqk, dqk[:, E:], dv, site=st + 1,
                             kpm=tp["kpm"])
            t = self._lin_bwd(a["qk"], dqk, s["xp"], res=dz1)
            d_x = self._lin_bwd(a["v"], dv, s["x"], res=t)
        dc5 = self._lin_bwd(self.input_proj, d_x, tp["c5"])
        torch.cuda.current_stream().wait_stream(self._side)       # join: all weight gradients are in the flat buffer
        self._keep = []
        if not keep_tape:
            self.tape = None
        return dc5

    # ------------------------------------------------------------------ optimizer
    def step(self, sumsq=None, reduced=False):
        """gradient all-reduce (one NCCL call on the flat buffer) + global-norm clip + AdamW + bf16 mirror refresh.  Same
        interface as train_flat.FlatModule.step: `sumsq` (fp64 [1] on the device) = the squared norm over ALL modules of the model
        when the clip is global (the reference clips the whole model, engine_glassrgbd.py:155-159; train_model.Trainer.step does
        that), None = this branch's own norm; `reduced` = the gradients have already been exchanged."""
        world = parallel.world_size() if reduced else parallel.allreduce_sum_(self.G)
        self.t += 1
        if sumsq is None:
            self.sumsq.zero_()
            ops.sumsq(self.G, self.sumsq)
            sumsq = self.sumsq
        ops.adamw_step(self.P, self.G, self.M, self.V, self.Wb, lr=self.lr, betas=self.betas, eps=self.eps,
                       weight_decay=self.weight_decay, step=self.t, max_norm=self.max_norm, grad_scale=1.0 / world,
                       sumsq_buf=sumsq)

    # ------------------------------------------------------------------ CUDA graphs
    def _captured(self, c5, producer=None):
        """forward and backward have no host synchronisation, so each is captured once per input shape (same memory pool:
        the backward graph reads the activations the forward graph leaves behind) and replayed as one launch.  With a
        `producer` (e.g. the frozen backbone: images -> C5) the input is the producer's input and its kernels are part of
        the forward graph."""
        key = (tuple(c5.shape), producer is not None)
        st = self._graphs.get(key)
        if st is None:
            st = {"c5": c5.clone()}
            feed = (lambda: producer(st["c5"])) if producer is not None else (lambda: st["c5"])
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up: lazy kernel attributes, cached tables, cuDNN plans
                lo, li = self.forward(feed())
                self.backward(torch.zeros_like(lo), torch.zeros_like(li))
            torch.cuda.current_stream().wait_stream(side)
            st["fwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["fwd"]):
                st["logits"], st["lines"] = self.forward(feed())
            st["tape"] = self.tape
            st["dlogits"], st["dlines"] = torch.zeros_like(st["logits"]), torch.zeros_like(st["lines"])
            st["bwd"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st["bwd"], pool=st["fwd"].pool()):
                st["dc5"] = self.backward(st["dlogits"], st["dlines"], keep_tape=True)
            self._graphs[key] = st
        return st

    def loss_and_grads(self, c5, targets, criterion, producer=None):
        """forward + SetCriterion + backward; returns (weighted total loss tensor, dict of losses, dC5).  `producer`: optional
        gradient-free front end (images -> C5, e.g. Engine.backbone); then `c5` is ITS input."""
        st = self._captured(c5, producer) if self.use_cuda_graph else None
        if st is not None:
            st["c5"].copy_(c5, non_blocking=True)
            st["fwd"].replay()
            logits, lines = st["logits"], st["lines"]
        else:
            with torch.no_grad():
                logits, lines = self.forward(producer(c5) if producer is not None else c5)
        losses, dlogits, dlines = criterion.forward_backward_stacked(logits, lines, targets)
        total = criterion.last_total
        self.last_cotangents = (dlogits, dlines)
        if st is not None:
            st["dlogits"].copy_(dlogits, non_blocking=True)
            st["dlines"].copy_(dlines, non_blocking=True)
            st["bwd"].replay()
            dc5 = st["dc5"]
        else:
            dc5 = self.backward(dlogits, dlines)
        return total.detach(), {k: v.detach() for k, v in losses.items()}, dc5

    def train_step(self, c5, targets, criterion, producer=None):
        total, losses, _ = self.loss_and_grads(c5, targets, criterion, producer)
        self.step()
        return total, losses
