"""Training path of a dense-branch STAGE ENTRY on B200: what `ReferTransformer.forward` does between two Swin stages
(src/models/multiscale_transformerr.py:1226-1243 for 1/8, :1263-1280 for 1/4, under torch.autograd):
    x    = proj_class(nearest_up(prev features)) + ConvA(backbone map)        ConvA = 3x3 conv + bias + GELU (:104-118)
    dtok = MlpNorm(nearest_up(prev depth token)),  stok likewise              MlpNorm = fc1 -> fc2 -> LayerNorm (:75-102)

B200 design: `proj_class` and the token MlpNorms are per-pixel maps, so they commute with the nearest up-sampling and run at
the LOW resolution (a quarter of the rows), exactly as the inference engine does; the backward of the x2 nearest up-sampling
is a 2x2 sum (gwd_avgpool x 4).  The ConvA weight gradient runs on the tcgen05 wgrad kernel; its bias gradient is the column
sum of dY.  Parameters live in one flat buffer (train_flat.FlatModule).
"""
import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, PackedWeight, conv_gemm
from .train_flat import join_wgrads, Conv3x3, FlatModule, Linear

P0 = "dense_encoder."


class StageEntry(FlatModule):
    def __init__(self, state_dict, si, device="cuda", **optim):
        """si = 1 (entry of the 1/16 stage: the class tokens are the `depth_token` / `seg_token` parameters, :1196-1199),
        2 (entry of the 1/8 stage) or 3 (entry of the 1/4 stage)"""
        assert si in (1, 2, 3)
        self.si = si
        sc = {1: None, 2: "8", 3: "4"}[si]
        self.names = dict(pc="proj_class%d" % si, cb="proj_backbn%d.conv" % si)
        keys = [self.names["pc"] + ".weight", self.names["pc"] + ".bias", self.names["cb"] + ".weight", self.names["cb"] + ".bias"]
        if si == 1:
            tensors = {k: state_dict[P0 + k] for k in keys}
            tensors.update({k: state_dict[P0 + k].reshape(-1) for k in ("depth_token", "seg_token")})      # [1,1,td] -> [td]
        else:
            self.names.update(d="old_depth_token_proj" + sc, s="old_seg_token_proj" + sc)
            for kind in ("d", "s"):
                keys += ["%s.%s.%s" % (self.names[kind], l, wb) for l in ("fc1", "fc2", "norm") for wb in ("weight", "bias")]
            tensors = {k: state_dict[P0 + k] for k in keys}
        super().__init__(tensors, None, device=device, **optim)
        self.pc = Linear(self, self.names["pc"] + ".weight", self.names["pc"] + ".bias")
        self.cb = Conv3x3(self, self.names["cb"] + ".weight")
        self.cb_bias, self.cb_gbias = self.view(self.P, self.names["cb"] + ".bias"), self.view(self.G, self.names["cb"] + ".bias")
        self.cb.pw = PackedWeight(self.cb.wb, self.cb_bias, 9, self.cb.n, self.cb.c_pad)
        self.tok = {} if si == 1 else {kind: (Linear(self, self.names[kind] + ".fc1.weight", self.names[kind] + ".fc1.bias"),
                                              Linear(self, self.names[kind] + ".fc2.weight", self.names[kind] + ".fc2.bias"),
                                              self.ln(self.names[kind] + ".norm")) for kind in ("d", "s")}
        self.C = self.pc.n
        self.td = self.index["depth_token"][2][0] if si == 1 else self.tok["d"][1].n
        self.tape = None

    def _weights(self):
        return [self.pc, self.cb] + [l for kind in self.tok for l in self.tok[kind][:2]]

    def _named(self, d):
        return {k: (v.view(1, 1, -1) if k.endswith("_token") else v) for k, v in d.items()}

    def state_dict(self):
        return self._named(super().state_dict(P0))

    def grads(self):
        return self._named(super().grads(P0))

    def forward(self, prev_x, prev_d, prev_s, feat):
        """prev_x bf16 [B,h,w,Cprev]; prev_d / prev_s bf16 [B*h*w, td] (None for si = 1); feat bf16 [B,2h,2w,Cb] (backbone map)
        -> x [B*H*W, C], d, s [B*H*W, td] at the doubled resolution"""
        B, h, w, _ = prev_x.shape
        H, W = feat.shape[1:3]
        if (H, W) != (2 * h, 2 * w):
            raise NotImplementedError("stage entries are built for the exact x2 nearest up-sampling (input sizes that are multiples of 32)")
        C, td = self.C, self.td
        pc = conv_gemm(prev_x, self.pc.pw)
        z_cb = torch.empty(B, H, W, self.cb.n_pad, dtype=torch.bfloat16, device=self.dev)
        cb = conv_gemm(feat, self.cb.pw, post_act=ACT_GELU, y_raw=z_cb)
        x = ops.upsample_nearest(pc.view(B, h, w, C), H, W, add=cb)
        toks, tp = [], dict(B=B, h=h, w=w, prev_x=prev_x, feat=feat, z_cb=z_cb)
        if self.si == 1:      # the token parameters, one copy per pixel
            toks = [self.view(self.Wb, k)[:td].view(1, td).expand(B * H * W, td).contiguous() for k in ("depth_token", "seg_token")]
        for kind, prev in (() if self.si == 1 else (("d", prev_d), ("s", prev_s))):
            fc1, fc2, ln = self.tok[kind]
            t1 = conv_gemm(prev, fc1.pw)
            t2 = conv_gemm(t1, fc2.pw)
            t = ops.layernorm(t2, ln[0], ln[1])
            toks.append(ops.upsample_nearest(t.view(B, h, w, td), H, W).view(B * H * W, td))
            tp[kind] = (prev, t1, t2)
        self.tape = tp
        return x.view(B * H * W, C), toks[0], toks[1]

    def backward(self, g_x, g_d, g_s, need_dfeat=False, keep_tape=False):
        """gradients of (x, d, s) (bf16, contiguous) -> (d prev_x [B,h,w,Cprev], d prev_d, d prev_s [B*h*w, td], d feat or None)"""
        tp = self.tape
        B, h, w = tp["B"], tp["h"], tp["w"]
        H, W, C, td = 2 * h, 2 * w, self.C, self.td
        self.refresh_transposes()
        self.G.zero_()
        # ConvA
        d_z = ops.act_bwd(g_x, tp["z_cb"].view(-1, self.cb.n_pad), ACT_GELU, from_input=True)
        self.cb_gbias.add_(d_z.float().sum(0))
        d_feat = self.conv_bwd(self.cb, d_z.view(B, H, W, -1), tp["feat"], need_dx=need_dfeat)
        # x2 nearest up-sampling backward = 2x2 sum, then proj_class at the low resolution
        low = lambda g, n: ops.act_bwd(ops.avgpool(g.view(B, H, W, n), 2), None, ACT_NONE, scale=4.0)
        d_prev_x = self.lin_bwd(self.pc, low(g_x, C), tp["prev_x"].view(B * h * w, -1)).view(B, h, w, -1)
        outs = [None, None]
        if self.si == 1:
            for k, g in (("depth_token", g_d), ("seg_token", g_s)):
                self.view(self.G, k)[:td].add_(g.float().sum(0))
        for kind, g in (() if self.si == 1 else (("d", g_d), ("s", g_s))):
            fc1, fc2, ln = self.tok[kind]
            prev, t1, t2 = tp[kind]
            d_t2 = ops.layernorm_bwd(low(g, td), t2, ln[0], ln[2], ln[3])
            outs[0 if kind == "d" else 1] = self.lin_bwd(fc1, self.lin_bwd(fc2, d_t2, t1), prev)
        join_wgrads()
        if not keep_tape:
            self.tape = None
        return d_prev_x, outs[0], outs[1], d_feat


class DepthHead16(FlatModule):
    """`depth_pred16` = Linear(C + td -> 64), Linear(64 -> 1), Sigmoid on cat[x1, depth token] (multiscale_transformerr.py:
    1044-1045, 1213-1215): the coarse depth map of the 1/16 stage, in [0, 1] units"""

    def __init__(self, state_dict, name="depth_pred16", device="cuda", **optim):
        self.name = name
        keys = ["%s.%d.%s" % (name, i, wb) for i in (0, 1) for wb in ("weight", "bias")]
        super().__init__({k: state_dict[P0 + k] for k in keys}, None, device=device, **optim)
        self.l0 = Linear(self, name + ".0.weight", name + ".0.bias")
        self.l1 = Linear(self, name + ".1.weight", name + ".1.bias")
        self.tape = None

    def _weights(self):
        return [self.l0, self.l1]

    def state_dict(self):
        return super().state_dict(P0)

    def grads(self):
        return super().grads(P0)

    def forward(self, x, d):
        """x bf16 [rows, C], d bf16 [rows, td] -> depth fp32 [rows]"""
        cat = torch.cat([x, d], dim=1)
        t = conv_gemm(cat, self.l0.pw)
        y = conv_gemm(t, self.l1.pw, post_act=ops.ACT_SIGMOID, out_f32=True)             # fp32 [rows, 1]
        self.tape = (cat, t, y, x.shape[1])
        return y.view(-1)

    def backward(self, d_depth, keep_tape=False):
        """d_depth fp32 [rows] -> (d x [rows, C], d token [rows, td]) as views of one buffer"""
        cat, t, y, C = self.tape
        self.refresh_transposes()
        self.G.zero_()
        d_z = ops.act_bwd(d_depth.contiguous().view(-1, 1), y, ops.ACT_SIGMOID, out_cols=self.l1.n_pad)
        d_cat = self.lin_bwd(self.l0, self.lin_bwd(self.l1, d_z, t), cat)
        join_wgrads()
        if not keep_tape:
            self.tape = None
        return d_cat[:, :C], d_cat[:, C:cat.shape[1]]
