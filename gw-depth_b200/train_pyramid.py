"""Training path of the PYRAMID LAYER of PointBasedPred on B200 -- the block that carries most of the model's FLOPs
(188 of 359 GFLOP per image at 480x640): forward with the pre-norm values kept, backward through hand-written kernels.

Reference: `PyramidLayer.forward`, `ConvLn`, `BasicBlock` (src/models/points/points_sample.py:12-43,106-125) under
torch.autograd: two ConvLn + GELU, five residual blocks x' = ConvLn(GELU(ConvLn(x))) + x, four branches
AvgPool(16/8/4/2) -> ConvLn -> GELU -> bilinear up-sampling (align_corners=True), concat, ConvLn(10K -> 4K) + GELU, 1x1
conv to the K mixture logits.  `layer4` is constructed by the reference but never run: it has no gradient and is not
stored here.

B200 design (see train_flat.FlatModule for the parameter storage)
* forward = the inference kernel sequence of engine.Engine.pyramid: 3x3 conv + LayerNorm + GELU (+ residual) in ONE
  tcgen05 GEMM launch each, the pre-norm value kept by the epilogue (y_raw); the five maps of the concat are channel
  slices of one buffer;
* backward per ConvLn: gwd_layernorm_bwd differentiates GELU(LN(z)) from z in one pass (statistics over the logical
  channels when the width is padded to 16), gwd_conv3x3_wgrad accumulates dW, the data gradient is gwd_conv_gemm with the
  flipped / transposed mirror and the residual gradient added in its epilogue;
* branches: gwd_bilinear_up_bwd gathers each branch's slice of the concat gradient with the forward's footprint
  arithmetic; gwd_avgpool_bwd spreads the pooled gradient and accumulates the four branches + the identity slice.
"""
import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, RES_AFTER, RES_NONE, conv_gemm, round_up
from .train_flat import join_wgrads, Conv3x3, FlatModule, Linear

POOLS = (16, 8, 4, 2)


class Pyramid(FlatModule):
    def __init__(self, state_dict, prefix, K, device="cuda", **optim):
        """prefix: e.g. 'dense_encoder.point_based_pred2.pyramid.'; K = mixture components (interval_sample_num)"""
        self.prefix, self.K = prefix, K
        self.Kp, self.C2, self.C2p, self.C4 = round_up(K, 16), 2 * K, round_up(2 * K, 16), 4 * K
        tensors = {k[len(prefix):]: v for k, v in state_dict.items()
                   if k.startswith(prefix) and v.is_floating_point() and not k[len(prefix):].startswith("layer4.")}
        cmap = torch.cat([torch.arange(self.C2) + g * self.C2p for g in range(5)])
        super().__init__(tensors, {"lastconv.0.conv.weight": dict(cin_pad=5 * self.C2p, col_map=cmap)}, device=device, **optim)
        cl = lambda name: (Conv3x3(self, name + ".conv.weight"), self.ln(name + ".layer_norm"))
        self.first = [cl("firstconv.0"), cl("firstconv.2")]
        self.blocks = [(cl("%s.%d.conv1.0" % (ln, b)), cl("%s.%d.conv2" % (ln, b)))
                       for ln, nb in (("layer1", 1), ("layer2", 2), ("layer3", 2)) for b in range(nb)]
        self.branches = [cl("branch%d.1" % i) for i in range(1, 5)]
        self.last0 = cl("lastconv.0")
        self.last2 = Linear(self, "lastconv.2.weight")
        self.tape = None

    def _weights(self):
        cls = self.first + [c for blk in self.blocks for c in blk] + self.branches + [self.last0]
        return [c[0] for c in cls] + [self.last2]

    def state_dict(self):
        return super().state_dict(self.prefix)

    def grads(self):
        return super().grads(self.prefix)

    # ------------------------------------------------------------------ forward
    def _conv_ln(self, x, cvln, act, res=None, out=None):
        cv, ln = cvln
        z = torch.empty(x.shape[:3] + (cv.n_pad,), dtype=torch.bfloat16, device=self.dev)
        y = conv_gemm(x, cv.pw, bias=False, ln=(ln[0], ln[1]), post_act=act, res=res,
                      res_mode=RES_AFTER if res is not None else RES_NONE, out=out, y_raw=z)
        return y, z

    def forward(self, rg):
        """rg: bf16 [B,H,W,Kp] (zero padding channels) -> mixture logits bf16 [B,H,W,Kp]"""
        B, H, W, Kp = rg.shape
        assert Kp == self.Kp and rg.dtype == torch.bfloat16 and rg.is_contiguous()
        if H < POOLS[0] or W < POOLS[0]:
            raise NotImplementedError("pad_before_pool path (feature map smaller than the 16-pixel pool) is not built")
        C2p = self.C2p
        tp = self.tape = {"shape": (B, H, W), "rg": rg}
        x0, tp["z_f0"] = self._conv_ln(rg, self.first[0], ACT_GELU)
        x, tp["z_f2"] = self._conv_ln(x0, self.first[1], ACT_GELU)
        tp["x_f0"] = x0
        cat = torch.empty(B, H, W, 5 * C2p, dtype=torch.bfloat16, device=self.dev)
        tp["blocks"] = []
        for i, (c1, c2) in enumerate(self.blocks):
            y, z1 = self._conv_ln(x, c1, ACT_GELU)
            xn, z2 = self._conv_ln(y, c2, ACT_NONE, res=x, out=cat if i == len(self.blocks) - 1 else None)
            tp["blocks"].append((x, z1, y, z2))
            x = xn
        tp["branches"] = []
        ups = []
        for pooled, br in zip(ops.avgpool_pyramid(cat, C=C2p), self.branches):        # POOLS order (16, 8, 4, 2), one pass
            b, zb = self._conv_ln(pooled, br, ACT_GELU)
            ups.append(b)
            tp["branches"].append((pooled, zb))
        ops.bilinear_up4_into(ups, cat, C2p, H, W)
        cv, ln = self.last0
        if cv.n_pad <= 256:
            y, z = self._conv_ln(cat, self.last0, ACT_GELU)
        else:   # LayerNorm over more channels than one accumulator tile: separate pass
            z = conv_gemm(cat, cv.pw, bias=False)
            y = ops.layernorm(z, ln[0], ln[1], act=ACT_GELU, n=cv.n)
        tp.update(cat=cat, z_last=z, y_last=y)
        return conv_gemm(y, self.last2.pw, bias=False, out_channels=self.Kp)

    # ------------------------------------------------------------------ backward
    def _conv_ln_bwd(self, d, z, cvln, x, act, res=None, need_dx=True):
        cv, ln = cvln
        dz = ops.layernorm_bwd(d, z, ln[0], ln[2], ln[3], beta=ln[1], post_act=act, n=cv.n).view(z.shape)
        return self.conv_bwd(cv, dz, x, need_dx=need_dx, res=res)

    def backward(self, d_logits, keep_tape=False, need_dx=True):
        """d_logits: bf16 [B,H,W,Kp] (zero padding channels).  Fills the flat gradient buffer; returns d(rg) [B,H,W,Kp]"""
        tp = self.tape
        B, H, W = tp["shape"]
        C2p, rows = self.C2p, B * H * W
        self.refresh_transposes()
        self.G.zero_()
        d = self.lin_bwd(self.last2, d_logits.view(rows, -1), tp["y_last"].view(rows, -1))
        d_cat = self._conv_ln_bwd(d, tp["z_last"], self.last0, tp["cat"], ACT_GELU)           # [B,H,W,5*C2p]
        acc = None
        for j, (pool, br) in enumerate(zip(POOLS, self.branches), start=1):
            pooled, zb = tp["branches"][j - 1]
            h, w = pooled.shape[1:3]
            d_b = ops.bilinear_up_bwd(d_cat[..., j * C2p:(j + 1) * C2p], h, w)
            d_pool = self._conv_ln_bwd(d_b, zb, br, pooled, ACT_GELU)
            if acc is None:     # identity slice of the concat + the first branch
                acc = ops.avgpool_bwd(d_pool, pool, H, W, add=d_cat[..., :C2p])
            else:
                ops.avgpool_bwd(d_pool, pool, H, W, add=acc, out=acc)
        d = acc
        for (c1, c2), (x, z1, y, z2) in zip(reversed(self.blocks), reversed(tp["blocks"])):
            dy = self._conv_ln_bwd(d, z2, c2, y, ACT_NONE)
            d = self._conv_ln_bwd(dy, z1, c1, x, ACT_GELU, res=d)
        d = self._conv_ln_bwd(d, tp["z_f2"], self.first[1], tp["x_f0"], ACT_GELU)
        d = self._conv_ln_bwd(d, tp["z_f0"], self.first[0], tp["rg"], ACT_GELU, need_dx=need_dx)
        join_wgrads()
        if not keep_tape:
            self.tape = None
        return d
