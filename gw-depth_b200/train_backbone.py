"""Training path of the BACKBONE on B200: torchvision's ResNet-50 v1.5 with FrozenBatchNorm2d as the
reference wraps it (src/models/backbone.py:19-92): the stem and layer1 are frozen (:62-64; they run as the inference kernels),
layer2 / layer3 / layer4 (13 bottlenecks, 23.5 M parameters) are trained with lr_backbone (src/main_glassrgbd.py:59-66).

B200 design
* every convolution of the trained layers lives in ONE flat fp32 buffer (train_flat.FlatModule) in the layout the tcgen05 GEMM
  reads: 1x1 convolutions as [N, C], stride-1 3x3 as [tap][N][C], the three STRIDE-2 3x3 convolutions as [N, 9C] -- they run as
  gwd_im2col3x3_s2 + one Linear (exactly the stride-2 FLOPs; their data gradient is a Linear + gwd_col2im3x3_s2, their weight
  gradient a Linear weight gradient), the stride-2 1x1 projections read every other pixel through the strided TMA view
  (forward) / gwd_subsample2 (weight gradient) and spread their data gradient with gwd_zero_stuff2;
* FrozenBatchNorm2d (eps inside the rsqrt, :46-55) is folded: the bf16 mirror the kernels read is w * s[n] (gwd_fold_mirror
  after every optimizer step), the shift is the GEMM bias, and the kernels' gradient w.r.t. the folded filter is multiplied by
  s[n] (gwd_scale_rows) to give the gradient of the parameter itself;
* a bottleneck's backward: ReLU' from the kept outputs (gwd_act_bwd), data gradients on gwd_conv_gemm with the transposed
  / flipped mirrors (the identity shortcut's gradient is added in the epilogue of conv1's data-gradient GEMM), weight
  gradients on gwd_linear_wgrad / gwd_conv3x3_wgrad (tcgen05 for the wide ones).
"""
import torch

from . import ops
from .ops import ACT_RELU, RES_AFTER, RES_BEFORE_NORM, RES_NONE, PackedWeight, conv_gemm, pack_conv3x3, pack_linear
from .train_flat import FUSE_ACT_GRAD, join_wgrads, Conv3x3, FlatModule, Linear

BODY = "backbone.0.body."
LAYERS = ((1, 3), (2, 4), (3, 6), (4, 3))


def _bn(sd, name, dev):
    """FrozenBatchNorm2d as (scale, shift), src/models/backbone.py:46-55"""
    g = lambda k: sd[name + "." + k].detach().to(dev, torch.float32)  # noqa: E731
    scale = g("weight") * (g("running_var") + 1e-5).rsqrt()
    return scale, g("bias") - g("running_mean") * scale


class BackboneTrain(FlatModule):
    def __init__(self, state_dict, device="cuda", **optim):
        sd = state_dict
        dev = torch.device(device)
        tensors, self._s2, bn_of = {}, set(), {}
        for li, nb in LAYERS[1:]:
            for bi in range(nb):
                q = "layer%d.%d." % (li, bi)
                for conv, bn in (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"), ("downsample.0", "downsample.1")):
                    k = BODY + q + conv + ".weight"
                    if k not in sd:
                        continue
                    v = sd[k].detach().float()
                    if conv == "conv2" and bi == 0:          # stride 2: [N, C, ky, kx] -> [N, (ky, kx, c)]
                        self._s2.add(q + conv + ".weight")
                        v = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
                    tensors[q + conv + ".weight"] = v
                    bn_of[q + conv + ".weight"] = BODY + q + bn
        super().__init__(tensors, None, device=device, **optim)
        # folded FrozenBatchNorm: flat per-element scale (the output channel's s[n]) + per-convolution shift
        self.S = torch.ones_like(self.P)
        self.shift = {}
        for short, bn in bn_of.items():
            scale, shift = _bn(sd, bn, dev)
            sv = self.view(self.S, short)
            if sv.dim() == 3:
                sv.copy_(scale.view(1, -1, 1).expand_as(sv))
            else:
                sv.copy_(scale.view(-1, 1).expand_as(sv))
            self.shift[short] = shift.contiguous()
        self.refresh_mirror()
        self.blocks = []
        for li, nb in LAYERS[1:]:
            stage = []
            for bi in range(nb):
                q = "layer%d.%d." % (li, bi)
                blk = {"stride": 2 if bi == 0 else 1, "c1": self._lin(q + "conv1.weight"), "c3": self._lin(q + "conv3.weight")}
                if bi == 0:
                    blk["c2"] = self._lin(q + "conv2.weight")
                    blk["down"] = self._lin(q + "downsample.0.weight")
                else:
                    cv = Conv3x3(self, q + "conv2.weight")
                    cv.pw = PackedWeight(cv.wb, self.shift[q + "conv2.weight"], 9, cv.n, cv.c_pad)
                    blk["c2"] = cv
                stage.append(blk)
            self.blocks.append(stage)
        self._pack_frozen(sd, dev)
        self.tape = None

    def _lin(self, short):
        lin = Linear(self, short, None)
        lin.pw = PackedWeight(lin.wb.view(1, lin.n_pad, lin.k), self.shift[short], 1, lin.n, lin.k)
        return lin

    def _pack_frozen(self, sd, dev):
        """stem + layer1 (requires_grad False in the reference, backbone.py:62-64): the inference packing of engine.Engine"""
        g = lambda k: sd[BODY + k].detach().to(dev, torch.float32)  # noqa: E731
        scale, shift = _bn(sd, BODY + "bn1", dev)
        self.stem = ops.pack_stem(g("conv1.weight") * scale.view(-1, 1, 1, 1), shift)
        self.layer1 = []
        for bi in range(LAYERS[0][1]):
            q = "layer1.%d." % bi

            def fold(conv, bn, taps):
                s, sh = _bn(sd, BODY + q + bn, dev)
                w = g(q + conv + ".weight") * s.view(-1, 1, 1, 1)
                return pack_linear(w.flatten(1), sh) if taps == 1 else pack_conv3x3(w, sh)
            blk = {"c1": fold("conv1", "bn1", 1), "c2": fold("conv2", "bn2", 9), "c3": fold("conv3", "bn3", 1)}
            if (BODY + q + "downsample.0.weight") in sd:
                blk["down"] = fold("downsample.0", "downsample.1", 1)
            self.layer1.append(blk)

    def _weights(self):
        return [b[k] for st in self.blocks for b in st for k in ("c1", "c2", "c3", "down") if k in b]

    def _names(self, d):
        out = {}
        for k, v in d.items():
            if k in self._s2:
                n = v.shape[0]
                v = v.view(n, 3, 3, -1).permute(0, 3, 1, 2).contiguous()
            out[BODY + k] = v
        return out

    def _adapt(self, short, v):
        if short in self._s2:
            return v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
        return v.reshape(self.index[short][2])

    def state_dict(self):
        return self._names(super().state_dict(""))

    def grads(self):
        return self._names(super().grads(""))

    # ------------------------------------------------------------------ folded FrozenBatchNorm
    def refresh_mirror(self):
        """bf16 mirror = parameter * s[n] (after construction and after every optimizer step)"""
        ops.fold_mirror(self.P, self.S, self.Wb)

    def finish_grads(self):
        """gradient w.r.t. the folded filter -> gradient w.r.t. the parameter (call once, after backward)"""
        ops.scale_rows(self.G, self.S)

    def step(self, sumsq=None, reduced=False):
        world = self._world if reduced else self.allreduce_grads()
        self.t += 1
        if sumsq is None:
            self.sumsq.zero_()
            ops.sumsq(self.G, self.sumsq)
            sumsq = self.sumsq
        ops.adamw_step(self.P, self.G, self.M, self.V, None, lr=self.lr, betas=self.betas, eps=self.eps,
                       weight_decay=self.weight_decay, step=self.t, max_norm=self.max_norm, grad_scale=1.0 / world,
                       sumsq_buf=sumsq)
        self.refresh_mirror()

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def frozen_front(self, images):
        """fp32 [B,3,H,W] -> C2 bf16 [B,H/4,W/4,256] (stem + layer1, no tape)"""
        x = ops.stem_conv_pool(images, *self.stem)
        for blk in self.layer1:
            y = conv_gemm(x, blk["c1"], post_act=ACT_RELU)
            y = conv_gemm(y, blk["c2"], post_act=ACT_RELU)
            idt = conv_gemm(x, blk["down"]) if "down" in blk else x
            x = conv_gemm(y, blk["c3"], res=idt, res_mode=RES_BEFORE_NORM, post_act=ACT_RELU)
        return x

    def forward(self, c2):
        """c2 bf16 [B,h,w,256] -> [C3, C4, C5] bf16 channels-last; keeps the tape"""
        x = c2
        feats, tape = [], []
        for stage in self.blocks:
            for blk in stage:
                y1 = conv_gemm(x, blk["c1"].pw, post_act=ACT_RELU)
                if blk["stride"] == 2:
                    col = ops.im2col3x3_s2(y1)
                    B, ho, wo, _ = col.shape
                    y2 = conv_gemm(col.view(B * ho * wo, -1), blk["c2"].pw, post_act=ACT_RELU).view(B, ho, wo, -1)
                    idt = conv_gemm(x, blk["down"].pw, subsample2=True)
                else:
                    col = None
                    y2 = conv_gemm(y1, blk["c2"].pw, post_act=ACT_RELU)
                    idt = x
                out = conv_gemm(y2, blk["c3"].pw, res=idt, res_mode=RES_BEFORE_NORM, post_act=ACT_RELU)
                tape.append((x, y1, col, y2, out))
                x = out
            feats.append(x)
        self.tape = tape
        return feats

    # ------------------------------------------------------------------ backward
    def backward(self, g_c3, g_c4, g_c5, keep_tape=False):
        """gradients of (C3, C4, C5) (bf16 channels-last, same shapes; None = zero) -> fills the flat gradient buffer (w.r.t. the
        PARAMETERS: the FrozenBatchNorm scale is applied).  C2 comes from the frozen layer1: nothing is returned."""
        self.refresh_transposes()
        self.G.zero_()
        it = iter(reversed(self.tape))
        g = None
        first = self.blocks[0][0]
        for stage, g_feat in zip(reversed(self.blocks), (g_c5, g_c4, g_c3)):
            if g_feat is not None:
                g = g_feat if g is None else ops.add_rows(g.view(-1, g.shape[-1]), g_feat.reshape(-1, g.shape[-1]), g.numel() // g.shape[-1]).view_as(g)
            for blk in reversed(stage):
                x, y1, col, y2, out = next(it)
                B, H, W, _ = x.shape
                ho, wo = out.shape[1:3]
                rows_o = B * ho * wo
                gs = ops.act_bwd(g.reshape(rows_o, -1), out.view(rows_o, -1), ACT_RELU)                    # d(sum) = d out * relu'
                if FUSE_ACT_GRAD:      # relu' of the conv2 output in the epilogue of conv3's data-gradient GEMM
                    d_y2 = self.lin_bwd(blk["c3"], gs, y2.view(rows_o, -1), act_grad=(y2.view(rows_o, -1), ACT_RELU, False, 1.0, 1.0))
                else:
                    d_y2 = ops.act_bwd(self.lin_bwd(blk["c3"], gs, y2.view(rows_o, -1)), y2.view(rows_o, -1), ACT_RELU)
                if blk["stride"] == 2:
                    d_col = self.lin_bwd(blk["c2"], d_y2, col.view(rows_o, -1))
                    d_y1 = ops.col2im3x3_s2(d_col.view(B, ho, wo, -1), H, W)
                    d_y1 = ops.act_bwd(d_y1.view(B * H * W, -1), y1.view(B * H * W, -1), ACT_RELU)
                elif FUSE_ACT_GRAD:
                    d_y1 = self.conv_bwd(blk["c2"], d_y2.view(B, ho, wo, -1), y1, act_grad=(y1, ACT_RELU, False, 1.0, 1.0)).view(B * H * W, -1)
                else:
                    d_y1 = ops.act_bwd(self.conv_bwd(blk["c2"], d_y2.view(B, ho, wo, -1), y1).view(B * H * W, -1), y1.view(B * H * W, -1), ACT_RELU)
                need_dx = blk is not first
                if blk["stride"] == 2:
                    d_x = self.lin_bwd(blk["c1"], d_y1, x.view(B * H * W, -1), need_dx=need_dx)
                    d_sub = self.lin_bwd(blk["down"], gs, ops.subsample2(x).view(rows_o, -1), need_dx=need_dx)
                    if need_dx:
                        d_x = ops.zero_stuff2(d_sub.view(B, ho, wo, -1), H, W, add=d_x.view(B, H, W, -1))
                else:
                    d_x = self.lin_bwd(blk["c1"], d_y1, x.view(B * H * W, -1), res=gs)                       # + identity shortcut
                g = d_x.view(B, H, W, -1) if d_x is not None else None
        join_wgrads()
        self.finish_grads()
        if not keep_tape:
            self.tape = None
