"""Training path of POINT-BASED DEPTH PREDICTION on B200: the block that turns the stage features
into a depth map as a soft-max mixture of K anchor depths sampled at the uncertainty points, forward with the tape kept
and the backward on hand-written kernels.  It wraps train_pyramid.Pyramid (row A19), which carries most of its FLOPs.

Reference: `PointBasedPred.forward` (src/models/points/points_sample.py:257-280) under torch.autograd:
    x_global = pre_proj(cat[x, depth_token]);  xg | xr = refer_proj(x_global)
    refer    = grid_sample(xr, coords) + grid_sample(pos, coords)            [B, dim, K]   (bilinear, align_corners=False)
    anchor   = grid_sample(pre_depth, coords)                               [B, K]
    rg       = xg @ refer * dim**-2  ->  PyramidLayer  ->  soft-max over K  ->  sum_k attn_k * anchor_k
The sample coordinates come from CertainSample (a top-k selection): they carry no gradient.

B200 design
* parameters: `pre_proj` / `refer_proj` in one flat fp32 master buffer (train_flat.FlatModule) in the layout the tcgen05
  GEMM reads -- `pre_proj` reads the stage buffer [x | depth token | seg token | ...] in place through a column map, so the
  concat is never materialised; the pyramid keeps its own flat buffers; `step` clips over both with ONE global norm;
* forward = the inference kernel sequence of engine.Engine.point_based_pred with the two Linears kept apart (their
  product is folded at inference; training needs both factors) and `x_global`, `xg | xr`, the per-image correlation
  weights, the anchors and the mixture logits kept;
* backward: gwd_anchor_mix_bwd (soft-max mixture: d logits as padded bf16 rows + d anchor by shuffle / shared-memory
  reduction) -> Pyramid.backward -> the correlation's two gradients: d xg = d rg @ refer on the per-image-weight tcgen05
  GEMM (written into the left half of the d(xg | xr) buffer), d refer = d rg^T @ xg per image on gwd_linear_wgrad ->
  gwd_sample_bilinear_bwd gathers the K point gradients into the right half (deterministic, zeros elsewhere) ->
  Linear dgrad / wgrad for `refer_proj`, `pre_proj`; gwd_sample_scalar_bwd spreads d anchor over the previous scale's depth
  map (added to that map's own loss gradient).
"""
import torch

from . import ops
from .ops import PackedWeight, conv_gemm, round_up
from .train_flat import join_wgrads, FlatModule, Linear
from .train_pyramid import Pyramid


class PointPred(FlatModule):
    def __init__(self, state_dict, prefix, dim, token_dim, K, in_width=None, device="cuda", **optim):
        """prefix: e.g. 'dense_encoder.point_based_pred1.'; dim: stage channels (D/4, D/8); token_dim: depth-token channels;
        K: sample points (interval_sample_num); in_width: channel count of the stage buffer whose first dim + token_dim
        columns are [x | depth token] (default: exactly those columns, padded to 16)"""
        self.prefix, self.dim, self.td, self.K = prefix, dim, token_dim, K
        self.Kp = round_up(K, 16)
        self.in_width = in_width or round_up(dim + token_dim, 16)
        assert dim % 16 == 0 and self.in_width % 16 == 0 and self.in_width >= dim + token_dim
        names = ("pre_proj.weight", "pre_proj.bias", "refer_proj.weight", "refer_proj.bias")
        tensors = {n: state_dict[prefix + n] for n in names}
        layout = {"pre_proj.weight": dict(cin_pad=self.in_width, col_map=torch.arange(dim + token_dim), shared_input=True)}
        super().__init__(tensors, layout, device=device, **optim)
        self.pre = Linear(self, "pre_proj.weight", "pre_proj.bias")
        self.refer = Linear(self, "refer_proj.weight", "refer_proj.bias")
        self.pyramid = Pyramid(state_dict, prefix + "pyramid.", K, device=device, **optim)
        self.tape = None

    def _weights(self):
        return [self.pre, self.refer]

    def state_dict(self):
        sd = super().state_dict(self.prefix)
        sd.update(self.pyramid.state_dict())
        return sd

    def grads(self):
        g = super().grads(self.prefix)
        g.update(self.pyramid.grads())
        return g

    def load_params(self, state_dict):
        super().load_params(state_dict)
        self.pyramid.load_params(state_dict)

    # ------------------------------------------------------------------ forward
    def forward(self, buf, pre_depth, coords, pos, B, H, W):
        """buf: bf16 [B*H*W, in_width] stage buffer; pre_depth: fp32 [B,h,w] previous scale's depth; coords: fp32 [B,K,2]
        (x, y) in [-1,1]; pos: fp32 position table [H*W, dim] or [B, H*W, dim].  Returns the depth map fp32 [B,H,W]."""
        dim, K, Kp = self.dim, self.K, self.Kp
        HW = H * W
        assert buf.shape == (B * HW, self.in_width) and buf.dtype == torch.bfloat16 and buf.is_contiguous()
        assert coords.shape == (B, K, 2) and coords.dtype == torch.float32 and coords.is_contiguous()
        pre_depth = pre_depth.contiguous()
        g = conv_gemm(buf, self.pre.pw, out_channels=dim)                               # x_global        [rows, dim]
        pr = conv_gemm(g, self.refer.pw, out_channels=2 * dim)                          # xg | xr         [rows, 2 dim]
        refer = ops.sample_bilinear(pr, dim, pos, B, H, W, dim, coords, K)              # fp32 [B,K,dim]
        anchor = ops.sample_scalar(pre_depth, coords, K)                                # fp32 [B,K]
        wimg = torch.zeros(B, Kp, dim, dtype=torch.bfloat16, device=self.dev)
        wimg[:, :K] = refer * (dim ** -2)                                               # points_sample.py:273
        rg = conv_gemm(pr.view(B, 1, HW, 2 * dim), PackedWeight(wimg, None, 1, K, dim), bias=False, w_per_image=True,
                       out_channels=Kp)
        logits = self.pyramid.forward(rg.view(B, H, W, Kp))
        self.tape = dict(B=B, H=H, W=W, buf=buf, g=g, pr=pr, wimg=wimg, anchor=anchor, logits=logits, coords=coords,
                         pre_hw=tuple(pre_depth.shape[1:]))
        return ops.anchor_mix(logits.view(B * HW, Kp), anchor, B, HW, K).view(B, H, W)

    # ------------------------------------------------------------------ backward
    def backward(self, d_pred, d_pre_depth=None, keep_tape=False):
        """d_pred: fp32 [B,H,W] gradient of the returned depth map; d_pre_depth: optional fp32 [B,h,w] gradient the previous
        scale's depth map already carries (its own loss).  Fills the flat gradient buffers (this module's and the
        pyramid's); returns (d(buf) bf16 [B*H*W, in_width], d(pre_depth) fp32 [B,h,w])."""
        tp = self.tape
        B, H, W = tp["B"], tp["H"], tp["W"]
        dim, K, Kp, HW = self.dim, self.K, self.Kp, tp["H"] * tp["W"]
        d_logits, d_anchor = ops.anchor_mix_bwd(tp["logits"].view(B * HW, Kp), tp["anchor"], d_pred.contiguous().float(), B, HW, K)
        d_rg = self.pyramid.backward(d_logits.view(B, H, W, Kp), keep_tape=keep_tape).view(B, HW, Kp)
        # correlation rg[b,p,k] = sum_c xg[b,p,c] wimg[b,k,c]
        d_pr = torch.empty(B * HW, 2 * dim, dtype=torch.bfloat16, device=self.dev)
        wT = tp["wimg"].transpose(1, 2).contiguous()                                     # [B, dim, Kp]
        conv_gemm(d_rg.view(B, 1, HW, Kp), PackedWeight(wT, None, 1, dim, Kp), bias=False, w_per_image=True,
                  out=d_pr.view(B, 1, HW, 2 * dim), y_coff=0)
        d_wimg = torch.zeros(B, Kp, dim, dtype=torch.float32, device=self.dev)
        pr3 = tp["pr"].view(B, HW, 2 * dim)
        for b in range(B):
            ops.linear_wgrad(d_rg[b], pr3[b], d_wimg[b], K=dim)
        d_refer = (d_wimg[:, :K] * (dim ** -2)).contiguous()
        ops.sample_bilinear_bwd(d_refer, tp["coords"], d_pr, dim, H, W)                  # d xr: right half of d_pr
        h, w = tp["pre_hw"]
        d_pre = ops.sample_scalar_bwd(d_anchor, tp["coords"], h, w, add=d_pre_depth)
        self.refresh_transposes()
        self.G.zero_()
        d_g = self.lin_bwd(self.refer, d_pr, tp["g"])
        d_buf = self.lin_bwd(self.pre, d_g, tp["buf"])
        join_wgrads()
        self.mask_grads()
        if not keep_tape:
            self.tape = None
        return d_buf, d_pre

    # ------------------------------------------------------------------ optimizer
    def step(self, sumsq=None):
        """gradient all-reduce of both flat buffers, clip over this module AND its pyramid with one norm (the reference clips
        the whole model, src/engine_glassrgbd.py:155-159), then AdamW on both"""
        self.allreduce_grads()
        self.pyramid.allreduce_grads()
        if sumsq is None:
            self.sumsq.zero_()
            ops.sumsq(self.G, self.sumsq)
            ops.sumsq(self.pyramid.G, self.sumsq)
            sumsq = self.sumsq
        super().step(sumsq, reduced=True)
        self.pyramid.step(sumsq, reduced=True)
