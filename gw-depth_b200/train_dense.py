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

from . import ops
from .engine import DEFAULT_CFG
from .ops import ACT_ELU, ACT_GELU, ACT_NONE, ACT_SIGMOID, conv_gemm
from .train_flat import join_wgrads, Conv3x3, FlatModule, Linear

PREFIX = "depth_decoder."
KINDS = ("depth", "seg")


class DenseHead(FlatModule):
    def __init__(self, state_dict, cfg=None, device="cuda", depth_weight=1.0, seg_weight=2.0, **optim):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.depth_weight, self.seg_weight = depth_weight, seg_weight
        c = self.cfg
        C, td = c["dense_trans_dim"] >> 3, c["class_token_dim"]
        self.C, self.td, self.width = C, td, C + 3 * td     # stage buffer: [x3 | depth token | seg token | depth_pred3 (+ pad)]
        feat, dtok, stok = torch.arange(C), torch.arange(td) + C, torch.arange(td) + C + td
        layout = {"depth_token_fuse.fc1.weight": dict(cin_pad=self.width, col_map=torch.cat([feat, torch.tensor([C + 2 * td]), dtok]),
                                                     shared_input=True),
                  "seg_token_fuse.fc1.weight": dict(cin_pad=self.width, col_map=torch.cat([feat, stok]), shared_input=True)}
        tensors = {k[len(PREFIX):]: v for k, v in state_dict.items() if k.startswith(PREFIX) and v.is_floating_point()}
        super().__init__(tensors, layout, device=device, **optim)
        self.losses = torch.zeros(2, dtype=torch.float32, device=self.dev)     # weighted (depth, seg) loss of the last step
        self.mods = {}
        for kind in KINDS:
            m = {"fc1": Linear(self, kind + "_token_fuse.fc1.weight", kind + "_token_fuse.fc1.bias"),
                 "fc2": Linear(self, kind + "_token_fuse.fc2.weight", kind + "_token_fuse.fc2.bias")}
            for cv in ("upconv1_%s.conv", "conv1_%s.0", "upconv2_%s.conv", "conv2_%s.0"):
                m[cv.split("_")[0]] = Conv3x3(self, (cv % kind) + ".weight")
            m["norm"] = self.ln("norm_" + kind)
            self.mods[kind] = m
        self.mods["depth"]["last"] = Conv3x3(self, "get_depth.0.weight")
        self.mods["seg"]["last"] = Conv3x3(self, "get_seg.weight")
        self.tape = None

    def _weights(self):
        return [v for m in self.mods.values() for v in m.values() if isinstance(v, (Conv3x3, Linear))]

    def state_dict(self):
        return super().state_dict(PREFIX)

    def grads(self):
        return super().grads(PREFIX)

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
            d = self.conv_bwd(m["last"], img(g, H, W), s["c2"])
            d = self.conv_bwd(m["conv2"], img(ops.act_bwd(d, s["c2"], ACT_ELU), H, W), s["u2"])
            d = self.conv_bwd(m["upconv2"], img(ops.act_bwd(d, s["u2"], ACT_ELU), H, W), s["c1u"])
            # nearest x2 up-sampling backward: every low-resolution pixel collects its 2x2 block (mean * 4)
            d = ops.act_bwd(ops.avgpool(d, 2), s["c1"], ACT_ELU, scale=4.0)
            d = self.conv_bwd(m["conv1"], img(d, H // 2, W // 2), s["u1"])
            gam, _, dgam, dbet = m["norm"]
            d = ops.layernorm_bwd(d, s["e1"], gam, dgam, dbet)
            d = self.conv_bwd(m["upconv1"], img(ops.act_bwd(d, s["e1"], ACT_ELU), H // 2, W // 2), s["fu"])
            d = ops.act_bwd(ops.avgpool(d, 2), None, ACT_NONE, scale=4.0)
            d = self.lin_bwd(m["fc2"], d, s["t"])
            d = ops.act_bwd(d, s["h_raw"], ACT_GELU, from_input=True)
            d_x = self.lin_bwd(m["fc1"], d, tp["x"], res=d_x)
        join_wgrads()
        self.mask_grads()
        if not keep_tape:
            self.tape = None
        return d_x

    # ------------------------------------------------------------------ losses
    def loss_grads(self, depth, seg, depth_gt, seg_gt):
        """SilogLoss (last scale, weight 1) + SegLoss x 2 on the forward outputs: fills self.losses (weighted (loss_depth,
        loss_seg)) and returns the bf16 [B*H*W, 16] gradient rows `backward` takes"""
        log_only = bool(self.cfg.get("log_depth_error", False))
        sums = ops.silog_sums(depth, depth_gt, log_only=log_only)
        g_depth = ops.silog_bwd(depth, depth_gt, sums, weight=self.depth_weight, log_only=log_only,
                                variance_focus=float(self.cfg.get("variance_focus", 0.85)), sig_scale=float(self.cfg["max_depth"]),
                                out_cols=16, loss_out=self.losses[0:1])
        _, g_seg = ops.seg_ce(seg, seg_gt.contiguous(), weight=self.seg_weight, out_cols=16, loss_out=self.losses[1:2])
        return g_depth, g_seg

    def cotangent_rows(self, depth, d_depth, d_seg):
        """external cotangents (the reference engine's autograd: d depth fp32 [B,1,H,W] in metres, d seg fp32 [B,2,H,W]) -> the
        bf16 [B*H*W, 16] gradient rows `backward` takes (sigmoid * max_depth derivative from the kept output)"""
        md = float(self.cfg["max_depth"])
        g_depth = ops.act_bwd(d_depth.reshape(-1, 1).float().contiguous(), depth.reshape(-1, 1), ACT_SIGMOID, out_cols=16,
                              y_mul=1.0 / md, scale=md)
        g_seg = ops.act_bwd(d_seg.permute(0, 2, 3, 1).reshape(-1, d_seg.shape[1]).float().contiguous(), None, ACT_NONE, out_cols=16)
        return g_depth, g_seg

    def loss_and_grads(self, buf4, depth_gt, seg_gt, H=None, W=None):
        """forward + SilogLoss (last scale) + SegLoss + backward.  depth_gt fp32 [B,1,H,W] metres, seg_gt int64 [B,1,H,W].
        Returns (depth, seg, losses fp32 [2] on the device = weighted (loss_depth, loss_seg), d(buf4))."""
        H, W = H or depth_gt.shape[-2], W or depth_gt.shape[-1]
        depth, seg = self.forward(buf4, H, W)
        g_depth, g_seg = self.loss_grads(depth, seg, depth_gt, seg_gt)
        d_buf4 = self.backward(g_depth, g_seg)
        return depth, seg, self.losses, d_buf4

    def train_step(self, buf4, depth_gt, seg_gt):
        _, _, losses, _ = self.loss_and_grads(buf4, depth_gt, seg_gt)
        self.step()
        return losses
