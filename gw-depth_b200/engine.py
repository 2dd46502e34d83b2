"""GW-Depth forward on B200: the execution plan behind `GlassRGBD.forward`.

Weights are re-laid-out ONCE (bf16, K-major tap-packed, scales / back-to-back Linears / frozen batch-norm folded)
and the forward is a fixed sequence of C-ABI kernel launches (gw-depth_b200/ops.py -> libgwd_b200.so) on bf16
channels-last activations with fp32 accumulation, fp32 statistics and fp32 selection inputs (line logits, coarse
depth maps).  There are no host synchronisations and no data-dependent Python control flow in the forward, so
it can be captured in a CUDA graph.

What still runs through PyTorch library calls (cuDNN / ATen) and is therefore NOT claimed as a hand-written
kernel: the six stride-2 convolutions of the ResNet-50 backbone (with their ReLU) and three tiny index ops (top-k
over 100 line logits, a gather of 20 lines, parameter broadcasts).

Reference call sites are cited per method (paths relative to the reference root).
"""
import math

import torch
import torch.nn.functional as F

from . import ops
from .ops import (ACT_ELU, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, RES_AFTER, RES_BEFORE_NORM, conv_gemm, pack_conv3x3,
                  pack_linear, pad_vec, round_up)

DEFAULT_CFG = dict(
    hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6, num_queries=100, dense_trans_dim=512, dense_trans_heads=16,
    dense_trans_layers=(4,), class_trans_layers=(2, 2, 1), class_token_dim=64, num_ref=20, with_dense_center=False,
    window=7, interval_sample_num=(30, 80), depth_interval=(0.1, 0.3, 0.5, 0.7, 0.9), min_depth_eval=1e-3,
    max_depth_eval=10.0, max_depth=10.0, aux_loss=True)


def sine_table(h, w, num_pos_feats, normalize, device):
    """PositionEmbeddingSine for an un-padded map (src/models/position_encoding.py:28-48) as an [h*w, 2*num_pos_feats]
    fp32 table (channel order: y block then x block, sin/cos interleaved)."""
    ys = torch.arange(1, h + 1, dtype=torch.float32, device=device)
    xs = torch.arange(1, w + 1, dtype=torch.float32, device=device)
    if normalize:
        ys = ys / (h + 1e-6) * (2 * math.pi)
        xs = xs / (w + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=device)
    dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):
        a = v[:, None] / dim_t
        return torch.stack((a[:, 0::2].sin(), a[:, 1::2].cos()), dim=2).flatten(1)

    py = enc(ys)[:, None, :].expand(h, w, num_pos_feats)
    px = enc(xs)[None, :, :].expand(h, w, num_pos_feats)
    return torch.cat((py, px), dim=2).reshape(h * w, 2 * num_pos_feats).contiguous()


def sine_tables_masked(mask, num_pos_feats, normalize):
    """PositionEmbeddingSine for a PADDED batch (src/models/position_encoding.py:28-48): mask [B,h,w] bool (True =
    padding) -> per-image fp32 tables [B, h*w, 2*num_pos_feats] (same channel order as sine_table).  A dozen small
    torch ops on [B,h,w] maps; no host synchronisation."""
    not_mask = ~mask
    y = not_mask.cumsum(1, dtype=torch.float32)
    x = not_mask.cumsum(2, dtype=torch.float32)
    if normalize:
        y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
        x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = 10000 ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)

    def enc(v):
        a = v[..., None] / dim_t
        return torch.stack((a[..., 0::2].sin(), a[..., 1::2].cos()), dim=4).flatten(3)

    B, h, w = mask.shape
    return torch.cat((enc(y), enc(x)), dim=3).reshape(B, h * w, 2 * num_pos_feats).contiguous()


def level_mask(mask, size):
    """the padding mask at a backbone level (src/models/backbone.py:79: legacy nearest interpolation)"""
    return F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]


def shift_mask(H, W, ws, shift, device):
    """the SW-MSA mask of multiscale_transformerr.py:937-955 (-100 between different regions), fp32 [nW, N, N]"""
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    region = torch.zeros(Hp, Wp, device=device)
    bounds = ((0, -ws), (-ws, -shift), (-shift, None))
    cnt = 0
    for hs in bounds:
        for wsl in bounds:
            region[hs[0]:hs[1], wsl[0]:wsl[1]] = cnt
            cnt += 1
    win = region.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff)).contiguous()


def rel_pos_bias(table, ws, nheads):
    """relative position bias gather of multiscale_transformerr.py:236-246,313-315 -> fp32 [heads, N, N]"""
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = c[:, :, None] - c[:, None, :]
    idx = ((rel[0] + ws - 1) * (2 * ws - 1) + rel[1] + ws - 1).to(table.device)
    return table[idx.view(-1)].view(ws * ws, ws * ws, nheads).permute(2, 0, 1).float().contiguous()


class _LN:
    def __init__(self, sd, name, n_pad=None):
        w, b = sd[name + ".weight"], sd[name + ".bias"]
        n_pad = n_pad or round_up(w.numel(), 8)
        self.g, self.b, self.n = pad_vec(w, n_pad), pad_vec(b, n_pad), w.numel()

    @property
    def pair(self):
        return (self.g, self.b)


def _compose(w1, b1, w2, b2):
    """Linear(w2,b2) o Linear(w1,b1) with no activation in between == one Linear"""
    return w2.double() @ w1.double(), w2.double() @ b1.double() + b2.double()


class Engine:
    def __init__(self, state_dict, cfg=None, device="cuda"):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.dev = torch.device(device)
        sd = {k: v.detach().to(self.dev, torch.float32) for k, v in state_dict.items() if v.is_floating_point()}
        self.sd = sd
        self._tables = {}
        self._graphs = {}
        self.max_graphs = 8
        self._side, self._side2 = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev)
        self._side3 = torch.cuda.Stream(device=self.dev)
        self._pack_backbone()
        self._pack_detr()
        self._pack_dense()
        self._pack_head()
        del self.sd

    # ------------------------------------------------------------------ setup: backbone (cuDNN, frozen BN folded)
    def _fold(self, conv, bn):
        """FrozenBatchNorm2d folded into the preceding conv (src/models/backbone.py:46-55, eps inside the rsqrt)"""
        sd = self.sd
        scale = sd[bn + ".weight"] * (sd[bn + ".running_var"] + 1e-5).rsqrt()
        shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
        w = (sd[conv + ".weight"] * scale.view(-1, 1, 1, 1)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        return w, shift.to(torch.bfloat16)

    def _fold_packed(self, conv, bn, taps):
        """same folding, packed for gwd_conv_gemm (1x1 -> taps=1, 3x3 stride 1 -> taps=9); the BN shift is the bias"""
        sd = self.sd
        scale = sd[bn + ".weight"] * (sd[bn + ".running_var"] + 1e-5).rsqrt()
        shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
        w = sd[conv + ".weight"] * scale.view(-1, 1, 1, 1)
        return pack_linear(w.flatten(1), shift) if taps == 1 else pack_conv3x3(w, shift)

    def _pack_backbone(self):
        """The stem (7x7/2 conv + BN + ReLU + max-pool) is one gwd_stem_conv_pool launch; every convolution of the bottlenecks
        runs on gwd_conv_gemm with the folded FrozenBN shift, the ReLU and the residual add fused in its epilogue: 1x1 and
        stride-1 3x3 directly, the three stride-2 3x3 as gwd_im2col3x3_s2 + a Linear, the three stride-2 1x1 projections through
        the strided TMA view.  No cuDNN call is left."""
        p = "backbone.0.body."
        sd = self.sd
        scale = sd[p + "bn1.weight"] * (sd[p + "bn1.running_var"] + 1e-5).rsqrt()
        self.stem = ops.pack_stem(sd[p + "conv1.weight"] * scale.view(-1, 1, 1, 1), sd[p + "bn1.bias"] - sd[p + "bn1.running_mean"] * scale)
        self.blocks = []
        for li, nb in enumerate((3, 4, 6, 3), start=1):
            stage = []
            for bi in range(nb):
                q = "%slayer%d.%d." % (p, li, bi)
                stride = 2 if (li > 1 and bi == 0) else 1
                blk = {"c1": self._fold_packed(q + "conv1", q + "bn1", 1), "c3": self._fold_packed(q + "conv3", q + "bn3", 1),
                       "stride": stride}
                if stride == 2:      # stride-2 3x3 = gwd_im2col3x3_s2 + one Linear over K = (ky, kx, c)
                    scale2 = sd[q + "bn2.weight"] * (sd[q + "bn2.running_var"] + 1e-5).rsqrt()
                    w2 = (sd[q + "conv2.weight"] * scale2.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).flatten(1)
                    blk["c2"] = pack_linear(w2, sd[q + "bn2.bias"] - sd[q + "bn2.running_mean"] * scale2)
                else:
                    blk["c2"] = self._fold_packed(q + "conv2", q + "bn2", 9)
                if (q + "downsample.0.weight") in self.sd:     # 1x1 projection; stride 2 = the same Linear on every other pixel
                    blk["down"] = self._fold_packed(q + "downsample.0", q + "downsample.1", 1)
                stage.append(blk)
            self.blocks.append(stage)

    def backbone(self, images):
        """torchvision-style ResNet-50 C2..C5 with frozen batch-norm (src/models/backbone.py:19-92), bf16 channels-last.
        Returns [B,h,w,C] bf16 maps."""
        x = ops.stem_conv_pool(images, *self.stem)                           # -> [B,h,w,64] channels-last buffer
        feats = []
        for stage in self.blocks:
            for blk in stage:
                y = conv_gemm(x, blk["c1"], post_act=ACT_RELU)
                if blk["stride"] == 2:
                    # stride-2 3x3: patches by gwd_im2col3x3_s2, then the tcgen05 Linear with the BN shift + ReLU fused;
                    # stride-2 1x1 projection: every other pixel through the strided TMA view
                    col = ops.im2col3x3_s2(y)
                    Bc, ho, wo, _ = col.shape
                    y = conv_gemm(col.view(Bc * ho * wo, -1), blk["c2"], post_act=ACT_RELU).view(Bc, ho, wo, -1)
                    idt = conv_gemm(x, blk["down"], subsample2=True)
                else:
                    y = conv_gemm(y, blk["c2"], post_act=ACT_RELU)
                    idt = conv_gemm(x, blk["down"]) if "down" in blk else x
                x = conv_gemm(y.contiguous(), blk["c3"], res=idt.contiguous(), res_mode=RES_BEFORE_NORM, post_act=ACT_RELU)
            feats.append(x)
        return feats

    # ------------------------------------------------------------------ setup: DETR transformer
    def _pack_mha(self, prefix, E, nheads):
        sd = self.sd
        w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
        s = (E // nheads) ** -0.5     # q scaling applied after the bias (multi_head_attention.py:236,276) -> fold both
        return {
            "q": pack_linear(w[:E] * s, b[:E] * s), "k": pack_linear(w[E:2 * E], b[E:2 * E]),
            "qk": pack_linear(torch.cat([w[:E] * s, w[E:2 * E]]), torch.cat([b[:E] * s, b[E:2 * E]])),
            "v": pack_linear(w[2 * E:], b[2 * E:]),
            "o": pack_linear(sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"]),
        }

    def _pack_detr(self):
        sd, c = self.sd, self.cfg
        E, nh = c["hidden_dim"], c["nheads"]
        self.input_proj = pack_linear(sd["input_proj.weight"].flatten(1), sd["input_proj.bias"])
        self.dense_input_proj = pack_linear(sd["dense_input_proj.weight"].flatten(1), sd["dense_input_proj.bias"])
        self.enc = []
        for i in range(c["enc_layers"]):
            p = "transformer.encoder.layers.%d." % i
            self.enc.append({"attn": self._pack_mha(p + "self_attn.", E, nh),
                             "l1": pack_linear(sd[p + "linear1.weight"], sd[p + "linear1.bias"]),
                             "l2": pack_linear(sd[p + "linear2.weight"], sd[p + "linear2.bias"]),
                             "n1": _LN(sd, p + "norm1"), "n2": _LN(sd, p + "norm2")})
        self.dec = []
        for i in range(c["dec_layers"]):
            p = "transformer.decoder.layers.%d." % i
            self.dec.append({"self": self._pack_mha(p + "self_attn.", E, nh),
                             "cross": self._pack_mha(p + "multihead_attn.", E, nh),
                             "l1": pack_linear(sd[p + "linear1.weight"], sd[p + "linear1.bias"]),
                             "l2": pack_linear(sd[p + "linear2.weight"], sd[p + "linear2.bias"]),
                             "n1": _LN(sd, p + "norm1"), "n2": _LN(sd, p + "norm2"), "n3": _LN(sd, p + "norm3")})
        self.dec_norm = _LN(sd, "transformer.decoder.norm")
        self.query_pos = sd["query_embed.weight"].to(torch.bfloat16).contiguous()
        self.class_embed = pack_linear(sd["class_embed.weight"], sd["class_embed.bias"])
        self.lines_embed = [pack_linear(sd["lines_embed.layers.%d.weight" % i], sd["lines_embed.layers.%d.bias" % i])
                            for i in range(3)]

    # ------------------------------------------------------------------ setup: dense encoder
    def _pack_mlp(self, p):
        sd = self.sd
        return (pack_linear(sd[p + "fc1.weight"], sd[p + "fc1.bias"]), pack_linear(sd[p + "fc2.weight"], sd[p + "fc2.bias"]))

    def _pack_composed(self, first, second, col_map=None, cin_pad=None):
        sd = self.sd
        w, b = _compose(sd[first + ".weight"], sd[first + ".bias"], sd[second + ".weight"], sd[second + ".bias"])
        return pack_linear(w.float(), b.float(), cin_pad=cin_pad, col_map=col_map)

    def _pack_pyramid(self, p, K):
        """PyramidLayer (points_sample.py:45-125): channel counts K, 2K, 5*2K, 4K padded to multiples of 16"""
        sd = self.sd
        Kp, C2, C2p, C4 = round_up(K, 16), 2 * K, round_up(2 * K, 16), 4 * K

        def cl(name, cin_pad, col_map=None):
            pw = pack_conv3x3(sd[p + name + ".conv.weight"], None, cin_pad=cin_pad, col_map=col_map)
            return pw, _LN(sd, p + name + ".layer_norm", pw.n_pad)

        py = {"K": K, "Kp": Kp, "C2p": C2p, "first0": cl("firstconv.0", Kp), "first2": cl("firstconv.2", Kp), "blocks": []}
        for lname, nblk in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
            for b in range(nblk):
                py["blocks"].append((cl("%s.%d.conv1.0" % (lname, b), C2p), cl("%s.%d.conv2" % (lname, b), C2p)))
        py["branches"] = [cl("branch%d.1" % i, C2p) for i in range(1, 5)]
        cmap = torch.cat([torch.arange(C2) + g * C2p for g in range(5)])
        py["last0"] = cl("lastconv.0", 5 * C2p, cmap)
        py["last2"] = pack_linear(sd[p + "lastconv.2.weight"].flatten(1), None, cin_pad=round_up(C4, 16))
        return py

    def _pack_dense(self):
        sd, c = self.sd, self.cfg
        D, nh, ws, td = c["dense_trans_dim"], c["dense_trans_heads"], c["window"], c["class_token_dim"]
        p = "dense_encoder."
        hd = D // nh
        s = hd ** -0.5
        self.line_blocks = []
        for i in range(c["dense_trans_layers"][0]):
            q = "%sdense_transformer.blocks.%d." % (p, i)
            wqkv, bqkv = sd[q + "attn.qkv.weight"].clone(), sd[q + "attn.qkv.bias"].clone()
            wqkv[:D] *= s
            bqkv[:D] *= s
            # ref_k = mu + exp(logsigma) * Linear(x_ref) (multiscale_transformerr.py:281-288) folded into the weights
            wr, br = sd[q + "attn.ref_qk.weight"].clone(), sd[q + "attn.ref_qk.bias"].clone()
            sig, mu = sd[q + "attn.diff_logsigma"].exp().flatten(), sd[q + "attn.diff_mu"].flatten()
            wr[:D] *= sig[:, None]
            br[:D] = mu + sig * br[:D]
            self.line_blocks.append({
                "n1": _LN(sd, q + "norm1"), "n2": _LN(sd, q + "norm2"), "qkv": pack_linear(wqkv, bqkv),
                "ref": pack_linear(wr, br), "proj": pack_linear(sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"]),
                # the 16x16x3x3 diffusion filter travels as a kernel parameter: keep it on the host
                "diff_w": sd[q + "attn.ref_attn_diffusion.weight"].float().cpu().contiguous(),
                "diff_b": sd[q + "attn.ref_attn_diffusion.bias"].float().cpu().contiguous(),
                "bias": rel_pos_bias(sd[q + "attn.relative_position_bias_table"], ws, nh), "mlp": self._pack_mlp(q + "mlp.")})
        self.depth32 = self._pack_composed(p + "depth_pred32.0", p + "depth_pred32.1")
        self.class_stages = []
        for si, depth in enumerate(c["class_trans_layers"], start=1):
            C = D >> si
            sc = (C // nh) ** -0.5
            blocks = []
            for i in range(depth):
                q = "%sclass_transformer%d.blocks.%d." % (p, si, i)
                wqkv, bqkv = sd[q + "attn.qkv.weight"].clone(), sd[q + "attn.qkv.bias"].clone()
                wqkv[:C] *= sc
                bqkv[:C] *= sc
                blocks.append({
                    "n1": _LN(sd, q + "norm1"), "n2": _LN(sd, q + "norm2"), "nd1": _LN(sd, q + "norm_depth1"),
                    "ns1": _LN(sd, q + "norm_seg1"), "nd2": _LN(sd, q + "norm_depth2"), "ns2": _LN(sd, q + "norm_seg2"),
                    "qkv": pack_linear(wqkv, bqkv), "proj": pack_linear(sd[q + "attn.proj.weight"], sd[q + "attn.proj.bias"]),
                    "dq": pack_linear(sd[q + "attn.cls_dth_q.weight"], sd[q + "attn.cls_dth_q.bias"]),
                    "sq": pack_linear(sd[q + "attn.cls_seg_q.weight"], sd[q + "attn.cls_seg_q.bias"]),
                    "gkv": pack_linear(torch.cat([sd[q + "attn.global_k.weight"], sd[q + "attn.global_v.weight"]]),
                                       torch.cat([sd[q + "attn.global_k.bias"], sd[q + "attn.global_v.bias"]])),
                    # the seg token is projected with proj_dth too (multiscale_transformerr.py:578)
                    "pdth": pack_linear(sd[q + "attn.proj_dth.weight"], sd[q + "attn.proj_dth.bias"]),
                    "bias": rel_pos_bias(sd[q + "attn.relative_position_bias_table"], ws, nh), "scale": sc,
                    "mlp": self._pack_mlp(q + "mlp."), "mlp_d": self._pack_mlp(q + "mlp_depth."),
                    "mlp_s": self._pack_mlp(q + "mlp_seg.")})
            st = {"C": C, "blocks": blocks,
                  "proj_class": pack_linear(sd["%sproj_class%d.weight" % (p, si)], sd["%sproj_class%d.bias" % (p, si)]),
                  "proj_backbn": pack_conv3x3(sd["%sproj_backbn%d.conv.weight" % (p, si)], sd["%sproj_backbn%d.conv.bias" % (p, si)])}
            if si > 1:
                scale_name = {2: "8", 3: "4"}[si]
                for kind in ("depth", "seg"):
                    q = "%sold_%s_token_proj%s" % (p, kind, scale_name)
                    st["tok_" + kind] = (self._pack_composed(q + ".fc1", q + ".fc2"), _LN(sd, q + ".norm"))
            self.class_stages.append(st)
        self.depth_token = sd[p + "depth_token"].to(torch.bfloat16).view(1, td)
        self.seg_token = sd[p + "seg_token"].to(torch.bfloat16).view(1, td)
        C1 = D >> 1
        self.depth16 = self._pack_composed(p + "depth_pred16.0", p + "depth_pred16.1", cin_pad=C1 + td)
        self.pbp = []
        for j, (si, K) in enumerate(zip((2, 3), c["interval_sample_num"]), start=1):
            C = D >> si
            q = "%spoint_based_pred%d." % (p, j)
            self.pbp.append({"dim": C, "K": K, "pre_refer": self._pack_composed(q + "pre_proj", q + "refer_proj", cin_pad=C + td),
                             "pyr": self._pack_pyramid(q + "pyramid.", K)})

    # ------------------------------------------------------------------ setup: dense prediction head
    def _pack_head(self):
        sd, td = self.sd, self.cfg["class_token_dim"]
        p = "depth_decoder."
        C = self.cfg["dense_trans_dim"] >> 3
        width = C + 3 * td    # stage buffer at 1/4: [x | depth token | seg token | depth_pred3 (+7 pad)]
        feat, dtok, stok = torch.arange(C), torch.arange(td) + C, torch.arange(td) + C + td
        d3 = torch.tensor([C + 2 * td])
        hid_d, hid_s = round_up(C + 1 + td, 16), round_up(C + td, 16)
        self.head = {}
        self.head["depth_fc1"] = pack_linear(sd[p + "depth_token_fuse.fc1.weight"], sd[p + "depth_token_fuse.fc1.bias"],
                                             cin_pad=width, col_map=torch.cat([feat, d3, dtok]))
        self.head["depth_fc2"] = pack_linear(sd[p + "depth_token_fuse.fc2.weight"], sd[p + "depth_token_fuse.fc2.bias"], cin_pad=hid_d)
        self.head["seg_fc1"] = pack_linear(sd[p + "seg_token_fuse.fc1.weight"], sd[p + "seg_token_fuse.fc1.bias"],
                                           cin_pad=width, col_map=torch.cat([feat, stok]))
        self.head["seg_fc2"] = pack_linear(sd[p + "seg_token_fuse.fc2.weight"], sd[p + "seg_token_fuse.fc2.bias"], cin_pad=hid_s)
        for kind in ("depth", "seg"):
            self.head["up1_" + kind] = ops.pack_upconv3x3(sd[p + "upconv1_%s.conv.weight" % kind])
            self.head["norm_" + kind] = _LN(sd, p + "norm_" + kind)
            self.head["conv1_" + kind] = pack_conv3x3(sd[p + "conv1_%s.0.weight" % kind])
            self.head["up2_" + kind] = ops.pack_upconv3x3(sd[p + "upconv2_%s.conv.weight" % kind])
            self.head["conv2_" + kind] = pack_conv3x3(sd[p + "conv2_%s.0.weight" % kind])
        self.head["get_depth"] = pack_conv3x3(sd[p + "get_depth.0.weight"])
        self.head["get_seg"] = pack_conv3x3(sd[p + "get_seg.weight"])
        self.width4 = width

    # ------------------------------------------------------------------ cached, input-independent tables
    def table(self, key, fn):
        if key not in self._tables:
            self._tables[key] = fn()
        return self._tables[key]

    # ------------------------------------------------------------------ DETR transformer
    def _mha(self, pk, q_in, k_in, v_in, B, Lq, Lk, E, nh, fused_qk, key_padding=None, kv=None):
        """multi_head_attention_forward (src/models/multi_head_attention.py:188-380) without the dead head-averaged
        weights.  q_in/k_in/v_in are [B*L, E] token matrices; returns the un-projected attention output."""
        hd = E // nh
        if kv is not None:       # cross attention: K / V of the memory were projected ahead of the decoder chain
            q, k, v, q_rs, k_rs = conv_gemm(q_in, pk["q"]), kv[0], kv[1], E, E
        else:
            # the V projection does not depend on the Q|K projection: a parallel branch (side stream / graph fork)
            main = torch.cuda.current_stream()
            v = torch.empty(v_in.shape[0], E, dtype=torch.bfloat16, device=self.dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                conv_gemm(v_in, pk["v"], out=v)
            qk = conv_gemm(q_in, pk["qk"])
            q, k, q_rs, k_rs = qk, qk[:, E:], 2 * E, 2 * E
            main.wait_stream(self._side)
        o = torch.empty(B * Lq, E, dtype=torch.bfloat16, device=self.dev)
        ops.attention(q, k, v, o, items=B, heads=nh, Lq=Lq, Lk=Lk, hd=hd, q_strides=(Lq * q_rs, q_rs),
                      k_strides=(Lk * k_rs, k_rs), v_strides=(Lk * E, E), o_strides=(Lq * E, E), key_padding=key_padding)
        return o

    def detr(self, src, B, h, w, mask5=None):
        """Transformer.forward (src/models/transformer.py:47-61): 6 post-norm encoder layers (:149-162), 6 decoder layers
        (:212-233) with decoder.norm on every layer output (:105-123).  src: [B*L, E] bf16 tokens."""
        c = self.cfg
        E, nh, L, Q = c["hidden_dim"], c["nheads"], h * w, c["num_queries"]
        if mask5 is None:
            pos, period, kpm = self.table(("pos5", h, w), lambda: sine_table(h, w, E // 2, True, self.dev).to(torch.bfloat16)), L, None
        else:       # padded batch: per-image position codes, padded keys masked out (transformer.py:52-57, mha.py:343-349)
            pos, period = sine_tables_masked(mask5, E // 2, True).view(B * L, E).to(torch.bfloat16), B * L
            kpm = mask5.reshape(B, L).to(torch.uint8).contiguous()
        x = src
        for ly in self.enc:
            xp = ops.add_rows(x, pos, period)
            o = self._mha(ly["attn"], xp, xp, x, B, L, L, E, nh, True, key_padding=kpm)
            x = conv_gemm(o, ly["attn"]["o"], res=x, res_mode=RES_BEFORE_NORM, ln=ly["n1"].pair)
            hmid = conv_gemm(x, ly["l1"], post_act=ACT_RELU)
            x = conv_gemm(hmid, ly["l2"], res=x, res_mode=RES_BEFORE_NORM, ln=ly["n2"].pair)
        memory = x
        mem_pos = ops.add_rows(memory, pos, period)
        tgt = torch.zeros(B * Q, E, dtype=torch.bfloat16, device=self.dev)
        hs = torch.empty(len(self.dec), B * Q, E, dtype=torch.bfloat16, device=self.dev)
        # cross-attention K / V projections of the memory for all decoder layers: independent of the decoder chain, so they
        # run as a parallel branch while the first self-attention block is in flight
        main = torch.cuda.current_stream()
        ckv = [(torch.empty(B * L, E, dtype=torch.bfloat16, device=self.dev), torch.empty(B * L, E, dtype=torch.bfloat16, device=self.dev))
               for _ in self.dec]
        self._side2.wait_stream(main)
        with torch.cuda.stream(self._side2):
            for ly, (ck, cv) in zip(self.dec, ckv):
                conv_gemm(mem_pos, ly["cross"]["k"], out=ck)
                conv_gemm(memory, ly["cross"]["v"], out=cv)
        for i, ly in enumerate(self.dec):
            tq = ops.add_rows(tgt, self.query_pos, Q)
            o = self._mha(ly["self"], tq, tq, tgt, B, Q, Q, E, nh, True)
            tgt = conv_gemm(o, ly["self"]["o"], res=tgt, res_mode=RES_BEFORE_NORM, ln=ly["n1"].pair)
            tq = ops.add_rows(tgt, self.query_pos, Q)
            if i == 0:
                main.wait_stream(self._side2)
            o = self._mha(ly["cross"], tq, mem_pos, memory, B, Q, L, E, nh, False, key_padding=kpm, kv=ckv[i])
            tgt = conv_gemm(o, ly["cross"]["o"], res=tgt, res_mode=RES_BEFORE_NORM, ln=ly["n2"].pair)
            hmid = conv_gemm(tgt, ly["l1"], post_act=ACT_RELU)
            tgt = conv_gemm(hmid, ly["l2"], res=tgt, res_mode=RES_BEFORE_NORM, ln=ly["n3"].pair)
            ops.layernorm(tgt, self.dec_norm.g, self.dec_norm.b, out=hs[i])
        return hs, memory

    def line_heads(self, hs, B):
        """class_embed / lines_embed on all decoder outputs (src/models/glassrgbd.py:89-90); fp32 outputs"""
        nl, Q = hs.shape[0], self.cfg["num_queries"]
        flat = hs.view(nl * B * Q, -1)
        logits = conv_gemm(flat, self.class_embed, out_f32=True)
        t = conv_gemm(flat, self.lines_embed[0], post_act=ACT_RELU)
        t = conv_gemm(t, self.lines_embed[1], post_act=ACT_RELU)
        lines = conv_gemm(t, self.lines_embed[2], post_act=ACT_SIGMOID, out_f32=True)
        return logits.view(nl, B, Q, -1), lines.view(nl, B, Q, -1)

    # ------------------------------------------------------------------ dense encoder pieces
    def _mlp_res(self, x_ln, x_res, mlp, out=None, y_coff=0):
        """x_res + fc2(gelu(fc1(x_ln)))  (multiscale_transformerr.py:55-73,755)"""
        hmid = conv_gemm(x_ln, mlp[0], post_act=ACT_GELU)
        return conv_gemm(hmid, mlp[1], res=x_res, res_mode=RES_AFTER, out=out, y_coff=y_coff)

    def line_stage(self, x, B, H, W, ref_xy, mask=None):
        """BasicLayer of WindowAttention blocks at 1/32 (multiscale_transformerr.py:267-332,646-755,926-979).
        x: [B*H*W, D] bf16; ref_xy: fp32 [B, R, 2] line end points in [-1,1]."""
        c = self.cfg
        D, nh, ws = c["dense_trans_dim"], c["dense_trans_heads"], c["window"]
        hd, N, R = D // nh, ws * ws, ref_xy.shape[1]
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        nW = (Hp // ws) * (Wp // ws)
        P = nW * N
        pos = (self.table(("pos32", H, W), lambda: sine_table(H, W, D // 2, False, self.dev)) if mask is None
               else sine_tables_masked(mask, D // 2, False))
        mask = self.table(("mask", H, W), lambda: shift_mask(H, W, ws, ws // 2, self.dev))
        a0 = torch.empty(B, nh, P, R, dtype=torch.float32, device=self.dev)
        a1 = torch.empty_like(a0)
        raw_ws = torch.empty_like(a0)
        stats_ws = torch.empty(B * nh * 2, dtype=torch.float64, device=self.dev)
        for i, blk in enumerate(self.line_blocks):
            shift = 0 if i % 2 == 0 else ws // 2
            xw = ops.window_gather(x, B, H, W, ws, shift, blk["n1"].g, blk["n1"].b)
            xref = ops.line_ref_gather(xw, pos, ref_xy, R, B, H, W, ws, shift, D)
            qkv = conv_gemm(xw, blk["qkv"])
            refkv = conv_gemm(xref.view(B * R, D), blk["ref"], out_f32=True)
            ops.ref_scores(qkv, 3 * D, refkv, 2 * D, a0, B, nW, N, nh, hd, R)
            ops.ref_diffuse(a0, a1, blk["diff_w"], blk["diff_b"], raw_ws, stats_ws, B, nh, P, R)
            ops.ref_diffuse(a1, a0, blk["diff_w"], blk["diff_b"], raw_ws, stats_ws, B, nh, P, R)
            ops.ref_diffuse(a0, a1, blk["diff_w"], blk["diff_b"], raw_ws, stats_ws, B, nh, P, R)
            qnew = torch.empty(B * P, D, dtype=torch.bfloat16, device=self.dev)
            ops.ref_requery(a1, refkv[:, D:], 2 * D, qnew, D, B, nW, N, nh, hd, R, hd ** -0.5)
            o = torch.empty(B * P, D, dtype=torch.bfloat16, device=self.dev)
            ops.attention(qnew, qkv[:, D:], qkv[:, 2 * D:], o, items=B * nW, heads=nh, Lq=N, Lk=N, hd=hd,
                          q_strides=(N * D, D), k_strides=(N * 3 * D, 3 * D), v_strides=(N * 3 * D, 3 * D),
                          o_strides=(N * D, D), bias=blk["bias"], mask=mask if shift else None)
            pr = conv_gemm(o, blk["proj"])
            x, x_ln = ops.window_merge(pr, x, B, H, W, ws, shift, blk["n2"].g, blk["n2"].b, want_ln=True)
            x = self._mlp_res(x_ln, x, blk["mlp"])
        return x

    def class_stage(self, st, buf, B, H, W):
        """BasicLayer of WindowClassAttention blocks (multiscale_transformerr.py:455-580,646-788).  `buf` is the stage
        buffer [B*H*W, C+3*td]: channels [0,C) features, [C,C+td) depth token, [C+td,C+2td) seg token; it is
        updated in place by every block (the last td channels are scratch for the coarse depth)."""
        c = self.cfg
        C, nh, ws, td = st["C"], c["dense_trans_heads"], c["window"], c["class_token_dim"]
        hd, N = C // nh, ws * ws
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        nW = (Hp // ws) * (Wp // ws)
        rows_w = B * nW * N
        tC = C + 2 * td
        mask = self.table(("mask", H, W), lambda: shift_mask(H, W, ws, ws // 2, self.dev))
        for i, blk in enumerate(st["blocks"]):
            shift = 0 if i % 2 == 0 else ws // 2
            xw = ops.window_gather(buf, B, H, W, ws, shift, blk["n1"].g, blk["n1"].b, C=C, x_coff=0)
            tx = torch.empty(rows_w, tC, dtype=torch.bfloat16, device=self.dev)   # cat[x_attn, depth tok, seg tok]
            ops.window_gather(buf, B, H, W, ws, shift, blk["nd1"].g, blk["nd1"].b, C=td, x_coff=C, out=tx, y_coff=C)
            ops.window_gather(buf, B, H, W, ws, shift, blk["ns1"].g, blk["ns1"].b, C=td, x_coff=C + td, out=tx, y_coff=C + td)
            qkv = conv_gemm(xw, blk["qkv"])
            o = torch.empty(rows_w, C, dtype=torch.bfloat16, device=self.dev)
            ops.attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o, items=B * nW, heads=nh, Lq=N, Lk=N, hd=hd,
                          q_strides=(N * 3 * C, 3 * C), k_strides=(N * 3 * C, 3 * C), v_strides=(N * 3 * C, 3 * C),
                          o_strides=(N * C, C), bias=blk["bias"], mask=mask if shift else None)
            conv_gemm(o, blk["proj"], out=tx, y_coff=0)
            dq = conv_gemm(tx, blk["dq"], x_coff=C)
            sq = conv_gemm(tx, blk["sq"], x_coff=C + td)
            gkv = conv_gemm(tx, blk["gkv"])                                      # [rows, 2*tC] = global_k | global_v
            dout = torch.empty(rows_w, td, dtype=torch.bfloat16, device=self.dev)
            sout = torch.empty_like(dout)
            ops.token_attention(dq, sq, gkv, gkv[:, tC:], dout, sout, items=B * nW, N=N, heads=nh, td=td // nh,
                                tc=tC // nh, q_rs=td, k_rs=2 * tC, v_rs=2 * tC, o_rs=td, scale=blk["scale"])
            dpr = conv_gemm(dout, blk["pdth"])
            spr = conv_gemm(sout, blk["pdth"])
            x_new, x_ln = ops.window_merge(tx, buf, B, H, W, ws, shift, blk["n2"].g, blk["n2"].b, want_ln=True, C=C, sc_coff=0)
            d_new, d_ln = ops.window_merge(dpr, buf, B, H, W, ws, shift, blk["nd2"].g, blk["nd2"].b, want_ln=True, C=td, sc_coff=C)
            s_new, s_ln = ops.window_merge(spr, buf, B, H, W, ws, shift, blk["ns2"].g, blk["ns2"].b, want_ln=True, C=td, sc_coff=C + td)
            self._mlp_res(x_ln, x_new, blk["mlp"], out=buf, y_coff=0)
            self._mlp_res(d_ln, d_new, blk["mlp_d"], out=buf, y_coff=C)
            self._mlp_res(s_ln, s_new, blk["mlp_s"], out=buf, y_coff=C + td)
        return buf

    def conv_ln(self, x, pk, act=ACT_NONE, res=None, out=None, y_coff=0):
        """ConvLn (+GELU / +residual): 3x3 conv, LayerNorm over channels fused in the epilogue (points_sample.py:12-43)"""
        pw, ln = pk
        return conv_gemm(x, pw, bias=False, ln=ln.pair, post_act=act, res=res, res_mode=RES_AFTER if res is not None else 0,
                         out=out, y_coff=y_coff)

    def pyramid(self, py, rg, B, H, W):
        """PyramidLayer.forward (points_sample.py:106-125) on rg [B,H,W,Kp] -> mixture logits [B,H,W,Kp]"""
        if H < 16 or W < 16:
            raise NotImplementedError("pad_before_pool path (feature map smaller than the 16-pixel pool) is not built")
        C2p = py["C2p"]
        x = self.conv_ln(rg, py["first0"], ACT_GELU)
        x = self.conv_ln(x, py["first2"], ACT_GELU)
        cat = torch.empty(B, H, W, 5 * C2p, dtype=torch.bfloat16, device=self.dev)
        nb = len(py["blocks"])
        for i, (c1, c2) in enumerate(py["blocks"]):
            y = self.conv_ln(x, c1, ACT_GELU)
            x = self.conv_ln(y, c2, res=x, out=cat if i == nb - 1 else None)     # last block lands in slice 0 of the concat
        # the four pooled maps in one pass over slice 0 of the concat, the four up-sampled branches in one pass over slices 1-4
        pooled = ops.avgpool_pyramid(cat, C=C2p)                      # pools (16, 8, 4, 2)
        ups = [self.conv_ln(p_, br, ACT_GELU) for p_, br in zip(pooled, py["branches"])]
        ops.bilinear_up4_into(ups, cat, C2p, H, W)
        pw, ln = py["last0"]
        if pw.n_pad <= 256:
            y = conv_gemm(cat, pw, bias=False, ln=ln.pair, post_act=ACT_GELU)
        else:   # LayerNorm over more channels than one accumulator tile: separate pass
            y = ops.layernorm(conv_gemm(cat, pw, bias=False), ln.g, ln.b, act=ACT_GELU, n=ln.n)
        return conv_gemm(y, py["last2"], bias=False, out_channels=py["Kp"])

    def point_based_pred(self, pb, buf, pre_depth, coords, B, H, W, pos):
        """PointBasedPred.forward (points_sample.py:257-280); buf: stage buffer [B*H*W, C+3td] (features | depth token)"""
        dim, K = pb["dim"], pb["K"]
        Kp = pb["pyr"]["Kp"]
        pr = conv_gemm(buf, pb["pre_refer"])                                   # [rows, 2*dim] = xg | xr
        refer = ops.sample_bilinear(pr, dim, pos, B, H, W, dim, coords, K)      # fp32 [B,K,dim]
        anchor = ops.sample_scalar(pre_depth, coords, K)                        # fp32 [B,K]
        wimg = torch.zeros(B, Kp, dim, dtype=torch.bfloat16, device=self.dev)
        wimg[:, :K] = refer * (dim ** -2)                                       # points_sample.py:273
        rg = conv_gemm(pr.view(B, 1, H * W, 2 * dim), ops.PackedWeight(wimg.view(B, Kp, dim), None, 1, K, dim),
                       bias=False, w_per_image=True, out_channels=Kp)
        logits = self.pyramid(pb["pyr"], rg.view(B, H, W, Kp), B, H, W)
        return ops.anchor_mix(logits, anchor, B, H * W, K).view(B, H, W)

    def dense_encoder(self, dense_in, feats, pred_lines, pred_logits, B, h5, w5, pinned, trace, masks=None, cbs=None):
        """ReferTransformer.forward (multiscale_transformerr.py:1151-1319)"""
        c = self.cfg
        D, td, R0 = c["dense_trans_dim"], c["class_token_dim"], c["num_ref"]
        # top-num_ref lines by raw line logit -> end points in [-1,1]  (:1165-1179)
        if "line_ids" in pinned:     # parity tests pin the selection to the oracle's
            ids = pinned["line_ids"]
            chosen = torch.gather(pred_lines, 1, ids[:, :, None].expand(-1, -1, pred_lines.shape[-1]))
            pts = chosen.reshape(B, R0, -1, 2) * 2 - 1.0
            if not c["with_dense_center"]:
                pts = pts[:, :, :2]
            ref_xy = pts.reshape(B, -1, 2).contiguous().float()
        else:
            ref_xy, ids = ops.select_lines(pred_logits.contiguous(), pred_lines.contiguous(), R0, 3 if c["with_dense_center"] else 2)
        x32 = self.line_stage(dense_in, B, h5, w5, ref_xy, mask=None if masks is None else masks[3])
        depth0 = conv_gemm(x32, self.depth32, post_act=ACT_SIGMOID, out_f32=True).view(B, h5, w5)
        edges = [c["min_depth_eval"] / c["max_depth_eval"]] + list(c["depth_interval"]) + [1.0]
        depths, prev, (ph, pw_) = [], x32.view(B, h5, w5, D), (h5, w5)
        prev_c = D
        bufs = []
        coords = None
        for si, st in enumerate(self.class_stages):
            f = feats[2 - si]
            H, W = f.shape[1:3]
            C = st["C"]
            buf = torch.empty(B * H * W, C + 3 * td, dtype=torch.bfloat16, device=self.dev)
            buf[:, C + 2 * td:] = 0
            pc = conv_gemm(prev, st["proj_class"])       # proj_class commutes with the nearest up-sampling: run it at low res
            cb = cbs[si] if cbs is not None else conv_gemm(f, st["proj_backbn"], post_act=ACT_GELU)
            ops.upsample_nearest(pc.view(B, ph, pw_, C), H, W, add=cb, out=buf.view(B, H, W, -1), y_coff=0)
            if si == 0:
                buf[:, C:C + td] = self.depth_token
                buf[:, C + td:C + 2 * td] = self.seg_token
            else:
                pbuf = bufs[-1]
                for kind, off in (("depth", 0), ("seg", td)):
                    pk, ln = st["tok_" + kind]
                    t = conv_gemm(pbuf, pk, x_coff=prev_c + off, ln=ln.pair)
                    ops.upsample_nearest(t.view(B, ph, pw_, td), H, W, out=buf.view(B, H, W, -1), y_coff=C + off)
            self.class_stage(st, buf, B, H, W)
            if si == 0:
                d = conv_gemm(buf, self.depth16, post_act=ACT_SIGMOID, out_f32=True).view(B, H, W)
            else:
                pb = self.pbp[si - 1]
                pos = (self.table(("pos", H, W, C), lambda: sine_table(H, W, C // 2, False, self.dev)) if masks is None
                       else sine_tables_masked(masks[2 - si], C // 2, False))
                d = self.point_based_pred(pb, buf, depths[-1], coords, B, H, W, pos)
            depths.append(d)
            if si < 2:
                key = "sample%d" % (si + 1)
                if key in pinned:
                    coords = pinned[key].to(self.dev).float().contiguous()
                else:
                    small = depth0 if si == 0 else depths[0]
                    coords, idx = ops.certain_sample(small, d, c["interval_sample_num"][si], edges)
                    if trace is not None:
                        trace[key + "_idx"] = idx
                if trace is not None:
                    trace[key] = coords
            bufs.append(buf)
            prev, (ph, pw_), prev_c = buf, (H, W), C
        if trace is not None:
            trace.update(line_ids=ids, x32=x32, depth0=depth0, bufs=bufs)
        return bufs[-1], depths

    # ------------------------------------------------------------------ dense prediction head
    def dense_head(self, buf4, depth3, B, H4, W4, H, W):
        """DensePrediction.forward (src/models/dense_upsample.py:160-182).  buf4 = [x3 | depth tok | seg tok | scratch]"""
        c = self.cfg
        C, td, hd_ = c["dense_trans_dim"] >> 3, c["class_token_dim"], self.head
        buf4[:, C + 2 * td:] = 0
        buf4[:, C + 2 * td] = depth3.reshape(-1).to(torch.bfloat16)
        outs = {}
        for kind in ("depth", "seg"):
            t = conv_gemm(buf4, hd_[kind + "_fc1"], post_act=ACT_GELU, out_channels=hd_[kind + "_fc2"].cin_pad)
            f = conv_gemm(t, hd_[kind + "_fc2"]).view(B, H4, W4, td)
            if (2 * H4, 2 * W4) != (H // 2, W // 2) or (H % 4 or W % 4):
                raise NotImplementedError("the fused up-sampling convolutions need input sizes that are multiples of 4")
            # upconv = nearest x2 + 3x3 conv + ELU (dense_upsample.py:82-90), the up-sampling folded into 4 phase filters
            u = conv_gemm(f, hd_["up1_" + kind], bias=False, pre_act=ACT_ELU, ln=hd_["norm_" + kind].pair)
            u = conv_gemm(u, hd_["conv1_" + kind], bias=False, post_act=ACT_ELU)
            u = conv_gemm(u, hd_["up2_" + kind], bias=False, post_act=ACT_ELU)
            u = conv_gemm(u, hd_["conv2_" + kind], bias=False, post_act=ACT_ELU)
            outs[kind] = u
        depth = conv_gemm(outs["depth"], hd_["get_depth"], bias=False, post_act=ACT_SIGMOID, out_scale=float(c["max_depth"]),
                          out_f32=True).view(B, 1, H, W)
        seg = conv_gemm(outs["seg"], hd_["get_seg"], bias=False, out_f32=True).permute(0, 3, 1, 2)
        return depth, seg

    # ------------------------------------------------------------------ CUDA-graph replay of the whole forward
    @torch.no_grad()
    def forward_graphed(self, images, mask=None, slot=0):
        """The forward has no host synchronisation and no data-dependent control flow, so the ~480 launches of a step
        are captured once per input shape and replayed as one CUDA graph (the reference cannot: 36 `torch.equal` syncs,
        `int()` reads in CertainSample, ...).  A padded batch is captured with its padding mask as a second STATIC input (the
        per-image position codes and key-padding masks are computed from it inside the graph).  Returns the graph's static
        output tensors: consume (or clone) them before the next call with the same shape.  At most `max_graphs` shapes are
        kept (least recently used first out): evaluation over many image sizes would otherwise hold one memory pool per shape."""
        # slot: independent graph instances (own static buffers and memory pool) of the same shape, so that a serving loop can
        # replay batch i+1 on another stream while batch i is still in flight (the latency-bound phases of one batch -- DETR
        # chain, coarse Swin stages -- leave most SMs idle; two interleaved replays fill them)
        key = (tuple(images.shape), mask is not None, slot)
        entry = self._graphs.pop(key, None)
        if entry is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            static_in = torch.empty_like(images)
            static_in.copy_(images)
            static_mask = mask.clone() if mask is not None else None
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up: lazy kernel attributes, cached tables
                for _ in range(2):
                    self.forward(static_in, mask=static_mask)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.forward(static_in, mask=static_mask)
            entry = (graph, static_in, static_out, static_mask)
        self._graphs[key] = entry                  # (re-)inserted last = most recently used
        graph, static_in, static_out, static_mask = entry
        static_in.copy_(images, non_blocking=True)
        if static_mask is not None:
            static_mask.copy_(mask, non_blocking=True)
        graph.replay()
        return static_out

    # ------------------------------------------------------------------ full forward
    @torch.no_grad()
    def forward(self, images, pinned=None, trace=None, mask=None):
        """GlassRGBD.forward (src/models/glassrgbd.py:74-123).  mask: None for an equal-size batch, else the bool
        [B,H,W] padding mask of nested_tensor_from_tensor_list (True = padding): the convolutions run on the zero-padded
        batch exactly as the reference's do; the mask enters through the per-image position codes (all four levels) and
        the key-padding mask of the encoder self-attention and decoder cross-attention."""
        c = self.cfg
        pinned = pinned or {}
        B, _, H, W = images.shape
        feats = self.backbone(images)
        masks = None if mask is None else [level_mask(mask, f.shape[1:3]) for f in feats]
        c5 = feats[3]
        h5, w5 = c5.shape[1:3]
        tok5 = c5.reshape(B * h5 * w5, c5.shape[-1])
        # the dense branch's projections of the backbone maps (dense_input_proj, proj_backbn1-3: 0.4 ms of wide GEMMs) do not depend
        # on the line branch: a parallel branch next to the latency-bound DETR chain (side stream / graph fork)
        main = torch.cuda.current_stream()
        self._side3.wait_stream(main)
        with torch.cuda.stream(self._side3):
            dense_in = conv_gemm(tok5, self.dense_input_proj)
            cbs = [conv_gemm(feats[2 - si], st["proj_backbn"], post_act=ACT_GELU) for si, st in enumerate(self.class_stages)]
        src = conv_gemm(tok5, self.input_proj)
        hs, memory = self.detr(src, B, h5, w5, mask5=None if masks is None else masks[3])
        logits, lines = self.line_heads(hs, B)
        out = {"pred_logits": logits[-1], "pred_lines": lines[-1]}
        if c["aux_loss"]:
            out["aux_outputs"] = [{"pred_logits": a, "pred_lines": b} for a, b in zip(logits[:-1], lines[:-1])]
        main.wait_stream(self._side3)
        buf4, depths = self.dense_encoder(dense_in, feats, out["pred_lines"], out["pred_logits"], B, h5, w5, pinned, trace, masks, cbs)
        H4, W4 = feats[0].shape[1:3]
        depth, seg = self.dense_head(buf4, depths[-1], B, H4, W4, H, W)
        out["pred_depth"] = [d.unsqueeze(1) for d in depths] + [depth]
        out["pred_seg"] = seg
        if trace is not None:
            trace.update(c5=c5, memory=memory, hs=hs, dense_in=dense_in, src=src)
        return out
