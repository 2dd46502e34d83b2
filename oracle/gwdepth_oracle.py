"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the GW-Depth forward / criterion hot path.

This is the parity oracle of the repo: a functional, plain-PyTorch (CPU, float32) restatement of
what the reference computes on the path named in BASELINE.json.  It is pinned
against the UNMODIFIED reference by tests/test_oracle_vs_reference.py (run in the build container,
where /root/reference exists) and against the committed fixtures in tests/golden/ (which were
produced by the reference itself through oracle/make_golden.py).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it; the product package never does.

Every function cites the reference lines it follows (paths relative to the reference root).
Input weights are a flat {state_dict key: tensor} mapping with the reference's own key names.
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CFG = dict(
    hidden_dim=256, nheads=8, enc_layers=6, dec_layers=6, num_queries=100,
    dense_trans_dim=512, dense_trans_heads=16, dense_trans_layers=(4,), class_trans_layers=(2, 2, 1),
    class_token_dim=64, num_ref=20, with_dense_center=False, window=7,
    interval_sample_num=(30, 80), depth_interval=(0.1, 0.3, 0.5, 0.7, 0.9),
    min_depth_eval=1e-3, max_depth_eval=10.0, max_depth=10.0, aux_loss=True,
    set_cost_class=1.0, set_cost_line=5.0, eos_coef=0.1, variance_focus=0.85, log_depth_error=False,
)


class P:
    """prefix view on a flat state dict"""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def sub(self, name):
        return P(self.sd, self.prefix + name + ".")

    def has(self, k):
        return (self.prefix + k) in self.sd


def linear(x, p, name):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"] if p.has(name + ".bias") else None)


def layer_norm(x, p, name):
    w = p[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, p[name + ".bias"], 1e-5)


# ----------------------------------------------------------------------------------------------
# backbone: torchvision resnet50 (v1.5) with frozen batch-norm     src/models/backbone.py:19-110
# ----------------------------------------------------------------------------------------------
def frozen_bn(x, p, name):  # backbone.py:46-55 (eps added before rsqrt)
    scale = p[name + ".weight"] * (p[name + ".running_var"] + 1e-5).rsqrt()
    shift = p[name + ".bias"] - p[name + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def bottleneck(x, p, stride):
    out = F.relu(frozen_bn(F.conv2d(x, p["conv1.weight"]), p, "bn1"))
    out = F.relu(frozen_bn(F.conv2d(out, p["conv2.weight"], stride=stride, padding=1), p, "bn2"))
    out = frozen_bn(F.conv2d(out, p["conv3.weight"]), p, "bn3")
    if p.has("downsample.0.weight"):
        x = frozen_bn(F.conv2d(x, p["downsample.0.weight"], stride=stride), p, "downsample.1")
    return F.relu(out + x)


def resnet50_features(img, p):
    """-> [C2, C3, C4, C5]  (IntermediateLayerGetter over layer1..layer4, backbone.py:66-69)"""
    x = F.relu(frozen_bn(F.conv2d(img, p["conv1.weight"], stride=2, padding=3), p, "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, nblocks in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblocks):
            x = bottleneck(x, p.sub("layer%d.%d" % (li, bi)), 2 if (li > 1 and bi == 0) else 1)
        feats.append(x)
    return feats


def downsample_mask(mask, size):  # backbone.py:79  (legacy nearest)
    return F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]


def sine_position(mask, num_pos_feats, normalize):
    """position_encoding.py:28-48; mask [B,h,w] bool (True = padding) -> [B, 2*num_pos_feats, h, w]"""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    if normalize:
        y_embed = y_embed / (y_embed[:, -1:, :] + 1e-6) * (2 * math.pi)
        x_embed = x_embed / (x_embed[:, :, -1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=mask.device)
    dim_t = 10000 ** (2 * (dim_t // 2) / num_pos_feats)
    px = x_embed[:, :, :, None] / dim_t
    py = y_embed[:, :, :, None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# DETR transformer            src/models/transformer.py:47-233, multi_head_attention.py:188-380
# ----------------------------------------------------------------------------------------------
def mha(q_in, k_in, v_in, p, nheads, key_padding_mask=None):
    """q_in [Lq,B,E], k_in/v_in [Lk,B,E].  Packed in-proj, bias before q scaling (mha.py:236-276)."""
    Lq, B, E = q_in.shape
    Lk = k_in.shape[0]
    hd = E // nheads
    w, b = p["in_proj_weight"], p["in_proj_bias"]
    q = F.linear(q_in, w[:E], b[:E]) * (hd ** -0.5)
    k = F.linear(k_in, w[E:2 * E], b[E:2 * E])
    v = F.linear(v_in, w[2 * E:], b[2 * E:])
    q = q.contiguous().view(Lq, B * nheads, hd).transpose(0, 1)
    k = k.contiguous().view(Lk, B * nheads, hd).transpose(0, 1)
    v = v.contiguous().view(Lk, B * nheads, hd).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    if key_padding_mask is not None:
        s = s.view(B, nheads, Lq, Lk).masked_fill(key_padding_mask[:, None, None, :], float("-inf")).view(B * nheads, Lq, Lk)
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).contiguous().view(Lq, B, E)
    return F.linear(o, p["out_proj.weight"], p["out_proj.bias"])


def encoder_layer(src, pos, mask, p, nheads):  # transformer.py:149-162 (post-norm, eval: dropout = id)
    q = src + pos
    src = layer_norm(src + mha(q, q, src, p.sub("self_attn"), nheads, mask), p, "norm1")
    ff = linear(F.relu(linear(src, p, "linear1")), p, "linear2")
    return layer_norm(src + ff, p, "norm2")


def decoder_layer(tgt, memory, pos, query_pos, mask, p, nheads):  # transformer.py:212-233
    q = tgt + query_pos
    tgt = layer_norm(tgt + mha(q, q, tgt, p.sub("self_attn"), nheads), p, "norm1")
    tgt = layer_norm(tgt + mha(tgt + query_pos, memory + pos, memory, p.sub("multihead_attn"), nheads, mask), p, "norm2")
    ff = linear(F.relu(linear(tgt, p, "linear1")), p, "linear2")
    return layer_norm(tgt + ff, p, "norm3")


def detr_transformer(src, mask, query_embed, pos, p, cfg):
    """transformer.py:47-61,96-125 -> hs [L_dec, B, Q, E], memory [HW, B, E]"""
    B = src.shape[0]
    x = src.flatten(2).permute(2, 0, 1)
    pos = pos.flatten(2).permute(2, 0, 1)
    qpos = query_embed.unsqueeze(1).repeat(1, B, 1)
    m = mask.flatten(1)
    for i in range(cfg["enc_layers"]):
        x = encoder_layer(x, pos, m, p.sub("encoder.layers.%d" % i), cfg["nheads"])
    memory = x
    tgt = torch.zeros_like(qpos)
    inter = []
    for i in range(cfg["dec_layers"]):
        tgt = decoder_layer(tgt, memory, pos, qpos, m, p.sub("decoder.layers.%d" % i), cfg["nheads"])
        inter.append(layer_norm(tgt, p, "decoder.norm"))
    return torch.stack(inter).transpose(1, 2), memory


# ----------------------------------------------------------------------------------------------
# window machinery                                   src/models/multiscale_transformerr.py:120-168
# ----------------------------------------------------------------------------------------------
def to_windows(x, ws):  # [B,H,W,C] -> [B*nW, ws*ws, C]
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def from_windows(win, ws, B, H, W):  # inverse of to_windows
    C = win.shape[-1]
    x = win.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H, W, C)


def shift_window_mask(Hp, Wp, ws, shift, device=None):  # multiscale_transformerr.py:937-955 (fill value -100)
    img = torch.zeros(1, Hp, Wp, 1, device=device)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = to_windows(img, ws).squeeze(-1)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(am != 0, torch.full_like(am, -100.0), torch.zeros_like(am))


def relative_position_bias(p, ws, nheads):  # multiscale_transformerr.py:313-315
    idx = p["relative_position_index"].view(-1)
    return p["relative_position_bias_table"][idx].view(ws * ws, ws * ws, nheads).permute(2, 0, 1)


def pad_and_shift(x, H, W, ws, shift):
    """[B,H*W,C] -> zero-pad bottom/right to multiples of ws, cyclic shift (mst.py:668-676)"""
    B, _, C = x.shape
    x = x.view(B, H, W, C)
    pr, pb = (ws - W % ws) % ws, (ws - H % ws) % ws
    x = F.pad(x, (0, 0, 0, pr, 0, pb))
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    return x


def unshift_and_crop(x, H, W, shift):  # mst.py:737-745
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    return x[:, :H, :W, :].contiguous()


def mlp(x, p, name):  # mst.py:55-73 (GELU exact)
    return linear(F.gelu(linear(x, p, name + ".fc1")), p, name + ".fc2")


def softmax_attention(q, k, v, bias, mask, nW):
    """q,k,v [B_,h,N,d]; bias [h,N,N]; mask [nW,N,N] or None"""
    a = q @ k.transpose(-2, -1) + bias.unsqueeze(0)
    if mask is not None:
        B_, h, N, _ = a.shape
        a = (a.view(B_ // nW, nW, h, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(B_, h, N, N)
    return torch.softmax(a, dim=-1) @ v


def line_window_attention(xw, x_ref, p, nheads, ws, mask):
    """WindowAttention.forward, mst.py:267-332: queries are re-expressed through the line end-point
    tokens ("glass-structure context") before the ordinary window attention.  q is scaled twice."""
    B_, N, C = xw.shape
    hd = C // nheads
    scale = hd ** -0.5
    qkv = linear(xw, p, "qkv").reshape(B_, N, 3, nheads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    rB, n_rf, _ = x_ref.shape
    n_win = B_ // rB
    ref = linear(x_ref, p, "ref_qk").reshape(rB, n_rf, 2, C)
    ref_k = p["diff_mu"] + p["diff_logsigma"].exp() * ref[:, :, 0]          # mst.py:286-288
    ref_v = ref[:, :, 1]
    ref_k = ref_k.reshape(rB, n_rf, nheads, hd).permute(0, 2, 1, 3).repeat_interleave(n_win, dim=0)
    ref_v = ref_v.reshape(rB, n_rf, nheads, hd).permute(0, 2, 1, 3).repeat_interleave(n_win, dim=0)
    q = q * scale
    ra = q @ ref_k.transpose(-2, -1)                                         # [B_, h, N, n_rf]
    ra = ra.view(rB, n_win, nheads, N, n_rf).permute(0, 2, 1, 3, 4).reshape(rB, nheads, n_win * N, n_rf)
    for _ in range(3):                                                       # mst.py:299-302
        upd = F.conv2d(ra, p["ref_attn_diffusion.weight"], p["ref_attn_diffusion.bias"], padding=1)
        ra = ra + F.gelu(F.layer_norm(upd, [n_win * N, n_rf]))
    ra = ra.reshape(rB, nheads, n_win, N, n_rf).permute(0, 2, 1, 3, 4).reshape(B_, nheads, N, n_rf)
    q_new = (torch.softmax(ra, dim=-1) @ ref_v) * scale                      # second scaling, mst.py:310
    out = softmax_attention(q_new, k, v, relative_position_bias(p, ws, nheads), mask, mask.shape[0] if mask is not None else 1)
    return linear(out.transpose(1, 2).reshape(B_, N, C), p, "proj")


def class_window_attention(xw, dtok, stok, p, nheads, ws, mask):
    """WindowClassAttention.forward with group_attention=False, mst.py:455-580.  The segmentation
    token is projected with proj_dth as well (mst.py:578), exactly like the reference."""
    B_, N, C = xw.shape
    hd = C // nheads
    scale = hd ** -0.5
    qkv = linear(xw, p, "qkv").reshape(B_, N, 3, nheads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    out = softmax_attention(q, k, v, relative_position_bias(p, ws, nheads), mask, mask.shape[0] if mask is not None else 1)
    x = linear(out.transpose(1, 2).reshape(B_, N, C), p, "proj")
    tdim = dtok.shape[-1]
    t_x = torch.cat([x, dtok, stok], dim=-1)
    tC = t_x.shape[-1]
    t_k = linear(t_x, p, "global_k").reshape(B_, N, nheads, tC // nheads).permute(0, 2, 1, 3)
    t_v = linear(t_x, p, "global_v").reshape(B_, N, nheads, tC // nheads).permute(0, 2, 1, 3)

    def channel_attention(tok, qname):
        tq = linear(tok, p, qname).reshape(B_, N, nheads, tdim // nheads).permute(0, 2, 1, 3) * scale
        a = torch.softmax(tq.transpose(-2, -1) @ t_k, dim=-1)               # [B_, h, tdim/h, tC/h]
        t = (a @ t_v.transpose(-2, -1)).reshape(B_, -1, N).permute(0, 2, 1)
        return linear(t, p, "proj_dth")

    return x, channel_attention(dtok, "cls_dth_q"), channel_attention(stok, "cls_seg_q")


def nearest_points(feat_nchw, grid):
    return F.grid_sample(feat_nchw, grid, mode="nearest", align_corners=False)


def line_swin_block(x, H, W, ref, ref_pos, p, nheads, ws, shift, mask):
    """SwinTransformerBlock.forward with WindowAttention, mst.py:646-755"""
    B, L, C = x.shape
    xs = pad_and_shift(layer_norm(x, p, "norm1"), H, W, ws, shift)
    Hp, Wp = xs.shape[1:3]
    if shift > 0:                                                            # mst.py:680-686
        rr = torch.zeros_like(ref)
        rr[..., 0] = ref[..., 0] - (shift / (Wp - 1)) * 2
        rr[..., 1] = ref[..., 1] - (shift / (Hp - 1)) * 2
        rr = torch.where(rr < -1, -1 - (1 + rr), rr)
        rpos = torch.roll(ref_pos, shifts=(-shift, -shift), dims=(2, 3))
    else:
        rr, rpos = ref, ref_pos
    x_ref = nearest_points(xs.permute(0, 3, 1, 2), rr) + nearest_points(rpos, rr)   # mst.py:694-697
    x_ref = x_ref.reshape(B, C, -1).permute(0, 2, 1)
    aw = line_window_attention(to_windows(xs, ws), x_ref, p.sub("attn"), nheads, ws, mask if shift > 0 else None)
    xa = unshift_and_crop(from_windows(aw, ws, B, Hp, Wp), H, W, shift).view(B, L, C)
    x = x + xa
    return x + mlp(layer_norm(x, p, "norm2"), p, "mlp")


def class_swin_block(x, dtok, stok, H, W, p, nheads, ws, shift, mask):
    """SwinTransformerBlock.forward with WindowClassAttention, mst.py:646-788 (x_ref is dead there)"""
    B, L, C = x.shape
    tC = dtok.shape[-1]
    xs = pad_and_shift(layer_norm(x, p, "norm1"), H, W, ws, shift)
    ds = pad_and_shift(layer_norm(dtok, p, "norm_depth1"), H, W, ws, shift)
    ss = pad_and_shift(layer_norm(stok, p, "norm_seg1"), H, W, ws, shift)
    Hp, Wp = xs.shape[1:3]
    aw, dw, sw = class_window_attention(to_windows(xs, ws), to_windows(ds, ws), to_windows(ss, ws), p.sub("attn"),
                                        nheads, ws, mask if shift > 0 else None)
    xa = unshift_and_crop(from_windows(aw, ws, B, Hp, Wp), H, W, shift).view(B, L, C)
    x = x + xa
    x = x + mlp(layer_norm(x, p, "norm2"), p, "mlp")
    d = dtok + unshift_and_crop(from_windows(dw, ws, B, Hp, Wp), H, W, shift).view(B, L, tC)
    d = d + mlp(layer_norm(d, p, "norm_depth2"), p, "mlp_depth")
    s = stok + unshift_and_crop(from_windows(sw, ws, B, Hp, Wp), H, W, shift).view(B, L, tC)
    s = s + mlp(layer_norm(s, p, "norm_seg2"), p, "mlp_seg")
    return x, d, s


def swin_stage(x, H, W, p, depth, nheads, ws, ref=None, ref_pos=None, dtok=None, stok=None):
    """BasicLayer.forward, mst.py:926-979: blocks alternate shift 0 / ws//2"""
    Hp, Wp = math.ceil(H / ws) * ws, math.ceil(W / ws) * ws
    mask = shift_window_mask(Hp, Wp, ws, ws // 2, x.device)
    for i in range(depth):
        shift = 0 if i % 2 == 0 else ws // 2
        bp = p.sub("blocks.%d" % i)
        if dtok is None:
            x = line_swin_block(x, H, W, ref, ref_pos, bp, nheads, ws, shift, mask)
        else:
            x, dtok, stok = class_swin_block(x, dtok, stok, H, W, bp, nheads, ws, shift, mask)
    return x, dtok, stok


# ----------------------------------------------------------------------------------------------
# uncertainty sampling + point-anchored depth          src/models/points/points_sample.py
# ----------------------------------------------------------------------------------------------
def conv_ln(x, p, name, padding=1):  # points_sample.py:12-25
    y = F.conv2d(x, p[name + ".conv.weight"], None, 1, padding)
    w = p[name + ".layer_norm.weight"]
    return F.layer_norm(y.permute(0, 2, 3, 1), (w.numel(),), w, p[name + ".layer_norm.bias"], 1e-5).permute(0, 3, 1, 2)


def pyramid(x, p, pool_sizes=(16, 8, 4, 2)):
    """PyramidLayer.forward, points_sample.py:106-125 (layer4 is built but never run)"""
    x = F.gelu(conv_ln(x, p, "firstconv.0"))
    x = F.gelu(conv_ln(x, p, "firstconv.2"))
    for lname, nblk in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        for b in range(nblk):
            bp = "%s.%d" % (lname, b)
            x = conv_ln(F.gelu(conv_ln(x, p, bp + ".conv1.0")), p, bp + ".conv2") + x   # BasicBlock, :37-43
    Hh, Ww = x.shape[-2:]
    if Hh < pool_sizes[0] or Ww < pool_sizes[0]:                                         # pad_before_pool, :94-104
        x = F.pad(x, (0, max(pool_sizes[0] - Ww, 0), 0, max(pool_sizes[0] - Hh, 0)))
    outs = [x]
    for i, ps in enumerate(pool_sizes, start=1):
        b = F.gelu(conv_ln(F.avg_pool2d(x, ps, ps), p, "branch%d.1" % i))
        outs.append(F.interpolate(b, size=x.shape[-2:], mode="bilinear", align_corners=True))
    y = F.gelu(conv_ln(torch.cat(outs, dim=1), p, "lastconv.0"))
    return F.conv2d(y, p["lastconv.2.weight"])


def point_based_pred(x, dtok, pre_depth, coords, H, W, pos, p, dim):
    """PointBasedPred.forward, points_sample.py:257-280 (correlation scaled by dim**-2, :273)"""
    xg_xr = linear(linear(torch.cat([x, dtok], dim=-1), p, "pre_proj"), p, "refer_proj")
    xg, xr = xg_xr[:, :, :dim], xg_xr[:, :, dim:]
    B = x.shape[0]
    xr = xr.permute(0, 2, 1).reshape(B, -1, H, W)
    refer = (F.grid_sample(xr, coords, align_corners=False) + F.grid_sample(pos, coords, align_corners=False)).flatten(2)
    anchor = F.grid_sample(pre_depth, coords, align_corners=False).permute(0, 2, 1, 3)   # [B,K,1,1]
    rg = (xg @ refer) * (dim ** -2)
    rg = rg.permute(0, 2, 1).reshape(B, -1, H, W)
    attn = torch.softmax(pyramid(rg, p.sub("pyramid")), dim=1)
    return (attn * anchor).sum(dim=1, keepdim=True)


def certain_sample(pred_small, pred_large, sample_num, interval, min_depth):
    """CertainSample.forward, points_sample.py:291-364.  Returns (coords [B,K,1,2] in [-1,1), idx [B,K] int64
    flat pixel indices y*W+x).  Every depth bin takes a prefix of the SAME global descending-variance list
    (the bin mask is not applied to the top-k, :319) and sorts it by index (:320)."""
    B, _, H, W = pred_large.shape
    variance = (F.interpolate(pred_small, size=(H, W), mode="bilinear", align_corners=True) - pred_large) ** 2
    edges = [min_depth] + list(interval) + [1.0]
    all_idx = []
    for b in range(B):
        flat = variance[b].flatten(0)
        picks, counts, already = [], [], 0
        for i in range(len(edges) - 1):
            inside = (pred_large[b] >= edges[i]) & (pred_large[b] < edges[i + 1])
            total = torch.sum(inside)
            n_i = int(torch.min(torch.floor((total / (H * W)) * sample_num), total))
            if n_i > 0:
                picks.append(torch.topk(flat, n_i).indices.sort().values)
                counts.append(n_i)
                already += n_i
        if picks:
            idx = torch.cat(picks)
            remain = sample_num - already
        else:
            idx = torch.topk(flat, sample_num).indices.sort().values
            remain = 0
        if remain > 0 and remain >= already:                                             # :343-346
            times = remain // already + 1
            idx = idx.repeat(times)
            remain = sample_num - already * times
        if remain > 0:                                                                   # :348-350
            idx = torch.cat([idx, idx[-remain:]])
        if remain < 0:                                                                   # :351-355
            m = int(torch.argmax(torch.tensor(counts)))  # noqa: host list, as in the reference
            picks[m] = picks[m][:remain]
            idx = torch.cat(picks)
        all_idx.append(idx)
    idx = torch.stack(all_idx)
    col = (idx % W).float()
    row = torch.div(idx, W, rounding_mode="floor").float()
    coords = torch.stack([col / W * 2 - 1, row / H * 2 - 1], dim=-1)[:, :, None, :]      # :361-363
    return coords, idx


# ----------------------------------------------------------------------------------------------
# dense encoder ("ReferTransformer")                 src/models/multiscale_transformerr.py:1151-1319
# ----------------------------------------------------------------------------------------------
def select_reference_points(pred_lines, pred_logits, num_ref, with_center):
    """mst.py:1165-1179: top-num_ref by RAW line logit, end points (and centre) mapped to [-1,1]"""
    B = pred_lines.shape[0]
    ids = torch.topk(pred_logits[:, :, 0], num_ref, dim=-1).indices
    pts = torch.stack([pred_lines[i][ids[i]] for i in range(B)]).reshape(B, num_ref, -1, 2) * 2 - 1.0
    return (pts if with_center else pts[:, :, :2]), ids


def conv_a(x, p, name):  # ConvA, mst.py:104-118 (3x3 conv with bias + GELU)
    return F.gelu(F.conv2d(x, p[name + ".conv.weight"], p[name + ".conv.bias"], 1, 1))


def mlp_norm(x, p, name):  # MlpNorm without activation, mst.py:75-102
    return layer_norm(linear(linear(x, p, name + ".fc1"), p, name + ".fc2"), p, name + ".norm")


def depth_head(x, p, name):  # nn.Sequential(Linear, Linear, Sigmoid), mst.py:1044-1045
    return torch.sigmoid(linear(linear(x, p, name + ".0"), p, name + ".1"))


def up_tokens(tok, Hs, Ws, size):
    B, _, C = tok.shape
    t = tok.reshape(B, Hs, Ws, C).permute(0, 3, 1, 2)
    return F.interpolate(t, size=size, mode="nearest")


def dense_encoder(c5_proj, mask5, feats, masks, pred_lines, pred_logits, p, cfg, pinned=None, trace=None):
    """ReferTransformer.forward.  feats = [C2,C3,C4] backbone maps, masks their padding masks.
    pinned: optional dict {'line_ids','sample1','sample2'} overriding the discrete selections."""
    B, C, H, W = c5_proj.shape
    ws, heads = cfg["window"], cfg["dense_trans_heads"]
    pinned = pinned or {}
    ref, ids = select_reference_points(pred_lines, pred_logits, cfg["num_ref"], cfg["with_dense_center"])
    if "line_ids" in pinned:
        ids = pinned["line_ids"]
        pts = torch.stack([pred_lines[i][ids[i]] for i in range(B)]).reshape(B, cfg["num_ref"], -1, 2) * 2 - 1.0
        ref = pts if cfg["with_dense_center"] else pts[:, :, :2]
    D = cfg["dense_trans_dim"]
    pos32 = sine_position(mask5, D // 2, False)
    x32, _, _ = swin_stage(c5_proj.flatten(2).permute(0, 2, 1), H, W, p.sub("dense_transformer"),
                           cfg["dense_trans_layers"][0], heads, ws, ref=ref, ref_pos=pos32)
    depth0 = depth_head(x32, p, "depth_pred32").permute(0, 2, 1).reshape(B, -1, H, W)
    dense_out = x32.permute(0, 2, 1).reshape(B, C, H, W)
    # ---- 1/16 --------------------------------------------------------------------------- mst.py:1191-1215
    H1, W1 = feats[2].shape[-2:]
    up = F.interpolate(dense_out, size=(H1, W1), mode="nearest")
    x = linear(up.flatten(2).permute(0, 2, 1), p, "proj_class1") + conv_a(feats[2], p, "proj_backbn1").flatten(2).permute(0, 2, 1)
    dtok = p["depth_token"].expand(B, H1 * W1, -1)
    stok = p["seg_token"].expand(B, H1 * W1, -1)
    x1, dtok, stok = swin_stage(x, H1, W1, p.sub("class_transformer1"), cfg["class_trans_layers"][0], heads, ws, dtok=dtok, stok=stok)
    depth1 = depth_head(torch.cat([x1, dtok], dim=-1), p, "depth_pred16").permute(0, 2, 1).reshape(B, -1, H1, W1)
    pts1, idx1 = certain_sample(depth0, depth1, cfg["interval_sample_num"][0], cfg["depth_interval"],
                                cfg["min_depth_eval"] / cfg["max_depth_eval"])
    if "sample1" in pinned:
        pts1, idx1 = pinned["sample1"]
    # ---- 1/8 ---------------------------------------------------------------------------- mst.py:1226-1260
    H2, W2 = feats[1].shape[-2:]
    f1 = x1.permute(0, 2, 1).reshape(B, -1, H1, W1)
    up = F.interpolate(f1, size=(H2, W2), mode="nearest")
    x = linear(up.flatten(2).permute(0, 2, 1), p, "proj_class2") + conv_a(feats[1], p, "proj_backbn2").flatten(2).permute(0, 2, 1)
    pos8 = sine_position(masks[1], D // 8, False)
    dtok = mlp_norm(up_tokens(dtok, H1, W1, (H2, W2)).flatten(2).permute(0, 2, 1), p, "old_depth_token_proj8")
    stok = mlp_norm(up_tokens(stok, H1, W1, (H2, W2)).flatten(2).permute(0, 2, 1), p, "old_seg_token_proj8")
    x2, dtok, stok = swin_stage(x, H2, W2, p.sub("class_transformer2"), cfg["class_trans_layers"][1], heads, ws, dtok=dtok, stok=stok)
    depth2 = point_based_pred(x2, dtok, depth1, pts1, H2, W2, pos8, p.sub("point_based_pred1"), D // 4)
    pts2, idx2 = certain_sample(depth1, depth2, cfg["interval_sample_num"][1], cfg["depth_interval"],
                                cfg["min_depth_eval"] / cfg["max_depth_eval"])
    if "sample2" in pinned:
        pts2, idx2 = pinned["sample2"]
    # ---- 1/4 ---------------------------------------------------------------------------- mst.py:1263-1289
    H3, W3 = feats[0].shape[-2:]
    f2 = x2.permute(0, 2, 1).reshape(B, -1, H2, W2)
    up = F.interpolate(f2, size=(H3, W3), mode="nearest")
    x = linear(up.flatten(2).permute(0, 2, 1), p, "proj_class3") + conv_a(feats[0], p, "proj_backbn3").flatten(2).permute(0, 2, 1)
    pos4 = sine_position(masks[0], D // 16, False)
    dtok = mlp_norm(up_tokens(dtok, H2, W2, (H3, W3)).flatten(2).permute(0, 2, 1), p, "old_depth_token_proj4")
    stok = mlp_norm(up_tokens(stok, H2, W2, (H3, W3)).flatten(2).permute(0, 2, 1), p, "old_seg_token_proj4")
    x3, dtok, stok = swin_stage(x, H3, W3, p.sub("class_transformer3"), cfg["class_trans_layers"][2], heads, ws, dtok=dtok, stok=stok)
    depth3 = point_based_pred(x3, dtok, depth2, pts2, H3, W3, pos4, p.sub("point_based_pred2"), D // 8)
    feat4 = x3.permute(0, 2, 1).reshape(B, -1, H3, W3)
    dt4 = dtok.permute(0, 2, 1).reshape(B, -1, H3, W3)
    st4 = stok.permute(0, 2, 1).reshape(B, -1, H3, W3)
    if trace is not None:
        trace.update(line_ids=ids, ref_points=ref, x32=x32, depth0=depth0, x1=x1, sample1_idx=idx1, sample1=pts1,
                     x2=x2, sample2_idx=idx2, sample2=pts2, x3=x3, depth_token4=dt4, seg_token4=st4)
    return feat4, dt4, st4, [depth1, depth2, depth3]


# ----------------------------------------------------------------------------------------------
# dense prediction head                                     src/models/dense_upsample.py:74-182
# ----------------------------------------------------------------------------------------------
def upconv(x, w, size=None):  # dense_upsample.py:82-90: nearest up -> 3x3 conv (no bias) -> ELU
    up = F.interpolate(x, scale_factor=2, mode="nearest") if size is None else F.interpolate(x, size=size, mode="nearest")
    return F.elu(F.conv2d(up, w, None, 1, 1))


def dense_head(feat4, depth3, dt4, st4, size, p, max_depth):
    def branch(fused, kind):
        B, _, H, W = fused.shape
        f = mlp(fused.flatten(2).permute(0, 2, 1), p, kind + "_token_fuse").permute(0, 2, 1).reshape(B, -1, H, W)
        u1 = upconv(f, p["upconv1_%s.conv.weight" % kind])
        u1 = layer_norm(u1.permute(0, 2, 3, 1), p, "norm_" + kind).permute(0, 3, 1, 2)
        c1 = F.elu(F.conv2d(u1, p["conv1_%s.0.weight" % kind], None, 1, 1))
        u2 = upconv(c1, p["upconv2_%s.conv.weight" % kind], size)
        return F.elu(F.conv2d(u2, p["conv2_%s.0.weight" % kind], None, 1, 1))
    d = branch(torch.cat([feat4, depth3, dt4], dim=1), "depth")
    depth = max_depth * torch.sigmoid(F.conv2d(d, p["get_depth.0.weight"], None, 1, 1))
    s = branch(torch.cat([feat4, st4], dim=1), "seg")
    seg = F.conv2d(s, p["get_seg.weight"], None, 1, 1)
    return depth, seg


# ----------------------------------------------------------------------------------------------
# full model                                                    src/models/glassrgbd.py:74-123
# ----------------------------------------------------------------------------------------------
def forward(sd, images, mask=None, cfg=None, pinned=None, trace=None, grad=False):
    """images [B,3,H,W] float32 (already padded to a common size), mask [B,H,W] bool (True = pad) or None.
    Returns the reference's output dict: pred_logits, pred_lines, aux_outputs, pred_depth (list of 4), pred_seg.
    grad=True keeps the autograd graph (weights with requires_grad: the gradient oracle of the training tests)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    p = P(sd)
    B, _, H, W = images.shape
    if mask is None:
        mask = torch.zeros(B, H, W, dtype=torch.bool, device=images.device)
    with torch.set_grad_enabled(grad):
        feats = resnet50_features(images, p.sub("backbone.0.body"))
        masks = [downsample_mask(mask, f.shape[-2:]) for f in feats]
        c5, m5 = feats[3], masks[3]
        pos5 = sine_position(m5, cfg["hidden_dim"] // 2, True)
        src = F.conv2d(c5, p["input_proj.weight"], p["input_proj.bias"])
        hs, memory = detr_transformer(src, m5, p["query_embed.weight"], pos5, p.sub("transformer"), cfg)
        logits = linear(hs, p, "class_embed")
        h = hs
        for i in range(3):
            h = linear(h, p, "lines_embed.layers.%d" % i)
            if i < 2:
                h = F.relu(h)
        lines = h.sigmoid()
        out = {"pred_logits": logits[-1], "pred_lines": lines[-1]}
        if cfg["aux_loss"]:
            out["aux_outputs"] = [{"pred_logits": a, "pred_lines": b} for a, b in zip(logits[:-1], lines[:-1])]
        dense_in = F.conv2d(c5, p["dense_input_proj.weight"], p["dense_input_proj.bias"])
        feat4, dt4, st4, depths = dense_encoder(dense_in, m5, feats[:3], masks[:3], out["pred_lines"], out["pred_logits"],
                                                p.sub("dense_encoder"), cfg, pinned, trace)
        depth, seg = dense_head(feat4, depths[-1], dt4, st4, (H, W), p.sub("depth_decoder"), cfg["max_depth"])
        out["pred_depth"] = depths + [depth]
        out["pred_seg"] = seg
        if trace is not None:
            trace.update(c5=c5, memory=memory, hs=hs, dense_in=dense_in, feat4=feat4)
    return out


# ----------------------------------------------------------------------------------------------
# criterion side: matcher cost, set loss, depth / seg losses, eval metrics
# ----------------------------------------------------------------------------------------------
def matcher_cost(pred_logits, pred_lines, tgt_lines_list, cost_class=1.0, cost_line=5.0):
    """matcher.py:52-71 restricted to the block diagonal the reference actually uses (:73-74):
    list over images of [Q, T_b] cost matrices  C = cost_line * L1(lines) - cost_class * p(line)."""
    prob = pred_logits.softmax(-1)
    out = []
    for b, tgt in enumerate(tgt_lines_list):
        l1 = torch.cdist(pred_lines[b], tgt, p=1)
        out.append(cost_line * l1 + cost_class * (-prob[b][:, :1].expand(-1, tgt.shape[0])))
    return out


def hungarian(costs):
    """matcher.py:74: scipy.optimize.linear_sum_assignment per image (third-party, scipy 1.18.1 here)."""
    from scipy.optimize import linear_sum_assignment
    res = []
    for c in costs:
        i, j = linear_sum_assignment(c.detach().cpu().numpy())
        res.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
    return res


def set_losses(pred_logits, pred_lines, tgt_lines_list, indices, num_items, eos_coef=0.1):
    """glassrgbd.py:154-175 (weighted CE, classes {0: line, 1: no-object}) and :231-244 (L1 sum / num_items)."""
    B, Q, _ = pred_logits.shape
    target = torch.full((B, Q), 1, dtype=torch.int64)
    src, tgt = [], []
    for b, (i, j) in enumerate(indices):
        target[b, i] = 0
        src.append(pred_lines[b][i])
        tgt.append(tgt_lines_list[b][j])
    w = torch.tensor([1.0, eos_coef])
    loss_ce = F.cross_entropy(pred_logits.transpose(1, 2), target, w)
    loss_line = F.l1_loss(torch.cat(src), torch.cat(tgt), reduction="none").sum() / num_items
    return loss_ce, loss_line


def set_criterion(out, tgt_lines_list, cfg=None):
    """SetCriterion.forward, glassrgbd.py:308-358 (single process: world size 1)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    num_items = max(float(sum(t.shape[0] for t in tgt_lines_list)), 1.0)
    losses = {}
    stages = [(out["pred_logits"], out["pred_lines"], "")]
    for i, aux in enumerate(out.get("aux_outputs", [])):
        stages.append((aux["pred_logits"], aux["pred_lines"], "_%d" % i))
    all_indices = []
    for logits, lines, suffix in stages:
        idx = hungarian(matcher_cost(logits, lines, tgt_lines_list, cfg["set_cost_class"], cfg["set_cost_line"]))
        all_indices.append(idx)
        ce, l1 = set_losses(logits, lines, tgt_lines_list, idx, num_items, cfg["eos_coef"])
        losses["loss_ce" + suffix] = ce
        losses["loss_line" + suffix] = l1
    return losses, all_indices


def silog_loss(pred, gt, mask, variance_focus=0.85, log_depth_error=False):  # glassrgbd.py:366-374
    if log_depth_error:
        d = torch.log(pred[mask]) - torch.log(gt[mask])
    else:
        d = (pred[mask] + torch.log(pred[mask])) - (gt[mask] + torch.log(gt[mask]))
    return torch.sqrt((d ** 2).mean() - variance_focus * (d.mean() ** 2)) * 10.0


def depth_losses(pred_depth_list, depth_gt, weights=(0.25, 0.25, 0.25, 1.0), **kw):
    """engine_glassrgbd.py:65-82: gt and validity mask nearest-resized to every prediction scale"""
    mask = (depth_gt >= 0.2) & (depth_gt < 10.0)
    out = []
    for w, pd in zip(weights, pred_depth_list):
        size = pd.shape[-2:]
        g = F.interpolate(depth_gt, size=size, mode="nearest")
        m = F.interpolate(mask.to(torch.uint8), size=size, mode="nearest").to(torch.bool)
        out.append(silog_loss(pd, g, m, **kw) * w)
    return out


def seg_loss(pred_seg, seg_gt, weight=2.0):  # glassrgbd.py:376-383, engine_glassrgbd.py:88-90
    return F.cross_entropy(pred_seg, seg_gt.squeeze(1)) * weight


def depth_metrics(pred, gt, min_depth=1e-3, max_depth=10.0):
    """engine_glassrgbd.py:243-264 + util/metrics.py:197-218 for ONE image: -> 9 float64 values
    [silog, abs_rel, log10, rms, sq_rel, log_rms, d1, d2, d3] (numpy float32 data, float64 means)."""
    import numpy as np
    pred = pred.detach().cpu().numpy().astype(np.float32).squeeze().copy()
    gt = gt.detach().cpu().numpy().astype(np.float32).squeeze()
    pred[pred < min_depth] = min_depth
    pred[pred > max_depth] = max_depth
    pred[np.isinf(pred)] = max_depth
    pred[np.isnan(pred)] = min_depth
    valid = np.logical_and(gt > min_depth, gt < max_depth)
    g, q = gt[valid], pred[valid]
    thresh = np.maximum(g / q, q / g)
    d1, d2, d3 = (thresh < 1.25).mean(), (thresh < 1.25 ** 2).mean(), (thresh < 1.25 ** 3).mean()
    rms = np.sqrt(((g - q) ** 2).mean())
    log_rms = np.sqrt(((np.log(g) - np.log(q)) ** 2).mean())
    abs_rel = np.mean(np.abs(g - q) / g)
    sq_rel = np.mean(((g - q) ** 2) / g)
    err = np.log(q) - np.log(g)
    silog = np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100
    log10 = np.mean(np.abs(np.log10(q) - np.log10(g)))
    return [float(v) for v in (silog, abs_rel, log10, rms, sq_rel, log_rms, d1, d2, d3)]
