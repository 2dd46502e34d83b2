"""Torch-tensor front-ends of the C-ABI kernels (raw device pointers + sizes go down, nothing else).

Tensors are channels-last bf16 device buffers: a feature map is [B, H, W, C] and a token
matrix is [rows, C] (treated as B=1, H=1, W=rows).  Every function enqueues on torch's
current CUDA stream.  There is no fallback path: a missing library or a failing call raises.
"""
import ctypes

import torch

from . import capi
from .capi import (ACT_ELU, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID,  # noqa: F401
                   RES_AFTER, RES_BEFORE_NORM, RES_NONE)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def round_up(v, m):
    return (v + m - 1) // m * m


# bench.py sets this to a list to time every gwd_conv_gemm launch with CUDA events on the launching stream:
# entries are (start_event, end_event, algorithmic_flops, description)
PROFILE = None


# ------------------------------------------------------------------------------------------
# weight packing (done once at model-load time; not part of the hot path)
# ------------------------------------------------------------------------------------------
class PackedWeight:
    """bf16 [taps][n_pad][cin_pad] K-major weights + fp32 [n_pad] bias for gwd_conv_gemm."""

    def __init__(self, w, bias, taps, n, cin):
        self.w, self.bias, self.taps, self.n, self.cin = w, bias, taps, n, cin
        self.n_pad, self.cin_pad = w.shape[1], w.shape[2]


def pack_linear(weight, bias=None, cin_pad=None, col_map=None):
    """weight [N, K] (nn.Linear layout).  `col_map`: optional LongTensor [K] giving the physical input
    channel of every logical input column (used when the input buffer has padded channel groups)."""
    n, k = weight.shape
    n_pad = round_up(n, 16)
    cin_pad = cin_pad or round_up(k, 16)
    w = torch.zeros(1, n_pad, cin_pad, dtype=torch.float32, device=weight.device)
    if col_map is None:
        w[0, :n, :k] = weight.float()
    else:
        w[0, :n, col_map.to(weight.device)] = weight.float()
    b = torch.zeros(n_pad, dtype=torch.float32, device=weight.device)
    if bias is not None:
        b[:n] = bias.float()
    return PackedWeight(w.to(torch.bfloat16).contiguous(), b, 1, n, cin_pad)


def pack_conv3x3(weight, bias=None, cin_pad=None, col_map=None):
    """weight [N, C, 3, 3] (nn.Conv2d layout) -> tap-major [dx*3+dy][n_pad][cin_pad]."""
    n, c = weight.shape[:2]
    n_pad = round_up(n, 16)
    cin_pad = cin_pad or round_up(c, 16)
    w = torch.zeros(9, n_pad, cin_pad, dtype=torch.float32, device=weight.device)
    wf = weight.float()
    for dy in range(3):
        for dx in range(3):
            if col_map is None:
                w[dx * 3 + dy, :n, :c] = wf[:, :, dy, dx]
            else:
                w[dx * 3 + dy, :n, col_map.to(weight.device)] = wf[:, :, dy, dx]
    b = torch.zeros(n_pad, dtype=torch.float32, device=weight.device)
    if bias is not None:
        b[:n] = bias.float()
    return PackedWeight(w.to(torch.bfloat16).contiguous(), b, 9, n, cin_pad)


def pack_upconv3x3(weight):
    """`upconv` (nearest x2 up-sampling followed by a bias-free 3x3 conv, src/models/dense_upsample.py:74-90) as ONE 3x3
    conv on the low-resolution input with 4*Cout outputs: output pixel (2y+oy, 2x+ox) only sees low-res rows
    {y-1, y} (oy = 0) or {y, y+1} (oy = 1), so the taps that land on the same low-res pixel are summed (in fp32).
    Phase p = 2*oy + ox owns output channels [p*Cout, (p+1)*Cout)."""
    n, c = weight.shape[:2]
    assert n % 16 == 0 and c % 16 == 0
    wf = weight.float()
    # rows of the 3x3 filter (a = 0,1,2 <-> up-sampled rows 2y+oy-1 .. 2y+oy+1) -> low-res tap dy in {0,1,2} (= y-1,y,y+1)
    taps = {0: {0: [0], 1: [1, 2], 2: []}, 1: {0: [], 1: [0, 1], 2: [2]}}
    full = torch.zeros(4 * n, c, 3, 3, dtype=torch.float32, device=weight.device)
    for oy in (0, 1):
        for ox in (0, 1):
            ph = 2 * oy + ox
            for dy in range(3):
                for dx in range(3):
                    for a in taps[oy][dy]:
                        for b in taps[ox][dx]:
                            full[ph * n:(ph + 1) * n, :, dy, dx] += wf[:, :, a, b]
    pw = pack_conv3x3(full)
    pw.upsample2 = True
    return pw


def pad_vec(v, n_pad, fill=0.0):
    out = torch.full((n_pad,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.float()
    return out


# ------------------------------------------------------------------------------------------
# gwd_conv_gemm
# ------------------------------------------------------------------------------------------
def _as_bhwc(t):
    if t.dim() == 2:
        return 1, 1, t.shape[0], t.shape[1]
    if t.dim() == 3:  # [B, L, C] tokens
        return 1, 1, t.shape[0] * t.shape[1], t.shape[2]
    assert t.dim() == 4
    return tuple(t.shape)


def conv_gemm(x, pw, *, x_coff=0, out=None, y_coff=0, out_channels=None, out_f32=False,
              bias=True, ln=None, ln_eps=1e-5, pre_act=ACT_NONE, post_act=ACT_NONE, out_scale=1.0,
              res=None, res_coff=0, res_mode=RES_NONE, y_raw=None, w_per_image=False, subsample2=False, act_grad=None):
    """y = epilogue(conv_or_linear(x[..., x_coff:x_coff+cin], pw)).  See include/gwd_b200.h.

    x      : bf16 channels-last [B,H,W,Cx] or [rows,Cx] / [B,L,Cx]
    pw     : PackedWeight
    ln     : None or (gamma_pad, beta_pad) fp32 [n_pad]
    out    : optional preallocated output (channels-last, same leading dims); the result is
             written at channel offset y_coff.  Otherwise a [..., out_channels or n_pad(8-aligned n)] buffer is made.
    """
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.is_contiguous()
    B, H, W, Cx = _as_bhwc(x)
    taps = pw.taps
    strides = None
    if subsample2:      # Linear over x[:, ::2, ::2, :] without materialising the view (TMA walks the strided pixels)
        assert taps == 1 and x.dim() == 4 and out is None
        strides = (2, 2 * W, H * W)
        H, W = (H + 1) // 2, (W + 1) // 2
    if taps == 9:
        assert x.dim() == 4
    # bf16 outputs carry all n_pad physical channels (pads are written as exact zeros) so that the next layer can
    # consume a 16-aligned K; fp32 outputs (small heads) carry the logical channels only
    store_n = pw.n if out_f32 else min(pw.n_pad, out_channels or pw.n_pad)
    up2 = bool(getattr(pw, "upsample2", False))
    if out is None:
        oc = out_channels or store_n
        shape = ((B, H, W) if subsample2 else tuple(x.shape[:-1])) + (oc,)
        if up2:     # fused nearest x2 up-sampling: [B,H,W,C] -> [B,2H,2W,Cout]
            shape = (B, 2 * H, 2 * W, pw.n // 4)
        out = torch.empty(shape, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    assert out.is_contiguous()
    d = capi.GemmDesc()
    d.x = x.data_ptr(); d.B, d.H, d.W = B, H, W
    if taps == 1 and not subsample2 and not w_per_image:
        d.B, d.H, d.W = 1, 1, B * H * W      # a dense Linear is one row axis (lets the kernel use its TMA epilogue)
    d.x_cstride, d.x_coff, d.cin = Cx, x_coff, pw.cin_pad
    d.w = pw.w.data_ptr(); d.taps = taps; d.n_pad = pw.n_pad; d.n = pw.n
    d.bias = pw.bias.data_ptr() if (bias and pw.bias is not None) else None
    if ln is not None:
        d.ln_g, d.ln_b = ln[0].data_ptr(), ln[1].data_ptr()
    d.ln_eps = ln_eps
    d.pre_act, d.post_act, d.out_scale = pre_act, post_act, out_scale
    if act_grad is not None:      # (saved activation, act, from_input, y_mul, scale): result *= act'(saved) * scale  (training)
        assert res is None and ln is None and pre_act == ACT_NONE and post_act == ACT_NONE and not out_f32
        ag_y, ag_act, ag_from_input, ag_y_mul, ag_scale = act_grad
        assert ag_y.dtype == torch.bfloat16 and ag_y.is_contiguous()
        d.res = ag_y.data_ptr(); d.res_cstride = ag_y.shape[-1]; d.res_coff = 0; d.res_mode = capi.RES_MUL_ACTGRAD
        d.ag_act, d.ag_from_input, d.ag_y_mul, d.ag_scale = int(ag_act), int(bool(ag_from_input)), float(ag_y_mul), float(ag_scale)
    if res is not None:
        assert res.dtype == torch.bfloat16 and res.is_contiguous() and res_mode != RES_NONE
        d.res = res.data_ptr(); d.res_cstride = res.shape[-1]; d.res_coff = res_coff; d.res_mode = res_mode
    d.y = out.data_ptr(); d.y_cstride = out.shape[-1]; d.y_coff = y_coff; d.y_f32 = 1 if out_f32 else 0
    if y_raw is not None:
        d.y_raw = y_raw.data_ptr(); d.yraw_cstride = y_raw.shape[-1]; d.yraw_coff = 0
    d.store_n = store_n
    d.w_per_image = 1 if w_per_image else 0
    d.upsample2 = 1 if up2 else 0
    if strides is not None:
        d.x_wstride, d.x_hstride, d.x_bstride = strides
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.check(capi.lib().gwd_conv_gemm(ctypes.byref(d), _stream()), "gwd_conv_gemm")
        e1.record()
        PROFILE.append((e0, e1, 2.0 * B * H * W * pw.n * pw.cin * taps,
                        "%dx%dx%d pixels, %d->%d ch, %d taps" % (B, H, W, pw.cin, pw.n, taps)))
        return out
    capi.check(capi.lib().gwd_conv_gemm(ctypes.byref(d), _stream()), "gwd_conv_gemm")
    return out


# ------------------------------------------------------------------------------------------
# attention family
# ------------------------------------------------------------------------------------------
def _L():
    return capi.lib()


def attention(q, k, v, o, *, items, heads, Lq, Lk, hd, q_strides, k_strides, v_strides, o_strides,
              bias=None, mask=None, key_padding=None, scale=1.0, dropout=None):
    """q/k/v/o: bf16 tensors (possibly channel slices via .data_ptr() of a narrowed view); *_strides = (item, row)
    in elements; head h is at +h*hd."""
    d = capi.AttnDesc()
    d.q, d.k, d.v, d.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    d.items, d.heads, d.Lq, d.Lk, d.hd = items, heads, Lq, Lk, hd
    d.q_item_stride, d.q_row_stride = q_strides
    d.k_item_stride, d.k_row_stride = k_strides
    d.v_item_stride, d.v_row_stride = v_strides
    d.o_item_stride, d.o_row_stride = o_strides
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        d.bias = bias.data_ptr()
    if mask is not None:
        assert mask.dtype == torch.float32 and mask.is_contiguous()
        d.mask = mask.data_ptr()
        d.mask_windows = mask.shape[0]
    if key_padding is not None:
        assert key_padding.dtype == torch.uint8 and key_padding.is_contiguous()
        d.key_padding = key_padding.data_ptr()
    d.scale = scale
    if dropout is not None:     # (seed tensor uint32/int32 [1] on the device, site id, p)
        d.dropout_seed, d.dropout_site, d.dropout_p = dropout[0].data_ptr(), int(dropout[1]), float(dropout[2])
    capi.check(_L().gwd_attention(ctypes.byref(d), _stream()), "gwd_attention")
    return o


def token_attention(dq, sq, tk, tv, dout, sout, *, items, N, heads, td, tc, q_rs, k_rs, v_rs, o_rs, scale):
    capi.check(_L().gwd_token_attention(_ptr(dq), _ptr(sq), _ptr(tk), _ptr(tv), _ptr(dout), _ptr(sout), items, N, heads,
                                        td, tc, q_rs, k_rs, v_rs, o_rs, scale, _stream()), "gwd_token_attention")


def ref_scores(q, q_rs, ref_k, ref_rs, out, B, nW, N, heads, hd, R, scale=1.0):
    capi.check(_L().gwd_ref_scores(_ptr(q), q_rs, _ptr(ref_k), ref_rs, _ptr(out), B, nW, N, heads, hd, R, scale,
                                   _stream()), "gwd_ref_scores")


def ref_diffuse(a_in, a_out, w, b, raw_ws, stats_ws, B, heads, P, R):
    capi.check(_L().gwd_ref_diffuse(_ptr(a_in), _ptr(a_out), _ptr(w), _ptr(b), _ptr(raw_ws), _ptr(stats_ws), B, heads, P, R,
                                    _stream()), "gwd_ref_diffuse")


def ref_requery(a, ref_v, ref_rs, out, o_rs, B, nW, N, heads, hd, R, scale):
    capi.check(_L().gwd_ref_requery(_ptr(a), _ptr(ref_v), ref_rs, _ptr(out), o_rs, B, nW, N, heads, hd, R, scale,
                                    _stream()), "gwd_ref_requery")


# ------------------------------------------------------------------------------------------
# bandwidth kernels.  Tensors are bf16 channels-last; `rows` = product of the leading dims.
# ------------------------------------------------------------------------------------------
def _rows(t):
    return t.numel() // t.shape[-1]


def _padded_affine(gamma, beta, C):
    """the row kernels read gamma / beta as 16-byte vectors over the physical width C (a multiple of 8)"""
    if gamma is not None and gamma.numel() < C:
        gamma, beta = pad_vec(gamma, C), pad_vec(beta, C)
    return gamma, beta


def layernorm(x, gamma=None, beta=None, *, res=None, act=ACT_NONE, out=None, n=None, eps=1e-5, C=None, x_coff=0):
    """rows of x[..., x_coff:x_coff+C] -> act(LN(x + res)) (contiguous [rows, C] unless `out` is given)"""
    C = C or x.shape[-1]
    gamma, beta = _padded_affine(gamma, beta, C)
    rows = _rows(x)
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (C,), dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_layernorm(_off(x, x_coff), x.shape[-1], _ptr(res), res.shape[-1] if res is not None else 0,
                                  _ptr(gamma), _ptr(beta), eps, act, _ptr(out), out.shape[-1], rows, C, n or C, _stream()),
               "gwd_layernorm")
    return out


def add_rows(x, addend, period, out=None):
    C = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    capi.check(_L().gwd_add_rows(_ptr(x), C, _ptr(addend), addend.shape[-1], period, _ptr(out), out.shape[-1], _rows(x),
                                 C, _stream()), "gwd_add_rows")
    return out


def _off(t, coff):
    return ctypes.c_void_p(t.data_ptr() + coff * t.element_size())


def window_gather(x, B, H, W, ws, shift, gamma=None, beta=None, n=None, eps=1e-5, C=None, x_coff=0, out=None, y_coff=0):
    """x [B,H,W,Cx] (channels [x_coff, x_coff+C)) -> windows [B*nW*ws*ws, C] of the LayerNorm'ed, zero-padded,
    cyclically shifted map; optionally written into channels [y_coff, y_coff+C) of a wider `out`."""
    C = C or x.shape[-1]
    gamma, beta = _padded_affine(gamma, beta, C)
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    if out is None:
        out = torch.empty(B * Hp * Wp, C, dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_window_gather(_off(x, x_coff), x.shape[-1], _ptr(gamma), _ptr(beta), eps, _off(out, y_coff),
                                      out.shape[-1], B, H, W, ws, shift, C, n or C, _stream()), "gwd_window_gather")
    return out


def window_merge(win, shortcut, B, H, W, ws, shift, gamma=None, beta=None, n=None, eps=1e-5, want_ln=False, C=None,
                 sc_coff=0, win_coff=0):
    """-> (shortcut + unwindowed(win[:, :C]), LN(of that) or None), both [B*H*W, C] contiguous.  win may be wider than C
    (row stride = win.shape[-1]); shortcut may be a channel slice [sc_coff, sc_coff+C) of a wider buffer."""
    C = C or shortcut.shape[-1]
    gamma, beta = _padded_affine(gamma, beta, C)
    rows = B * H * W
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=win.device)
    out_ln = torch.empty(rows, C, dtype=torch.bfloat16, device=win.device) if want_ln else None
    capi.check(_L().gwd_window_merge(_off(win, win_coff), win.shape[-1], _off(shortcut, sc_coff), shortcut.shape[-1], _ptr(out), C,
                                     _ptr(gamma), _ptr(beta), eps, _ptr(out_ln), C, B, H, W, ws, shift, C, n or C,
                                     _stream()), "gwd_window_merge")
    return out, out_ln


def upsample_nearest(x, H, W, add=None, out=None, y_coff=0, C=None, x_coff=0):
    """nearest resize of x[..., x_coff:x_coff+C] ([B,h,w,Cx]) to HxW (+ optional same-size `add`), optionally into a
    channel slice of a wider `out`"""
    B, h, w, Cx = x.shape
    C = C or Cx
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=x.device) if out is None else out
    capi.check(_L().gwd_upsample_nearest(_off(x, x_coff), Cx, B, h, w, _off(out, y_coff), out.shape[-1], H, W, C, _ptr(add),
                                         add.shape[-1] if add is not None else 0, _stream()), "gwd_upsample_nearest")
    return out


def avgpool(x, k, C=None):
    """average-pool the first C channels of x [B,H,W,Cx]"""
    B, H, W, Cx = x.shape
    C = C or Cx
    out = torch.empty(B, H // k, W // k, C, dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_avgpool(_ptr(x), Cx, B, H, W, k, _ptr(out), C, C, _stream()), "gwd_avgpool")
    return out


def avgpool_pyramid(x, C=None):
    """AvgPool2d(k, k) of the first C channels of x [B,H,W,Cx] for k = 16, 8, 4, 2 in one pass -> [p16, p8, p4, p2]"""
    B, H, W, Cx = x.shape
    C = C or Cx
    outs = [torch.empty(B, H // k, W // k, C, dtype=torch.bfloat16, device=x.device) for k in (16, 8, 4, 2)]
    capi.check(_L().gwd_avgpool_pyramid(_ptr(x), Cx, B, H, W, _ptr(outs[3]), _ptr(outs[2]), _ptr(outs[1]), _ptr(outs[0]), C, _stream()),
               "gwd_avgpool_pyramid")
    return outs


def bilinear_up4_into(xs, out, y_coff, H, W):
    """bilinear_up_into of four contiguous maps x_j [B,h_j,w_j,C] into the adjacent slices [y_coff + j*C, y_coff + (j+1)*C) of out"""
    B, _, _, C = xs[0].shape
    assert len(xs) == 4 and all(x.is_contiguous() and x.shape[0] == B and x.shape[3] == C for x in xs)
    hw = (ctypes.c_int32 * 8)(*[v for x in xs for v in x.shape[1:3]])
    dst = ctypes.c_void_p(out.data_ptr() + y_coff * 2)
    capi.check(_L().gwd_bilinear_up4(_ptr(xs[0]), _ptr(xs[1]), _ptr(xs[2]), _ptr(xs[3]), hw, B, dst, out.shape[-1], H, W, C, _stream()),
               "gwd_bilinear_up4")
    return out


def bilinear_up_into(x, out, y_coff, H, W):
    """align_corners=True bilinear resize of x [B,h,w,C] into channels [y_coff, y_coff+C) of out [B,H,W,Ctot]"""
    B, h, w, C = x.shape
    dst = ctypes.c_void_p(out.data_ptr() + y_coff * 2)
    capi.check(_L().gwd_bilinear_up(_ptr(x), C, B, h, w, dst, out.shape[-1], H, W, C, _stream()), "gwd_bilinear_up")
    return out


def _table_bstride(table, B):
    """0 for a table shared by the batch ([H*W, C]), else the per-image stride of a [B, H*W, C] table"""
    if table is None or table.dim() == 2:
        return 0
    assert table.dim() == 3 and table.shape[0] == B and table.is_contiguous()
    return table.shape[1] * table.shape[2]


def sample_bilinear(x, x_coff, table, B, H, W, C, coords, K):
    out = torch.empty(B, K, C, dtype=torch.float32, device=coords.device)
    capi.check(_L().gwd_sample_bilinear(_ptr(x), x.shape[-1] if x is not None else 0, x_coff, _ptr(table), _table_bstride(table, B),
                                        B, H, W, C,
                                        _ptr(coords), K, _ptr(out), _stream()), "gwd_sample_bilinear")
    return out


def sample_scalar(x, coords, K):
    B, H, W = x.shape
    out = torch.empty(B, K, dtype=torch.float32, device=x.device)
    capi.check(_L().gwd_sample_scalar(_ptr(x), B, H, W, _ptr(coords), K, _ptr(out), _stream()), "gwd_sample_scalar")
    return out


def line_ref_gather(win, pos, coords, R, B, H, W, ws, shift, C):
    out = torch.empty(B, R, C, dtype=torch.bfloat16, device=win.device)
    capi.check(_L().gwd_line_ref_gather(_ptr(win), win.shape[-1], _ptr(pos), _table_bstride(pos, B), _ptr(coords), R, _ptr(out), C, B, H, W, ws,
                                        shift, C, _stream()), "gwd_line_ref_gather")
    return out


def anchor_mix(logits, anchor, B, HW, K):
    out = torch.empty(B, HW, dtype=torch.float32, device=logits.device)
    capi.check(_L().gwd_anchor_mix(_ptr(logits), logits.shape[-1], _ptr(anchor), B, HW, K, _ptr(out), _stream()),
               "gwd_anchor_mix")
    return out


def nchw_to_nhwc(x, Cp):
    B, C, H, W = x.shape
    out = torch.empty(B, H, W, Cp, dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_nchw_to_nhwc(_ptr(x), B, C, H * W, _ptr(out), Cp, _stream()), "gwd_nchw_to_nhwc")
    return out


IMAGE_MEAN, IMAGE_STD = (0.538, 0.494, 0.453), (0.257, 0.263, 0.273)      # src/datasets/coco.py:78


# ---------------------------------------------------------------------------------------------- training data path (pixel side)
_PIL_TABLES = {}


def pil_bilinear_tables(in_size, out_size, device):
    """Pillow's BILINEAR resize tables of one axis as device int32 tensors (xmin [out], cnt [out], kk [out, ksize]); cached"""
    key = ("bl", in_size, out_size, str(device))
    t = _PIL_TABLES.get(key)
    if t is None:
        ks = _L().gwd_pil_bilinear_ksize(in_size, out_size)
        xmin, cnt = torch.empty(out_size, dtype=torch.int32), torch.empty(out_size, dtype=torch.int32)
        kk = torch.empty(out_size, ks, dtype=torch.int32)
        capi.check(_L().gwd_pil_bilinear_coeffs(in_size, out_size, _ptr(xmin), _ptr(cnt), _ptr(kk)), "gwd_pil_bilinear_coeffs")
        t = _PIL_TABLES[key] = (xmin.to(device), cnt.to(device), kk.to(device))
        if len(_PIL_TABLES) > 512:
            _PIL_TABLES.pop(next(iter(_PIL_TABLES)))
    return t


def pil_nearest_index(in_size, out_size, device):
    """source index of every output index of a Pillow NEAREST resize, device int32 [out]; cached"""
    key = ("nn", in_size, out_size, str(device))
    t = _PIL_TABLES.get(key)
    if t is None:
        idx = torch.empty(out_size, dtype=torch.int32)
        capi.check(_L().gwd_pil_nearest_index(in_size, out_size, _ptr(idx)), "gwd_pil_nearest_index")
        t = _PIL_TABLES[key] = idx.to(device)
    return t


def resize_bilinear_u8(img, oh, ow, hflip=False, vflip=False):
    """Pillow `resize((ow, oh), BILINEAR)` of a uint8 image [H,W,C] on the device, after an optional horizontal / vertical flip
    (transforms_depth.py:206-263,316-372).  img may be a crop view (rows any stride, pixels contiguous)."""
    H, W, C = img.shape
    assert img.dtype == torch.uint8 and img.is_cuda and img.stride(2) == 1 and img.stride(1) == C
    out, rs = img, img.stride(0)
    if ow != W or hflip:
        xmin, cnt, kk = pil_bilinear_tables(W, ow, img.device)
        tmp = torch.empty(H, ow, C, dtype=torch.uint8, device=img.device)
        capi.check(_L().gwd_resample_u8(_ptr(out), rs, H, W, C, _ptr(tmp), ow, 1, _ptr(xmin), _ptr(cnt), _ptr(kk), kk.shape[1],
                                        int(hflip), _stream()), "gwd_resample_u8")
        out, rs, W = tmp, ow * C, ow
    if oh != H or vflip:
        xmin, cnt, kk = pil_bilinear_tables(H, oh, img.device)
        dst = torch.empty(oh, W, C, dtype=torch.uint8, device=img.device)
        capi.check(_L().gwd_resample_u8(_ptr(out), rs, H, W, C, _ptr(dst), oh, 0, _ptr(xmin), _ptr(cnt), _ptr(kk), kk.shape[1],
                                        int(vflip), _stream()), "gwd_resample_u8")
        out = dst
    return out if out.is_contiguous() else out.contiguous()


def gather2d(mat, oh=None, ow=None, hflip=False, vflip=False, nearest=True):
    """auxiliary map [H,W] (any 1/2/4/8-byte dtype; may be a crop view) -> [oh,ow]: Pillow NEAREST resize after optional flips"""
    H, W = mat.shape
    oh, ow = oh or H, ow or W
    assert mat.is_cuda and mat.stride(1) == 1
    iy = pil_nearest_index(H, oh, mat.device) if oh != H else None
    ix = pil_nearest_index(W, ow, mat.device) if ow != W else None
    out = torch.empty(oh, ow, dtype=mat.dtype, device=mat.device)
    capi.check(_L().gwd_gather2d(_ptr(mat), mat.stride(0), mat.element_size(), H, W, _ptr(out), oh, ow, _ptr(iy), _ptr(ix), int(hflip),
                                 int(vflip), _stream()), "gwd_gather2d")
    return out


JITTER_BRIGHTNESS, JITTER_CONTRAST, JITTER_SATURATION, JITTER_HUE = 0, 1, 2, 3


def color_jitter_u8(img, order, factors):
    """ColorJitter (transforms_depth.py:582-604) in place on a contiguous uint8 [H,W,3] device image: `order` = op ids in the drawn
    order, `factors` = their factors.  One launch, or two around a contrast op (which needs the mean grey level of the image at that
    point)."""
    assert img.dtype == torch.uint8 and img.is_cuda and img.is_contiguous() and img.shape[-1] == 3
    npix = img.numel() // 3
    segments, cur = [], []
    for op, f in zip(order, factors):
        if op == JITTER_CONTRAST and cur:      # a contrast op opens a new launch
            segments.append(cur)
            cur = []
        cur.append((int(op), float(f)))
    if cur:
        segments.append(cur)
    gray = None
    for i, seg in enumerate(segments):
        contrast_first = seg[0][0] == JITTER_CONTRAST
        if contrast_first and i == 0:          # contrast is the very first op: measure the input image
            gray = torch.empty(1, dtype=torch.int64, device=img.device)
            capi.check(_L().gwd_jitter_u8(_ptr(img), npix, 0, None, None, None, _ptr(gray), _stream()), "gwd_jitter_u8")
        gout = torch.empty(1, dtype=torch.int64, device=img.device) if i + 1 < len(segments) else None
        n = len(seg)
        capi.check(_L().gwd_jitter_u8(_ptr(img), npix, n, (ctypes.c_int32 * n)(*[o for o, _ in seg]), (ctypes.c_float * n)(*[f for _, f in seg]),
                                      _ptr(gray) if contrast_first else None, _ptr(gout), _stream()), "gwd_jitter_u8")
        gray = gout
    return img


def images_to_batch(images, mean=IMAGE_MEAN, std=IMAGE_STD, out=None, want_mask=True, table=None):
    """uint8 HWC device images -> (normalised fp32 [B,3,H,W] padded batch, bool [B,H,W] padding mask or None, padded flag).
    images: one uint8 [B,H,W,3] tensor or a list of [h,w,3] tensors (padded bottom / right to the largest)"""
    if isinstance(images, torch.Tensor):
        assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[-1] == 3 and images.is_contiguous()
        B, H, W = images.shape[:3]
        step = H * W * 3
        rows = [[images.data_ptr() + b * step, H, W] for b in range(B)]
        dev, padded = images.device, False
    else:
        assert all(t.dtype == torch.uint8 and t.dim() == 3 and t.shape[-1] == 3 and t.is_contiguous() for t in images)
        B, H, W = len(images), max(t.shape[0] for t in images), max(t.shape[1] for t in images)
        rows = [[t.data_ptr(), t.shape[0], t.shape[1]] for t in images]
        dev, padded = images[0].device, any(t.shape[0] != H or t.shape[1] != W for t in images)
    if table is None:       # callers that reuse their device buffers pass the table of the previous call
        table = torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)
    images_to_batch.last_table = table
    if out is None:
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
    mask = torch.empty(B, H, W, dtype=torch.bool, device=dev) if want_mask else None
    m3, s3 = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    capi.check(_L().gwd_images_to_batch(_ptr(table), B, H, W, m3, s3, _ptr(out), _ptr(mask), _stream()), "gwd_images_to_batch")
    return out, mask, padded


def pack_stem(w, shift):
    """[64,3,7,7] BN-scaled stem filter -> bf16 [64,160] with column ky*22 + kx*3 + c (gwd_stem_conv_pool), fp32 shift"""
    assert tuple(w.shape) == (64, 3, 7, 7)
    packed = torch.zeros(64, 160, dtype=torch.float32, device=w.device)
    wk = w.float().permute(0, 2, 3, 1).reshape(64, 7, 21)        # [n][ky][kx*3 + c]
    packed[:, :154].view(64, 7, 22)[:, :, :21] = wk
    return packed.to(torch.bfloat16).contiguous(), shift.float().contiguous()


def stem_conv_pool(images, w_packed, bias):
    """fp32 [B,3,H,W] -> bf16 [B,PH,PW,64]: conv 7x7/2 + BN shift + ReLU + max-pool 3x3/2 in one launch"""
    assert images.dtype == torch.float32 and images.is_contiguous() and images.shape[1] == 3
    B, _, H, W = images.shape
    ch, cw = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    ph, pw = (ch - 1) // 2 + 1, (cw - 1) // 2 + 1
    out = torch.empty(B, ph, pw, 64, dtype=torch.bfloat16, device=images.device)
    capi.check(_L().gwd_stem_conv_pool(_ptr(images), _ptr(w_packed), _ptr(bias), _ptr(out), B, H, W, _stream()),
               "gwd_stem_conv_pool")
    return out


# ------------------------------------------------------------------------------------------
# selection / reduction kernels
# ------------------------------------------------------------------------------------------
def certain_sample(pred_small, pred_large, K, edges):
    """pred_small [B,h,w], pred_large [B,H,W] fp32 -> (coords fp32 [B,K,1,2], index int32 [B,K])"""
    B, h, w = pred_small.shape
    _, H, W = pred_large.shape
    coords = torch.empty(B, K, 1, 2, dtype=torch.float32, device=pred_large.device)
    index = torch.empty(B, K, dtype=torch.int32, device=pred_large.device)
    e = (ctypes.c_float * len(edges))(*edges)
    capi.check(_L().gwd_certain_sample(_ptr(pred_small), h, w, _ptr(pred_large), H, W, B, K, e, len(edges) - 1,
                                       _ptr(coords), _ptr(index), _stream()), "gwd_certain_sample")
    return coords, index


def match_cost(logits, lines, tgt_lines, tgt_labels, tgt_offsets, w_class, w_line):
    """-> (flat block-diagonal cost fp32 [Q * sum T_b], row_min fp32 [B,Q])"""
    B, Q, ncls = logits.shape
    total_t = tgt_lines.shape[0]
    cost = torch.empty(Q * total_t, dtype=torch.float32, device=logits.device)
    row_min = torch.empty(B, Q, dtype=torch.float32, device=logits.device)
    if total_t == 0:        # no target line in the whole batch: nothing to cost (the reference builds a [B*Q, 0] matrix)
        return cost, row_min.fill_(float("inf"))
    capi.check(_L().gwd_match_cost(_ptr(logits), _ptr(lines), _ptr(tgt_lines), _ptr(tgt_labels), _ptr(tgt_offsets), B, Q,
                                   ncls, lines.shape[-1], w_class, w_line, _ptr(cost), _ptr(row_min), _stream()),
               "gwd_match_cost")
    return cost, row_min


def lsap_batch(cost_flat, offsets, sizes, Q, n_threads=0, raw=False):
    """HOST: cost_flat float32 numpy [sum Q*T_p], offsets / sizes per problem -> list of (query_idx, target_idx) int64 numpy
    pairs, index-for-index what scipy.optimize.linear_sum_assignment returns for every [Q, T_p] block"""
    import numpy as np
    n = len(sizes)
    cost_flat = np.ascontiguousarray(cost_flat, dtype=np.float32)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    T = np.ascontiguousarray(sizes, dtype=np.int32)
    stride = max(Q, 1)
    qi = np.empty((n, stride), dtype=np.int32)
    ti = np.empty((n, stride), dtype=np.int32)
    cnt = np.empty(n, dtype=np.int32)
    capi.check(_L().gwd_lsap_batch(cost_flat.ctypes.data, off.ctypes.data, T.ctypes.data, Q, n, qi.ctypes.data, ti.ctypes.data,
                                   cnt.ctypes.data, n_threads), "gwd_lsap_batch")
    if raw:         # padded int32 arrays [n, max(Q,1)] + counts: for callers that vectorise over the problems
        return qi, ti, cnt
    return [(qi[p, :cnt[p]].astype(np.int64), ti[p, :cnt[p]].astype(np.int64)) for p in range(n)]


def depth_metrics(pred, gt, min_depth=1e-3, max_depth=10.0):
    """pred, gt fp32 [B,H,W] -> fp64 [B,9] (silog, abs_rel, log10, rms, sq_rel, log_rms, d1, d2, d3)"""
    B = pred.shape[0]
    HW = pred[0].numel()
    ws = torch.empty(B, 10, dtype=torch.float64, device=pred.device)
    out = torch.empty(B, 9, dtype=torch.float64, device=pred.device)
    capi.check(_L().gwd_depth_metrics(_ptr(pred), _ptr(gt), B, HW, min_depth, max_depth, _ptr(ws), _ptr(out), _stream()),
               "gwd_depth_metrics")
    return out


def seg_confusion(pred_seg, seg_gt, confusion=None, ignore_index=255):
    """pred_seg fp32 [B,C,H,W] logits (any strides: the forward returns a channels-last view), seg_gt int64 [B,H,W] or
    [B,1,H,W] -> int64 [C,C] confusion[gt][argmax pred], accumulated into `confusion` when given"""
    B, C, H, W = pred_seg.shape
    assert pred_seg.dtype == torch.float32 and pred_seg.stride(0) == C * H * W and pred_seg.stride(2) == W * pred_seg.stride(3)
    gt = seg_gt.reshape(B, H * W).to(torch.int64).contiguous()
    if confusion is None:
        confusion = torch.zeros(C, C, dtype=torch.int64, device=pred_seg.device)
    capi.check(_L().gwd_seg_confusion(_ptr(pred_seg), pred_seg.stride(3), pred_seg.stride(1), pred_seg.stride(0), _ptr(gt), B, H * W, C,
                                      ignore_index, _ptr(confusion), _stream()), "gwd_seg_confusion")
    return confusion


def seg_scores(confusion):
    """compute_mean_ioU's numbers from a confusion matrix (src/util/metrics.py:66-77), fp64 on the device:
    -> (IoU per class * 100, pixel accuracy, mean accuracy, mean IoU)"""
    cm = confusion.double()
    pos, res, tp = cm.sum(1), cm.sum(0), cm.diag()
    iou = tp / torch.clamp(pos + res - tp, min=1.0) * 100
    return iou, tp.sum() / pos.sum() * 100, (tp / torch.clamp(pos, min=1.0)).mean() * 100, iou.mean()


def silog_sums(pred, gt, lo=0.2, hi=10.0, log_only=False):
    """pred fp32 [B,1,h,w], gt fp32 [B,1,H,W] -> fp64 [3] = count, sum d, sum d^2"""
    B, _, h, w = pred.shape
    H, W = gt.shape[-2:]
    sums = torch.empty(3, dtype=torch.float64, device=pred.device)
    capi.check(_L().gwd_silog_sums(_ptr(pred), B, h, w, _ptr(gt), H, W, lo, hi, 1 if log_only else 0, _ptr(sums),
                                   _stream()), "gwd_silog_sums")
    return sums


# ------------------------------------------------------------------------------------------
# training side of the line branch: backward + optimizer kernels (gwd_train.cu)
# ------------------------------------------------------------------------------------------
def _rs(t):
    """row stride of a [..., C] tensor or of a channel slice of a wider buffer"""
    return t.stride(-2) if t.dim() >= 2 and t.shape[-2] > 1 else t.shape[-1]


def layernorm_bwd(dy, z, gamma, dgamma, dbeta, add=None, eps=1e-5, beta=None, post_act=ACT_NONE, n=0):
    """dz (+ add) of y = post_act(LN(z)); dgamma / dbeta (fp32 views, may be None) are accumulated.  dy / add may be channel
    slices of wider buffers; n: logical channels when z carries zero padding up to C"""
    C = z.shape[-1]
    rows = _rows(z)
    assert z.is_contiguous() and dy.stride(-1) == 1
    dz = torch.empty(rows, C, dtype=torch.bfloat16, device=z.device)
    capi.check(_L().gwd_layernorm_bwd(_ptr(dy), _rs(dy), _ptr(z), C, _ptr(gamma), _ptr(beta), post_act, eps, _ptr(add),
                                      _rs(add) if add is not None else 0, _ptr(dz), C, _ptr(dgamma), _ptr(dbeta), rows, C, n,
                                      _stream()), "gwd_layernorm_bwd")
    return dz


def bilinear_up_bwd(dy, h, w, C=None):
    """backward of bilinear_up_into: dy bf16 [B,H,W,C] (may be a channel slice of the concat gradient) -> [B,h,w,C]"""
    B, H, W, Cs = dy.shape
    C = C or Cs
    dx = torch.empty(B, h, w, C, dtype=torch.bfloat16, device=dy.device)
    capi.check(_L().gwd_bilinear_up_bwd(_ptr(dy), dy.stride(2), B, H, W, _ptr(dx), C, h, w, C, _stream()), "gwd_bilinear_up_bwd")
    return dx


def avgpool_bwd(d, k, H, W, add=None, out=None, scale=1.0):
    """backward of avgpool(k): d bf16 [B,H//k,W//k,C] -> out [B,H,W,C] = add + d spread over the k x k cells * scale / k^2
    (out may be `add` itself or a channel slice)"""
    B, oh, ow, C = d.shape
    assert (oh, ow) == (H // k, W // k) and d.is_contiguous()
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=d.device) if out is None else out
    capi.check(_L().gwd_avgpool_bwd(_ptr(d), C, k, scale, _ptr(add), add.stride(2) if add is not None else 0, _ptr(out),
                                    out.stride(2), B, H, W, C, _stream()), "gwd_avgpool_bwd")
    return out


def act_bwd(dy, y, act, out_cols=None, y_mul=1.0, scale=1.0, from_input=False):
    """bf16 [rows, out_cols] = dy * act'(.) * scale, act' from the output y * y_mul (or from the activation's input with
    from_input=True); zeros in the padding columns"""
    n = dy.shape[-1]
    rows = _rows(dy)
    out_cols = out_cols or round_up(n, 8)
    out = torch.empty(rows, out_cols, dtype=torch.bfloat16, device=dy.device)
    capi.check(_L().gwd_act_bwd(_ptr(dy), int(dy.dtype == torch.float32), n, _ptr(y),
                                int(y is not None and y.dtype == torch.float32), y.shape[-1] if y is not None else 0, act,
                                _ptr(out), out_cols, rows, n, out_cols, y_mul, scale, int(from_input), _stream()), "gwd_act_bwd")
    return out


def silog_bwd(pred, gt, sums, *, weight=1.0, lo=0.2, hi=10.0, log_only=False, variance_focus=0.85, sig_scale=0.0,
              out_cols=0, loss_out=None):
    """gradient of weight * SilogLoss to pred fp32 [B,1,h,w] (or, with sig_scale > 0, to the pre-sigmoid value) from the
    device-resident sums of `silog_sums`: fp32 [B*h*w] (out_cols 0) or bf16 [B*h*w, out_cols] with the gradient in column 0"""
    B, _, h, w = pred.shape
    H, W = gt.shape[-2:]
    n = B * h * w
    out = (torch.empty(n, dtype=torch.float32, device=pred.device) if out_cols == 0
           else torch.empty(n, out_cols, dtype=torch.bfloat16, device=pred.device))
    capi.check(_L().gwd_silog_bwd(_ptr(pred), B, h, w, _ptr(gt), H, W, lo, hi, 1 if log_only else 0, _ptr(sums), variance_focus,
                                  weight, sig_scale, _ptr(out), out_cols, _ptr(loss_out), _stream()), "gwd_silog_bwd")
    return out


def seg_ce(logits, seg_gt, *, weight=1.0, ignore_index=-100, out_cols=16, want_grad=True, loss_out=None, sums=None):
    """weight * mean cross entropy of fp32 logits [B,C,H,W] (any strides: NCHW or a permuted channels-last buffer) against
    int64 seg_gt [B,(1,)H,W]; returns (sums fp64 [2] = (n, sum nll), dlogits bf16 [B*H*W, out_cols] channels-last or None)"""
    assert logits.dtype == torch.float32 and logits.dim() == 4 and seg_gt.dtype == torch.int64 and seg_gt.is_contiguous()
    B, C, H, W = logits.shape
    assert logits.stride(2) == W * logits.stride(3), "rows of the logits map must be contiguous in pixels"
    sums = torch.empty(2, dtype=torch.float64, device=logits.device) if sums is None else sums
    d = torch.empty(B * H * W, out_cols, dtype=torch.bfloat16, device=logits.device) if want_grad else None
    capi.check(_L().gwd_seg_ce(_ptr(logits), logits.stride(3), logits.stride(1), logits.stride(0), _ptr(seg_gt), B, H * W, C,
                               ignore_index, weight, _ptr(sums), _ptr(d), out_cols, _ptr(loss_out), _stream()), "gwd_seg_ce")
    return sums, d


def transpose(x, colsum=None, C=None, x_coff=0, pad_to=64, out=None):
    """bf16 [rows, Cx] (columns [x_coff, x_coff+C)) -> [C, round_up(rows, pad_to)] with zero padding columns;
    colsum (fp32 [C] view) accumulates the column sums of x"""
    rows, Cx = _rows(x), x.shape[-1]
    C = C or Cx
    rp = round_up(rows, pad_to)
    if out is None:
        out = torch.empty(C, rp, dtype=torch.bfloat16, device=x.device)
    assert out.shape == (C, rp) and out.is_contiguous()
    capi.check(_L().gwd_transpose(_off(x, x_coff), Cx, _ptr(out), rp, rows, rp, C, _ptr(colsum), _stream()), "gwd_transpose")
    return out


def transpose_batch_tables(pairs):
    """pairs: list of (src [rows, cols] bf16 view with unit column stride, dst [cols, rows] bf16 contiguous) ->
    device tables for transpose_batch"""
    tab, prefix = [], [0]
    for src, dst in pairs:
        rows, cols = src.shape
        assert src.dtype == dst.dtype == torch.bfloat16 and src.stride(1) == 1 and dst.shape == (cols, rows) and dst.is_contiguous()
        assert rows % 2 == 0 and cols % 2 == 0
        tab.append([src.data_ptr(), dst.data_ptr(), rows, cols, src.stride(0), dst.stride(0)])
        prefix.append(prefix[-1] + ((rows + 63) // 64) * ((cols + 63) // 64))
    dev = pairs[0][0].device
    return (torch.tensor(tab, dtype=torch.int64, device=dev), torch.tensor(prefix, dtype=torch.int32, device=dev), len(pairs), prefix[-1])


def anchor_mix_bwd(logits, anchor, dpred, B, HW, K):
    """backward of anchor_mix: -> (dlogits bf16 [B*HW, Kp] with zero padding columns, danchor fp32 [B, K])"""
    Kp = logits.shape[-1]
    dlogits = torch.empty(B * HW, Kp, dtype=torch.bfloat16, device=logits.device)
    danchor = torch.zeros(B, K, dtype=torch.float32, device=logits.device)
    assert dpred.dtype == torch.float32 and dpred.is_contiguous() and dpred.numel() == B * HW
    capi.check(_L().gwd_anchor_mix_bwd(_ptr(logits), Kp, _ptr(anchor), _ptr(dpred), B, HW, K, Kp, _ptr(dlogits), Kp, _ptr(danchor),
                                       _stream()), "gwd_anchor_mix_bwd")
    return dlogits, danchor


def sample_bilinear_bwd(d, coords, out, y_coff, H, W):
    """backward of sample_bilinear w.r.t. the map: d fp32 [B,K,C] -> out[..., y_coff:y_coff+C] (bf16 [B*H*W, Cout] buffer)"""
    B, K, C = d.shape
    assert d.dtype == torch.float32 and d.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    capi.check(_L().gwd_sample_bilinear_bwd(_ptr(d), _ptr(coords), K, _off(out, y_coff), out.shape[-1], B, H, W, C, _stream()),
               "gwd_sample_bilinear_bwd")
    return out


def sample_scalar_bwd(d, coords, H, W, add=None):
    """backward of sample_scalar: d fp32 [B,K] -> fp32 [B,H,W] (+ add)"""
    B, K = d.shape
    out = torch.empty(B, H, W, dtype=torch.float32, device=d.device)
    assert add is None or (add.dtype == torch.float32 and add.is_contiguous() and add.numel() == out.numel())
    capi.check(_L().gwd_sample_scalar_bwd(_ptr(d), _ptr(coords), K, _ptr(add), _ptr(out), B, H, W, _stream()),
               "gwd_sample_scalar_bwd")
    return out


def window_attention_bwd(qkv, d_o, *, items, heads, N, hd, scale, bias=None, mask=None, dbias=None):
    """backward of attention() on fused window projections: qkv bf16 [items*N, 3C], d_o bf16 [items*N, C] -> dqkv [items*N, 3C];
    dbias fp32 [heads, N, N] (optional) is accumulated"""
    C = heads * hd
    assert qkv.dtype == d_o.dtype == torch.bfloat16 and qkv.is_contiguous() and d_o.is_contiguous()
    assert qkv.shape == (items * N, 3 * C) and d_o.shape == (items * N, C)
    dqkv = torch.empty_like(qkv)
    capi.check(_L().gwd_window_attention_bwd(_ptr(qkv), 3 * C, _ptr(d_o), C, _ptr(dqkv), 3 * C, _ptr(bias), _ptr(mask),
                                             mask.shape[0] if mask is not None else 0, _ptr(dbias), items, heads, N, hd, scale,
                                             _stream()), "gwd_window_attention_bwd")
    return dqkv


def token_attention_bwd(dq, sq, gkv, d_dout, d_sout, *, items, N, heads, td, tc, scale):
    """backward of token_attention(dq, sq, gkv[:, :tC], gkv[:, tC:]): -> (g_dq, g_sq [rows, heads*td], g_gkv [rows, 2*heads*tc])"""
    tC = heads * tc
    assert gkv.shape[-1] == 2 * tC and all(t.is_contiguous() for t in (dq, sq, gkv, d_dout, d_sout))
    g_dq, g_sq, g_gkv = torch.empty_like(dq), torch.empty_like(sq), torch.empty_like(gkv)
    capi.check(_L().gwd_token_attention_bwd(_ptr(dq), _ptr(sq), _ptr(gkv), _off(gkv, tC), _ptr(d_dout), _ptr(d_sout), _ptr(g_dq),
                                            _ptr(g_sq), _ptr(g_gkv), _off(g_gkv, tC), items, N, heads, td, tc, dq.shape[-1],
                                            2 * tC, 2 * tC, d_dout.shape[-1], g_dq.shape[-1], 2 * tC, 2 * tC, scale, _stream()),
               "gwd_token_attention_bwd")
    return g_dq, g_sq, g_gkv


def transpose_batch(tables):
    table, prefix, n, total = tables
    capi.check(_L().gwd_transpose_batch(_ptr(table), _ptr(prefix), n, total, _stream()), "gwd_transpose_batch")


def linear_wgrad(dy, x, dw, db=None, N=None, K=None, x_coff=0):
    """dw [N, K] (fp32 view) += dy[:, :N]^T x[:, x_coff:x_coff+K]; db [N] += column sums of dy.  dy, x: bf16 [rows, *]"""
    N = N or dw.shape[0]
    K = K or dw.shape[1]
    assert dy.dtype == x.dtype == torch.bfloat16 and dw.dtype == torch.float32 and dw.stride(-1) == 1
    capi.check(_L().gwd_linear_wgrad(_ptr(dy), dy.shape[-1], _off(x, x_coff), x.shape[-1], _rows(dy), N, K, _ptr(dw), dw.stride(0),
                                     _ptr(db), _stream()), "gwd_linear_wgrad")


def conv3x3_wgrad(dy, x, dw, db=None, N=None, C=None):
    """dw fp32 [9, N, C] (tap = dx*3+dy, the packed layout of pack_conv3x3) += 3x3-conv weight gradient; dy, x: bf16
    channels-last [B,H,W,*]"""
    B, H, W = x.shape[:3]
    N, C = N or dw.shape[1], C or dw.shape[2]
    assert dy.dtype == x.dtype == torch.bfloat16 and dw.dtype == torch.float32 and dw.is_contiguous() and dw.shape == (9, N, C)
    capi.check(_L().gwd_conv3x3_wgrad(_ptr(dy), dy.shape[-1], _ptr(x), x.shape[-1], B, H, W, N, C, _ptr(dw), _ptr(db), _stream()),
               "gwd_conv3x3_wgrad")


def pack_conv3x3_dgrad(weight, cin_pad=None):
    """the filter whose gwd_conv_gemm on dY gives dX of a stride-1, pad-1 3x3 conv with `weight` [N, C, 3, 3]:
    transposed in (n, c), flipped in (dy, dx)"""
    return pack_conv3x3(weight.transpose(0, 1).flip(2, 3).contiguous(), None, cin_pad=cin_pad)


def unpack_conv3x3_grad(dw, n, c):
    """packed gradient [9, N_pad, C_pad] (tap = dx*3+dy) -> the reference's [n, c, 3(dy), 3(dx)] layout"""
    return dw.view(3, 3, dw.shape[1], dw.shape[2])[:, :, :n, :c].permute(2, 3, 1, 0).contiguous()


def attention_bwd(q, k, v, d_o, dq, dk, dv, *, items, heads, Lq, Lk, hd, q_strides, k_strides, v_strides, do_strides,
                  dq_strides, dk_strides, dv_strides, scale=1.0, o=None, o_strides=None, dq_mul=0.0, dk_mul=0.0, dropout=None,
                  key_padding=None):
    """o: the forward output (bf16) -> tensor-core kernel; None -> CUDA-core kernel that recomputes D = rowsum(P dP).
    key_padding: uint8 / bool [items, Lk], 1 = padded key.  Lq, Lk <= 1280 (beyond 512: two launches + a scratch, needs o)"""
    d = capi.AttnBwdDesc()
    d.q, d.k, d.v, d.d_o = q.data_ptr(), k.data_ptr(), v.data_ptr(), d_o.data_ptr()
    d.dq, d.dk, d.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    d.items, d.heads, d.Lq, d.Lk, d.hd = items, heads, Lq, Lk, hd
    d.q_item_stride, d.q_row_stride = q_strides
    d.k_item_stride, d.k_row_stride = k_strides
    d.v_item_stride, d.v_row_stride = v_strides
    d.do_item_stride, d.do_row_stride = do_strides
    d.dq_item_stride, d.dq_row_stride = dq_strides
    d.dk_item_stride, d.dk_row_stride = dk_strides
    d.dv_item_stride, d.dv_row_stride = dv_strides
    d.scale, d.dq_mul, d.dk_mul = scale, dq_mul, dk_mul
    if o is not None:
        d.o = o.data_ptr()
        d.o_item_stride, d.o_row_stride = o_strides or do_strides
    if dropout is not None:
        d.dropout_seed, d.dropout_site, d.dropout_p = dropout[0].data_ptr(), int(dropout[1]), float(dropout[2])
    if key_padding is not None:
        kp = key_padding.to(torch.uint8) if key_padding.dtype != torch.uint8 else key_padding
        assert kp.is_contiguous() and kp.numel() == items * Lk
        d.key_padding = kp.data_ptr()
    ws = None
    if Lq > 512 or Lk > 512:
        ws = torch.empty(items * heads * Lq * 2, dtype=torch.float32, device=q.device)
        d.stats_ws = ws.data_ptr()
    capi.check(_L().gwd_attention_bwd(ctypes.byref(d), _stream()), "gwd_attention_bwd")


def set_loss(logits, lines, tgt_lines, tgt_labels, match, stage_off, class_w, w_ce, w_line, num_items):
    """-> (losses fp32 [S,2], dlogits, dlines) of the weighted set criterion; see gwd_set_loss in include/gwd_b200.h"""
    S, B, Q, C = logits.shape
    D = lines.shape[-1]
    assert logits.dtype == lines.dtype == torch.float32 and logits.is_contiguous() and lines.is_contiguous()
    assert match.dtype == torch.int32 and match.is_contiguous() and match.shape[0] == 4
    losses = torch.empty(S, 2, dtype=torch.float32, device=logits.device)
    dlogits, dlines = torch.empty_like(logits), torch.empty_like(lines)
    M = match.shape[1]
    if M == 0:              # no matched pair at all: the kernel still needs non-null (unread) pointers
        match = torch.zeros(4, 1, dtype=torch.int32, device=logits.device)
    if tgt_lines.numel() == 0:
        tgt_lines = torch.zeros(1, D, dtype=torch.float32, device=logits.device)
        tgt_labels = torch.zeros(1, dtype=torch.int64, device=logits.device)
    capi.check(_L().gwd_set_loss(_ptr(logits), _ptr(lines), _ptr(tgt_lines), _ptr(tgt_labels), _ptr(match), _ptr(stage_off),
                                 _ptr(class_w), _ptr(w_ce), _ptr(w_line), _ptr(num_items), S, B, Q, C, D, M, _ptr(losses),
                                 _ptr(dlogits), _ptr(dlines), _stream()), "gwd_set_loss")
    return losses, dlogits, dlines


def sumsq(g, out):
    """out (fp64 [1], zeroed by the caller) += sum g^2"""
    capi.check(_L().gwd_sumsq(_ptr(g), g.numel(), _ptr(out), _stream()), "gwd_sumsq")


def adamw_step(p, g, m, v, mirror, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step=1, max_norm=0.0,
               grad_scale=1.0, sumsq_buf=None):
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
    capi.check(_L().gwd_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(mirror), p.numel(), lr, betas[0], betas[1], eps,
                                   weight_decay, step, max_norm, grad_scale, _ptr(sumsq_buf), _stream()), "gwd_adamw_step")


# ------------------------------------------------------------------------------------------
# training side of the 1/32 line-window stage (gwd_train_line.cu) and backbone-backward helpers
# ------------------------------------------------------------------------------------------
def diffuse_filter_pack(w_phys, bias, fwd=None, bwd=None):
    """w_phys fp32 [9,16,16] (tap = kx*3+ky: the flat-buffer layout of the 16x16x3x3 filter), bias fp32 [>=16] -> (fwd, bwd)
    device filters (2 320 floats) for ref_diffuse_dev / ref_diffuse_bwd"""
    fwd = torch.empty(2320, dtype=torch.float32, device=w_phys.device) if fwd is None else fwd
    bwd = torch.empty(2320, dtype=torch.float32, device=w_phys.device) if bwd is None else bwd
    capi.check(_L().gwd_diffuse_filter_pack(_ptr(w_phys), _ptr(bias), _ptr(fwd), _ptr(bwd), _stream()), "gwd_diffuse_filter_pack")
    return fwd, bwd


def ref_diffuse_dev(a_in, filt, B, heads, P, R):
    """one diffusion round with a device filter -> (a_out, raw, stats) (all kept for the backward)"""
    a_out, raw = torch.empty_like(a_in), torch.empty_like(a_in)
    stats = torch.empty(B * heads * 2, dtype=torch.float64, device=a_in.device)
    capi.check(_L().gwd_ref_diffuse_dev(_ptr(a_in), _ptr(a_out), _ptr(filt), _ptr(raw), _ptr(stats), B, heads, P, R, _stream()),
               "gwd_ref_diffuse_dev")
    return a_out, raw, stats


def ref_diffuse_bwd(g, raw, stats, a_in, filt_bwd, dw_phys, db, B, heads, P, R, ws=None):
    """-> d a_in; accumulates dw_phys [9,16,16] / db [16].  ws: optional (d_raw, stats2) workspaces"""
    d_raw = torch.empty_like(g) if ws is None else ws[0]
    stats2 = torch.empty(4 * B * heads, dtype=torch.float64, device=g.device) if ws is None else ws[1]
    d_in = torch.empty_like(g)
    capi.check(_L().gwd_ref_diffuse_bwd(_ptr(g), _ptr(raw), _ptr(stats), _ptr(a_in), _ptr(filt_bwd), _ptr(d_raw), _ptr(stats2),
                                        _ptr(d_in), _ptr(dw_phys), _ptr(db), B, heads, P, R, _stream()), "gwd_ref_diffuse_bwd")
    return d_in


def ref_affine(ref, mu, logsigma, D):
    """ref fp32 [rows, >= D] -> ref_k fp32 [rows, D] = mu + exp(logsigma) * ref[:, :D]"""
    rows = ref.shape[0]
    out = torch.empty(rows, D, dtype=torch.float32, device=ref.device)
    capi.check(_L().gwd_ref_affine(_ptr(ref), ref.shape[-1], _ptr(mu), _ptr(logsigma), _ptr(out), rows, D, _stream()), "gwd_ref_affine")
    return out


def ref_affine_bwd(d_kv, ref, logsigma, dmu, dlogsigma, D):
    """d_kv fp32 [rows, 2D] -> d_ref bf16 [rows, 2D]; dmu / dlogsigma (fp32 [D] views) accumulated"""
    rows = d_kv.shape[0]
    assert d_kv.shape[1] == 2 * D and d_kv.is_contiguous() and ref.is_contiguous()
    d_ref = torch.empty(rows, 2 * D, dtype=torch.bfloat16, device=d_kv.device)
    capi.check(_L().gwd_ref_affine_bwd(_ptr(d_kv), _ptr(ref), ref.shape[-1], _ptr(logsigma), _ptr(d_ref), _ptr(dmu), _ptr(dlogsigma),
                                       rows, D, _stream()), "gwd_ref_affine_bwd")
    return d_ref


def ref_requery_bwd(a, refv, ref_rs, d_qnew, dq_rs, d_refv, drv_rs, B, T, heads, hd, R, scale):
    """-> d a (soft-max backward included); d_refv (a channel slice of an fp32 [B*R, *] buffer) is written"""
    d_a = torch.empty_like(a)
    capi.check(_L().gwd_ref_requery_bwd(_ptr(a), _ptr(refv), ref_rs, _ptr(d_qnew), dq_rs, _ptr(d_a), _ptr(d_refv), drv_rs, B, T, heads,
                                        hd, R, scale, _stream()), "gwd_ref_requery_bwd")
    return d_a


def ref_scores_bwd(d_a, refk, ref_rs, q, q_rs, d_q, dq_rs, d_refk, drk_rs, B, T, heads, hd, R, scale):
    capi.check(_L().gwd_ref_scores_bwd(_ptr(d_a), _ptr(refk), ref_rs, _ptr(q), q_rs, _ptr(d_q), dq_rs, _ptr(d_refk), drk_rs, B, T, heads,
                                       hd, R, scale, _stream()), "gwd_ref_scores_bwd")


def line_ref_scatter(d_ref, coords, R, d_win, B, H, W, ws, shift, C):
    """d_win[row of point (b, r)] += d_ref[b, r]: the adjoint of the feature part of line_ref_gather (in place)"""
    capi.check(_L().gwd_line_ref_scatter(_ptr(d_ref), d_ref.shape[-1], _ptr(coords), R, _ptr(d_win), d_win.shape[-1], B, H, W, ws, shift,
                                         C, _stream()), "gwd_line_ref_scatter")
    return d_win


def subsample2(x):
    """bf16 [B,H,W,C] -> x[:, ::2, ::2, :] contiguous"""
    B, H, W, C = x.shape
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    y = torch.empty(B, (H + 1) // 2, (W + 1) // 2, C, dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_subsample2(_ptr(x), _ptr(y), B, H, W, C, _stream()), "gwd_subsample2")
    return y


def zero_stuff2(s, H, W, add=None):
    """bf16 [B,h,w,C] -> [B,H,W,C] with s at the even pixels (zeros elsewhere) (+ add)"""
    B, h, w, C = s.shape
    assert (h, w) == ((H + 1) // 2, (W + 1) // 2) and s.is_contiguous() and (add is None or (add.is_contiguous() and add.numel() == B * H * W * C))
    y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=s.device)
    capi.check(_L().gwd_zero_stuff2(_ptr(s), _ptr(add), _ptr(y), B, H, W, C, _stream()), "gwd_zero_stuff2")
    return y


def scale_rows(g, scale):
    capi.check(_L().gwd_scale_rows(_ptr(g), _ptr(scale), g.numel(), _stream()), "gwd_scale_rows")


def fold_mirror(p, scale, mirror):
    capi.check(_L().gwd_fold_mirror(_ptr(p), _ptr(scale), _ptr(mirror), p.numel(), _stream()), "gwd_fold_mirror")


def im2col3x3_s2(x):
    """bf16 [B,H,W,C] -> [B,ho,wo,9C]: the patches of a stride-2, padding-1 3x3 convolution, K order (ky, kx, c)"""
    B, H, W, C = x.shape
    assert x.is_contiguous() and x.dtype == torch.bfloat16
    col = torch.empty(B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 9 * C, dtype=torch.bfloat16, device=x.device)
    capi.check(_L().gwd_im2col3x3_s2(_ptr(x), _ptr(col), B, H, W, C, _stream()), "gwd_im2col3x3_s2")
    return col


def col2im3x3_s2(dcol, H, W, add=None):
    """adjoint of im2col3x3_s2: bf16 [B,ho,wo,9C] -> [B,H,W,C] (+ add)"""
    B, ho, wo, C9 = dcol.shape
    C = C9 // 9
    assert dcol.is_contiguous() and (ho, wo) == ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
    assert add is None or (add.is_contiguous() and add.numel() == B * H * W * C)
    dx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dcol.device)
    capi.check(_L().gwd_col2im3x3_s2(_ptr(dcol), _ptr(add), _ptr(dx), B, H, W, C, _stream()), "gwd_col2im3x3_s2")
    return dx


def select_lines(logits, lines, num_ref, points_per_line):
    """logits fp32 [B,Q,C], lines fp32 [B,Q,D] -> (ref_xy fp32 [B, num_ref * points_per_line, 2] in [-1,1], ids int64 [B,num_ref]):
    the top-num_ref lines by raw line logit and their points (multiscale_transformerr.py:1165-1179)"""
    B, Q, C = logits.shape
    assert logits.dtype == lines.dtype == torch.float32 and logits.is_contiguous() and lines.is_contiguous()
    ref_xy = torch.empty(B, num_ref * points_per_line, 2, dtype=torch.float32, device=logits.device)
    ids = torch.empty(B, num_ref, dtype=torch.int64, device=logits.device)
    capi.check(_L().gwd_select_lines(_ptr(logits), C, _ptr(lines), lines.shape[-1], B, Q, num_ref, points_per_line, _ptr(ref_xy),
                                     _ptr(ids), _stream()), "gwd_select_lines")
    return ref_xy, ids


def dropout(x, seed, site, p, res=None, out=None):
    """bf16 element-wise dropout: out = res + (keep ? x / (1 - p) : 0); seed: int32 [1] device tensor, site: call-site id.
    Applied to a gradient with the same (seed, site) it is its own backward (the mask is regenerated, never stored)."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0
    assert res is None or (res.dtype == torch.bfloat16 and res.is_contiguous() and res.numel() == x.numel())
    out = torch.empty_like(x) if out is None else out
    capi.check(_L().gwd_dropout(_ptr(x), _ptr(res), _ptr(out), x.numel(), _ptr(seed), int(site), float(p), _stream()), "gwd_dropout")
    return out
