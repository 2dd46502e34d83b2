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
              res=None, res_coff=0, res_mode=RES_NONE, y_raw=None):
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
    if taps == 9:
        assert x.dim() == 4
    store_n = round_up(pw.n, 8) if not out_f32 else pw.n
    if out is None:
        oc = out_channels or store_n
        out = torch.empty(tuple(x.shape[:-1]) + (oc,), dtype=torch.float32 if out_f32 else torch.bfloat16,
                          device=x.device)
    assert out.is_contiguous()
    d = capi.GemmDesc()
    d.x = x.data_ptr(); d.B, d.H, d.W = B, H, W
    d.x_cstride, d.x_coff, d.cin = Cx, x_coff, pw.cin_pad
    d.w = pw.w.data_ptr(); d.taps = taps; d.n_pad = pw.n_pad; d.n = pw.n
    d.bias = pw.bias.data_ptr() if (bias and pw.bias is not None) else None
    if ln is not None:
        d.ln_g, d.ln_b = ln[0].data_ptr(), ln[1].data_ptr()
    d.ln_eps = ln_eps
    d.pre_act, d.post_act, d.out_scale = pre_act, post_act, out_scale
    if res is not None:
        assert res.dtype == torch.bfloat16 and res.is_contiguous() and res_mode != RES_NONE
        d.res = res.data_ptr(); d.res_cstride = res.shape[-1]; d.res_coff = res_coff; d.res_mode = res_mode
    d.y = out.data_ptr(); d.y_cstride = out.shape[-1]; d.y_coff = y_coff; d.y_f32 = 1 if out_f32 else 0
    if y_raw is not None:
        d.y_raw = y_raw.data_ptr(); d.yraw_cstride = y_raw.shape[-1]; d.yraw_coff = 0
    d.store_n = store_n
    capi.check(capi.lib().gwd_conv_gemm(ctypes.byref(d), _stream()), "gwd_conv_gemm")
    return out
