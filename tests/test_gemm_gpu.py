"""GPU parity of gwd_conv_gemm (tcgen05/TMA implicit GEMM) against fp32 torch on bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def _rand(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale)


def _act(v, a):
    return {0: lambda t: t, 1: F.relu, 2: F.gelu, 3: F.elu, 4: torch.sigmoid}[a](v)


def _ref_epilogue(acc, bias, pre_act, res, res_mode, ln, post_act, scale, n):
    v = acc + (bias if bias is not None else 0)
    v = _act(v, pre_act)
    if res_mode == 1:
        v = v + res
    if ln is not None:
        v = F.layer_norm(v, (n,), ln[0], ln[1], 1e-5)
    v = _act(v, post_act) * scale
    if res_mode == 2:
        v = v + res
    return v


LINEAR_CASES = [
    # M, K, N, pre, post, res_mode, ln
    (300, 256, 256, 0, 1, 0, False),
    (1000, 256, 2048, 0, 1, 0, False),
    (777, 2048, 256, 0, 0, 1, True),
    (4800, 64, 192, 0, 2, 0, False),
    (130, 512, 1536, 0, 0, 0, False),
    (19200, 192, 64, 0, 0, 2, False),
    (513, 80, 160, 0, 2, 0, True),
    (100, 256, 6, 0, 4, 0, False),
    # staged epilogue (residual tile fetched / result stored through per-warp shared-memory tiles)
    (5000, 64, 256, 0, 1, 1, False),      # ResNet conv3: relu(x W + res), ragged last tile
    (2500, 256, 1024, 0, 1, 1, False),    # four N tiles
    (1111, 512, 128, 0, 0, 2, False),     # MLP fc2 + residual after
    (3000, 128, 64, 0, 2, 2, False),      # 16-column staging rows
    (700, 64, 120, 0, 1, 1, False),       # logical width below the padded width
    # TMA epilogue (>= 296 M tiles, 64 columns per epilogue warp): residual tile in / result out through shared memory
    (40001, 64, 256, 0, 1, 1, False),     # ResNet conv3 shape, ragged last tile
    (38013, 256, 1024, 0, 0, 2, False),   # four N tiles, residual after
    (39000, 128, 250, 0, 1, 0, False),    # logical width below the padded width, no residual
    (20011, 64, 192, 0, 2, 0, False),     # 12 epilogue warps (3 groups of 64 columns), GELU
    (19999, 192, 384, 0, 0, 1, False),    # two N tiles of 192
    (21000, 64, 128, 0, 2, 2, False),     # 8 epilogue warps
    # large M with streamed weights
    (76057, 256, 512, 0, 1, 0, False),
    (75800, 512, 320, 0, 0, 2, False),
]


@pytest.mark.parametrize("M,K,N,pre,post,res_mode,use_ln", LINEAR_CASES)
def test_linear(M, K, N, pre, post, res_mode, use_ln):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + K + N)
    x = _rand((M, K), g).bfloat16()
    w = _rand((N, K), g, K ** -0.5).bfloat16()
    b = _rand((N,), g, 0.5)
    n8 = ops.round_up(N, 16)
    res = _rand((M, n8), g).bfloat16() if res_mode else None
    ln = (1 + 0.1 * _rand((N,), g), 0.1 * _rand((N,), g)) if use_ln else None
    acc = x.float() @ w.float().t()
    ref = _ref_epilogue(acc, b, pre, res.float()[:, :N] if res is not None else None, res_mode, ln, post, 1.0, N)

    pw = ops.pack_linear(w.cuda(), b.cuda())
    lnp = (ops.pad_vec(ln[0].cuda(), pw.n_pad), ops.pad_vec(ln[1].cuda(), pw.n_pad)) if use_ln else None
    y = ops.conv_gemm(x.cuda(), pw, ln=lnp, pre_act=pre, post_act=post,
                      res=res.cuda() if res is not None else None, res_mode=res_mode)
    torch.cuda.synchronize()
    got = y.float().cpu()[:, :N]
    err = (got - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    assert err < tol, (err, tol)
    if n8 > N:
        assert (y.float().cpu()[:, N:] == 0).all()


CONV_CASES = [
    # B, H, W, C, N, pre, post, res_mode, ln
    (2, 30, 40, 64, 64, 0, 3, 0, False),
    (2, 18, 22, 64, 64, 3, 0, 0, True),
    (1, 24, 32, 160, 160, 0, 2, 0, True),
    (2, 17, 23, 80, 80, 0, 2, 0, True),
    (1, 16, 24, 160, 160, 0, 0, 2, True),
    (1, 20, 28, 32, 32, 0, 3, 0, False),
    (1, 12, 16, 320, 320, 0, 0, 0, False),
    (1, 9, 13, 1024, 256, 0, 2, 0, False),
    (3, 7, 10, 160, 160, 0, 2, 0, True),
    # >= 592 M tiles, filter too large to stay resident
    (4, 120, 160, 160, 160, 0, 2, 0, True),
    (7, 97, 100, 80, 160, 0, 0, 0, False),    # 637 M tiles (odd), ragged borders
    # CTA-pair path (cta_group::2: >= 296 M tiles, streamed weights): LayerNorm + residual, two N tiles, an odd tile count whose
    # last pair has a ghost tile, the 64-byte-swizzle K chunk (C = 32 * odd)
    (4, 120, 160, 160, 160, 0, 0, 2, True),
    (8, 64, 80, 64, 320, 0, 3, 0, False),
    (3, 125, 103, 96, 192, 0, 2, 1, False),
]


@pytest.mark.parametrize("B,H,W,C,N,pre,post,res_mode,use_ln", CONV_CASES)
def test_conv3x3(B, H, W, C, N, pre, post, res_mode, use_ln):
    ops = _ops()
    g = torch.Generator().manual_seed(B + H * 3 + W * 5 + C + N)
    x = _rand((B, H, W, C), g).bfloat16()
    w = _rand((N, C, 3, 3), g, (9 * C) ** -0.5).bfloat16()
    b = _rand((N,), g, 0.5)
    res = _rand((B, H, W, N), g).bfloat16() if res_mode else None
    ln = (1 + 0.1 * _rand((N,), g), 0.1 * _rand((N,), g)) if use_ln else None
    acc = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, 1, 1).permute(0, 2, 3, 1)
    ref = _ref_epilogue(acc, b, pre, res.float() if res is not None else None, res_mode, ln, post, 1.0, N)

    pw = ops.pack_conv3x3(w.cuda(), b.cuda())
    lnp = (ops.pad_vec(ln[0].cuda(), pw.n_pad), ops.pad_vec(ln[1].cuda(), pw.n_pad)) if use_ln else None
    y = ops.conv_gemm(x.cuda(), pw, ln=lnp, pre_act=pre, post_act=post,
                      res=res.cuda() if res is not None else None, res_mode=res_mode)
    torch.cuda.synchronize()
    got = y.float().cpu()
    err = (got - ref).abs().max().item()
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    assert err < tol, (err, tol)


def test_channel_slices_and_f32_out():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = _rand((2, 10, 12, 96), g).bfloat16()
    w = _rand((2, 32, 3, 3), g, 0.1).bfloat16()
    acc = F.conv2d(x.float()[..., 32:64].permute(0, 3, 1, 2), w.float(), None, 1, 1).permute(0, 2, 3, 1)
    pw = ops.pack_conv3x3(w.cuda())
    y = ops.conv_gemm(x.cuda(), pw, x_coff=32, out_f32=True, bias=False)
    torch.cuda.synchronize()
    assert y.shape[-1] == 2
    assert (y.cpu() - acc).abs().max().item() < 2e-2
    # write into a channel slice of a wider bf16 buffer
    w2 = _rand((16, 32, 3, 3), g, 0.1).bfloat16()
    acc2 = F.conv2d(x.float()[..., 64:96].permute(0, 3, 1, 2), w2.float(), None, 1, 1).permute(0, 2, 3, 1)
    buf = torch.zeros(2, 10, 12, 48, dtype=torch.bfloat16, device="cuda")
    ops.conv_gemm(x.cuda(), ops.pack_conv3x3(w2.cuda()), x_coff=64, out=buf, y_coff=16, bias=False)
    torch.cuda.synchronize()
    got = buf.float().cpu()
    assert (got[..., 16:32] - acc2).abs().max().item() < 3e-2
    assert (got[..., :16] == 0).all() and (got[..., 32:] == 0).all()


@pytest.mark.parametrize("C,N,use_ln", [(64, 64, True), (64, 32, False), (32, 16, False)])
def test_fused_upsample_conv(C, N, use_ln):
    """`upconv`: nearest x2 + 3x3 conv (+ELU, + per-pixel LayerNorm) computed on the low-resolution input"""
    ops = _ops()
    g = torch.Generator().manual_seed(C + N)
    B, H, W = 2, 13, 18
    x = _rand((B, H, W, C), g).bfloat16()
    w = _rand((N, C, 3, 3), g, (9 * C) ** -0.5).bfloat16()
    ln = (1 + 0.1 * _rand((N,), g), 0.1 * _rand((N,), g)) if use_ln else None
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.elu(F.conv2d(up, w.float(), None, 1, 1)).permute(0, 2, 3, 1)
    if use_ln:
        ref = F.layer_norm(ref, (N,), ln[0], ln[1], 1e-5)
    pw = ops.pack_upconv3x3(w.cuda())
    lnp = (ln[0].cuda().contiguous(), ln[1].cuda().contiguous()) if use_ln else None
    y = ops.conv_gemm(x.cuda(), pw, bias=False, ln=lnp, pre_act=3 if use_ln else 0, post_act=0 if use_ln else 3)
    torch.cuda.synchronize()
    assert y.shape == (B, 2 * H, 2 * W, N)
    err = (y.float().cpu() - ref).abs().max().item()
    assert err < 3e-2 * max(1.0, ref.abs().max().item()), err     # phase filters are sums of bf16 taps, re-rounded to bf16


@pytest.mark.parametrize("B,H,W,K,N", [(2, 30, 40, 256, 512), (1, 15, 21, 64, 128), (3, 7, 9, 512, 1024)])
def test_linear_on_strided_pixels(B, H, W, K, N):
    """stride-2 1x1 projection (ResNet downsample): a Linear over x[:, ::2, ::2, :] read through a strided tensor map;
    odd extents included"""
    ops = _ops()
    g = torch.Generator().manual_seed(B + H + W + K)
    x = _rand((B, H, W, K), g).bfloat16()
    w = _rand((N, K), g, K ** -0.5).bfloat16()
    b = _rand((N,), g, 0.5)
    ref = x.float()[:, ::2, ::2, :] @ w.float().t() + b
    pw = ops.pack_linear(w.cuda(), b.cuda())
    out = ops.conv_gemm(x.cuda(), pw, subsample2=True)
    assert tuple(out.shape) == tuple(ref.shape)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err <= 2e-2 * max(1.0, ref.abs().max().item()), err      # bf16 output rounding


@pytest.mark.parametrize("rows,K,N,act,from_input", [(4800, 256, 2048, 1, False), (40000, 256, 64, 2, True), (9000, 2048, 256, 1, False),
                                                     (70000, 64, 256, 2, True), (777, 128, 96, 3, False), (5000, 512, 128, 4, False)])
def test_linear_with_activation_gradient_epilogue(rows, K, N, act, from_input):
    """res_mode = GWD_RES_MUL_ACTGRAD (training): y = (x W^T) * act'(saved) * scale, the activation backward fused into the
    data-gradient GEMM, on the TMA-epilogue path (N = 128 / 256 / 2048) and the thread-per-row path (N = 64 / 96)"""
    ops = _ops()
    g = torch.Generator().manual_seed(rows + N)
    x = _rand((rows, K), g).bfloat16()
    w = _rand((N, K), g, K ** -0.5).bfloat16()
    saved = _rand((rows, N), g).bfloat16()
    if act == 4 and not from_input:
        saved = torch.sigmoid(saved.float()).bfloat16()
    scale = 1.25
    acc = x.float() @ w.float().t()
    sv = saved.float()
    if from_input:
        v = sv.clone().requires_grad_(True)
        {1: torch.relu, 2: F.gelu, 3: F.elu, 4: torch.sigmoid}[act](v).sum().backward()
        fac = v.grad
    else:
        fac = {1: (sv > 0).float(), 3: torch.where(sv > 0, torch.ones_like(sv), sv + 1), 4: sv * (1 - sv)}[act]
    ref = acc * fac * scale
    pw = ops.pack_linear(w.cuda(), None)
    y = ops.conv_gemm(x.cuda(), pw, bias=False, act_grad=(saved.cuda(), act, from_input, 1.0, scale))
    torch.cuda.synchronize()
    err = (y.float().cpu()[:, :N] - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err


def test_conv3x3_with_activation_gradient_epilogue():
    """the same epilogue behind a 3x3 data-gradient convolution, single-CTA and CTA-pair tiles"""
    ops = _ops()
    for (B, H, W, C, N) in [(2, 30, 40, 64, 32), (4, 120, 160, 160, 160)]:
        g = torch.Generator().manual_seed(B + C)
        x = _rand((B, H, W, C), g).bfloat16()
        w = _rand((N, C, 3, 3), g, (9 * C) ** -0.5).bfloat16()
        saved = _rand((B, H, W, N), g).bfloat16()
        acc = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, 1, 1).permute(0, 2, 3, 1)
        sv = saved.float()
        ref = acc * torch.where(sv > 0, torch.ones_like(sv), sv + 1)          # ELU' from the output
        y = ops.conv_gemm(x.cuda(), ops.pack_conv3x3(w.cuda(), None), bias=False, act_grad=(saved.cuda(), 3, False, 1.0, 1.0))
        torch.cuda.synchronize()
        err = (y.float().cpu() - ref).abs().max().item()
        assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
