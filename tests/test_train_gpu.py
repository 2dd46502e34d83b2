"""GPU parity of the line-branch training path: every backward / optimizer kernel against plain torch fp32 on the same
seeded inputs, then the gradients of the whole branch (input_proj -> encoder -> decoder -> heads -> SetCriterion)
against torch.autograd over the CPU oracle, and a few optimisation steps.

Tolerances: activations and activation gradients are bf16 (8-bit mantissa); a kernel alone must match fp32 torch to
bf16 output rounding, the branch gradients (12 transformer layers deep) to a few per cent in relative L2 norm."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from helpers import oracle, synth, synth_weights

pytestmark = pytest.mark.gpu


def _ops():
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import ops
    return ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


def rel_l2(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("rows,C,with_add", [(600, 256, False), (4800, 256, True), (77, 512, True), (5, 64, False),
                                             (4099, 64, True), (1001, 128, False), (3333, 32, True), (2050, 96, False), (5000, 160, True), (4097, 160, False),
                                             (9001, 144, True), (4100, 168, False), (4098, 136, True), (30000, 152, False)])
def test_layernorm_bwd(rows, C, with_add):
    ops = _ops()
    g = _g(rows + C)
    z = (torch.randn(rows, C, generator=g) * 2 + 0.3).bfloat16()
    dy = torch.randn(rows, C, generator=g).bfloat16()
    gamma = torch.rand(C, generator=g) + 0.5
    add = torch.randn(rows, C, generator=g).bfloat16() if with_add else None
    zr = z.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = torch.zeros(C, requires_grad=True)
    F.layer_norm(zr, (C,), gr, br, 1e-5).backward(dy.float())
    ref_dz = zr.grad + (add.float() if with_add else 0)
    dgamma = torch.full((C,), 1.0, device="cuda")        # accumulation semantics: starts from a non-zero value
    dbeta = torch.zeros(C, device="cuda")
    dz = ops.layernorm_bwd(dy.cuda(), z.cuda(), gamma.cuda(), dgamma, dbeta, add=add.cuda() if with_add else None)
    assert rel_l2(dz, ref_dz) < 6e-3            # bf16 output rounding (2^-9 per element)
    assert rel_l2(dgamma - 1.0, gr.grad) < 1e-4
    assert rel_l2(dbeta, br.grad) < 1e-4


def test_conv_layernorm_gelu_block_backward():
    """ConvLn + GELU + residual (src/models/points/points_sample.py:12-43, the PyramidLayer block that carries most of the
    model's FLOPs): forward on gwd_conv_gemm with the pre-norm value kept, backward = gwd_layernorm_bwd(post_act=GELU) ->
    gwd_conv3x3_wgrad + dgrad on gwd_conv_gemm, against torch.autograd (erf GELU) on the same bf16 operands"""
    ops = _ops()
    g = _g(23)
    B, H, W, C = 2, 24, 32, 160
    x = torch.randn(B, H, W, C, generator=g).bfloat16()
    w = (torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5).bfloat16()
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    dy = torch.randn(B, H, W, C, generator=g).bfloat16()
    # reference
    xr, wr = x.float().permute(0, 3, 1, 2).requires_grad_(True), w.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    conv = F.conv2d(xr, wr, padding=1).permute(0, 2, 3, 1)
    y = F.gelu(F.layer_norm(conv, (C,), gr, br, 1e-5)) + xr.permute(0, 2, 3, 1)
    y.backward(dy.float())
    # CUDA path
    xc = x.cuda()
    z = torch.empty(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    yc = ops.conv_gemm(xc, ops.pack_conv3x3(w.float().cuda()), bias=False, ln=(gamma.cuda(), beta.cuda()), post_act=ops.ACT_GELU,
                       res=xc, res_mode=ops.RES_AFTER, y_raw=z)
    assert rel_l2(yc, y.detach()) < 1e-2
    dgam, dbet = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dz = ops.layernorm_bwd(dy.cuda(), z.view(-1, C), gamma.cuda(), dgam, dbet, beta=beta.cuda(), post_act=ops.ACT_GELU).view(B, H, W, C)
    dw = torch.zeros(9, C, C, device="cuda")
    ops.conv3x3_wgrad(dz, xc, dw)
    dx = ops.conv_gemm(dz, ops.pack_conv3x3_dgrad(w.float().cuda()), bias=False, res=dy.cuda(), res_mode=ops.RES_AFTER)
    assert rel_l2(ops.unpack_conv3x3_grad(dw, C, C), wr.grad) < 2e-2
    assert rel_l2(dgam, gr.grad) < 2e-2 and rel_l2(dbet, br.grad) < 2e-2
    assert rel_l2(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-2


def test_act_bwd_and_padding():
    ops = _ops()
    g = _g(3)
    dy = torch.randn(1200, 6, generator=g)
    y = torch.rand(1200, 6, generator=g)
    out = ops.act_bwd(dy.cuda(), y.cuda(), ops.ACT_SIGMOID, out_cols=16).float().cpu()
    ref = dy * y * (1 - y)
    assert out.shape == (1200, 16) and torch.equal(out[:, 6:], torch.zeros(1200, 10))
    assert torch.equal(out[:, :6], ref.bfloat16().float())
    h = torch.randn(333, 2048, generator=g).relu().bfloat16()
    dh = torch.randn(333, 2048, generator=g).bfloat16()
    out = ops.act_bwd(dh.cuda(), h.cuda(), ops.ACT_RELU).cpu()
    assert torch.equal(out, torch.where(h > 0, dh, torch.zeros_like(dh)))
    e = F.elu(torch.randn(333, 64, generator=g)).bfloat16()
    de = torch.randn(333, 64, generator=g).bfloat16()
    out = ops.act_bwd(de.cuda(), e.cuda(), ops.ACT_ELU).float().cpu()
    ref_e = torch.where(e.float() > 0, de.float(), de.float() * (e.float() + 1)).bfloat16().float()
    assert torch.equal(out, ref_e)
    out = ops.act_bwd(dy[:, :2].contiguous().cuda(), None, ops.ACT_NONE, out_cols=16).float().cpu()
    assert torch.equal(out[:, :2], dy[:, :2].bfloat16().float()) and float(out[:, 2:].abs().max()) == 0.0


@pytest.mark.parametrize("rows,C", [(600, 256), (200, 16), (4800, 2048), (130, 768)])
def test_transpose_and_colsum(rows, C):
    ops = _ops()
    x = torch.randn(rows, C, generator=_g(rows)).bfloat16()
    cs = torch.zeros(C, device="cuda")
    out = ops.transpose(x.cuda(), colsum=cs).cpu()
    rp = (rows + 63) // 64 * 64
    assert out.shape == (C, rp)
    assert torch.equal(out[:, :rows], x.t()) and float(out[:, rows:].float().abs().max() if rp > rows else 0.0) == 0.0
    assert rel_l2(cs, x.float().sum(0)) < 1e-5


def test_transpose_batch():
    ops = _ops()
    g = _g(9)
    flat = torch.randn(1_400_000, generator=g).bfloat16().cuda()
    shapes, off, pairs = [(16, 256), (768, 256), (2048, 256), (256, 2048), (112, 256), (64, 80)], 0, []
    for r, c in shapes:
        pairs.append((flat[off:off + r * c].view(r, c), torch.zeros(c, r, dtype=torch.bfloat16, device="cuda")))
        off += r * c
    pairs.append((pairs[1][0][256:512], torch.zeros(256, 256, dtype=torch.bfloat16, device="cuda")))     # a row slice
    ops.transpose_batch(ops.transpose_batch_tables(pairs))
    for src, dst in pairs:
        assert torch.equal(dst, src.t())


def test_weight_gradient_gemm_on_transposed_operands():
    """dW = dY^T X and dX = dY W through gwd_conv_gemm (fp32 output, multi-tile N, K = rows padded to 64)"""
    ops = _ops()
    g = _g(11)
    for R, N, K in ((600, 256, 2048), (200, 2048, 256), (1200, 16, 256)):
        dY = (torch.randn(R, N, generator=g) * 0.1).bfloat16()
        X = torch.randn(R, K, generator=g).bfloat16()
        W = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
        dYT, XT = ops.transpose(dY.cuda()), ops.transpose(X.cuda())
        gw = torch.empty(N, K, dtype=torch.float32, device="cuda")
        ops.conv_gemm(dYT, ops.PackedWeight(XT.view(1, K, XT.shape[1]), None, 1, K, XT.shape[1]), out=gw, out_f32=True, bias=False)
        assert rel_l2(gw, dY.float().t() @ X.float()) < 1e-3
        WT = ops.transpose(W.cuda(), pad_to=16)
        dX = ops.conv_gemm(dY.cuda(), ops.PackedWeight(WT.view(1, K, N), None, 1, K, N), bias=False)
        assert rel_l2(dX, dY.float() @ W.float()) < 6e-3


@pytest.mark.parametrize("R,N,K", [(600, 256, 2048), (4800, 2048, 256), (1200, 16, 256), (9600, 256, 256), (70, 512, 264),
                                   (40000, 192, 64), (33001, 64, 64), (50000, 384, 192), (40000, 160, 80), (36000, 16, 128), (40000, 256, 64), (34000, 512, 256), (33000, 1024, 64)])
def test_linear_wgrad_kernel(R, N, K):
    """gwd_linear_wgrad: dW += dY^T X, db += column sums (accumulating into a non-zero buffer), operands read in place.
    From 32 768 rows on (N, K multiples of 16) the tcgen05 kernel of gwd_wgrad_tc.cu runs (one "tap", 64-row TMA boxes, both role
    assignments, a partial last box) with gwd_colsum_kernel for the bias gradient; below, the split-K mma.sync kernel."""
    ops = _ops()
    g = _g(R + N)
    dY = (torch.randn(R, N, generator=g) * 0.1).bfloat16()
    X = torch.randn(R, K + 8, generator=g).bfloat16()            # wider buffer: the kernel reads a column slice
    dw = torch.full((N, K), 0.5, device="cuda")
    db = torch.full((N,), -1.0, device="cuda")
    ops.linear_wgrad(dY.cuda(), X.cuda(), dw, db, K=K, x_coff=8)
    ref = dY.double().t() @ X[:, 8:].double()
    assert rel_l2(dw - 0.5, ref) < 2e-5                          # fp32 accumulation of exact bf16 products
    assert rel_l2(db + 1.0, dY.double().sum(0)) < 2e-5


@pytest.mark.parametrize("B,H,W,C,N", [(2, 24, 40, 160, 160), (1, 17, 23, 64, 32), (3, 8, 8, 32, 16), (1, 30, 40, 800, 320),
                                       (2, 37, 70, 32, 32), (1, 64, 130, 64, 64), (2, 50, 100, 32, 16), (1, 70, 64, 64, 32),
                                       (3, 5, 200, 16, 64), (1, 120, 160, 64, 64)])
def test_conv3x3_backward_building_blocks(B, H, W, C, N):
    """weight gradient (gwd_conv3x3_wgrad, tap-shifted split-K mma.sync) and data gradient (gwd_conv_gemm with the
    transposed / flipped filter) of a stride-1 3x3 convolution vs torch.autograd on the same bf16 operands"""
    ops = _ops()
    g = _g(B * H + C)
    x = torch.randn(B, H, W, C, generator=g).bfloat16()
    w = (torch.randn(N, C, 3, 3, generator=g) * (9 * C) ** -0.5).bfloat16()
    dy = torch.randn(B, H, W, N, generator=g).bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    dw = torch.full((9, N, C), 0.25, device="cuda")          # accumulation semantics
    db = torch.zeros(N, device="cuda")
    ops.conv3x3_wgrad(dy.cuda(), x.cuda(), dw, db)
    assert rel_l2(ops.unpack_conv3x3_grad(dw - 0.25, N, C), wr.grad) < 2e-4
    assert rel_l2(db, dy.float().sum((0, 1, 2))) < 2e-5
    dx = ops.conv_gemm(dy.cuda(), ops.pack_conv3x3_dgrad(w.float().cuda()), bias=False)
    assert rel_l2(dx[..., :C], xr.grad.permute(0, 2, 3, 1)) < 6e-3


@pytest.mark.parametrize("B,H,W,C,N,cs", [(2, 64, 80, 160, 160, 0), (1, 70, 90, 800, 320, 0), (3, 37, 50, 80, 160, 32),
                                          (1, 120, 160, 160, 80, 0), (2, 48, 48, 320, 320, 16), (1, 66, 67, 96, 208, 0)])
def test_conv3x3_wgrad_tensor_core_path(B, H, W, C, N, cs):
    """gwd_conv3x3_wgrad without bias on maps of >= 4096 pixels and more than 64 channels runs on the tcgen05 kernel
    (gwd_wgrad_tc.cu: MN-major TMA operands, tap shift = box offset, zero padding = out-of-bounds fill; partial 16 x 4 pixel
    blocks, channel counts that are not multiples of 64 / 128, both role assignments, channel-slice operands, += semantics):
    equal to torch.autograd on the same bf16 operands"""
    ops = _ops()
    g = _g(B * H + C + N)
    xw = torch.randn(B, H, W, C + cs, generator=g).bfloat16()
    dyw = torch.randn(B, H, W, N + cs, generator=g).bfloat16()
    w = torch.zeros(N, C, 3, 3, requires_grad=True)
    F.conv2d(xw[..., :C].float().permute(0, 3, 1, 2), w, padding=1).backward(dyw[..., :N].float().permute(0, 3, 1, 2))
    dw = torch.full((9, N, C), 0.25, device="cuda")          # accumulation semantics
    xc, dyc = xw.cuda(), dyw.cuda()
    capi = __import__("gwdepth_b200").capi
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(capi.lib().gwd_conv3x3_wgrad(ctypes.c_void_p(dyc.data_ptr()), N + cs, ctypes.c_void_p(xc.data_ptr()), C + cs, B, H, W, N, C,
                                            ctypes.c_void_p(dw.data_ptr()), None, st), "gwd_conv3x3_wgrad")
    assert rel_l2(ops.unpack_conv3x3_grad(dw - 0.25, N, C), w.grad) < 2e-4


@pytest.mark.parametrize("use_o", [True, False])      # True: tensor-core kernel (needs the forward output), False: CUDA-core kernel
@pytest.mark.parametrize("B,Lq,Lk,fused", [(2, 300, 300, True), (3, 100, 300, False), (2, 100, 100, True), (1, 37, 480, False),
                                           (1, 512, 512, True), (2, 1, 5, False)])
def test_attention_bwd(B, Lq, Lk, fused, use_o):
    ops = _ops()
    heads, hd = 8, 32
    E = heads * hd
    g = _g(Lq * 7 + Lk)
    scale = hd ** -0.5
    if fused:           # q | k side by side in one [rows, 2E] buffer, as the self-attention projections produce them
        qk = torch.randn(B * Lq, 2 * E, generator=g).bfloat16().cuda()
        q, k, q_rs, k_rs = qk, qk[:, E:], 2 * E, 2 * E
    else:
        q, k = torch.randn(B * Lq, E, generator=g).bfloat16().cuda(), torch.randn(B * Lk, E, generator=g).bfloat16().cuda()
        q_rs = k_rs = E
    v = torch.randn(B * Lk, E, generator=g).bfloat16().cuda()
    d_o = torch.randn(B * Lq, E, generator=g).bfloat16().cuda()

    def heads_of(t, L, rs, off=0):
        return t.float().view(B, L, rs)[:, :, off:off + E].reshape(B, L, heads, hd).permute(0, 2, 1, 3).clone().requires_grad_(True)

    qh, kh, vh = heads_of(q, Lq, q_rs), heads_of(qk if fused else k, Lk, k_rs, E if fused else 0), heads_of(v, Lk, E)
    o = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1) @ vh
    o.backward(d_o.float().view(B, Lq, heads, hd).permute(0, 2, 1, 3))
    flat = lambda t, L: t.permute(0, 2, 1, 3).reshape(B * L, E)
    if fused:
        dqk = torch.empty(B * Lq, 2 * E, dtype=torch.bfloat16, device="cuda")
        dq, dk = dqk, dqk[:, E:]
    else:
        dq = torch.empty(B * Lq, E, dtype=torch.bfloat16, device="cuda")
        dk = torch.empty(B * Lk, E, dtype=torch.bfloat16, device="cuda")
    dv = torch.empty(B * Lk, E, dtype=torch.bfloat16, device="cuda")
    o_fwd = flat(o.detach(), Lq).bfloat16().cuda().contiguous() if use_o else None
    ops.attention_bwd(q, k, v, d_o, dq, dk, dv, items=B, heads=heads, Lq=Lq, Lk=Lk, hd=hd, q_strides=(Lq * q_rs, q_rs),
                      k_strides=(Lk * k_rs, k_rs), v_strides=(Lk * E, E), do_strides=(Lq * E, E), dq_strides=(Lq * q_rs, q_rs),
                      dk_strides=(Lk * k_rs, k_rs), dv_strides=(Lk * E, E), scale=scale, o=o_fwd, o_strides=(Lq * E, E))
    got_dq = dqk[:, :E] if fused else dq
    got_dk = dqk[:, E:] if fused else dk
    # CUDA-core kernel: fp32 throughout, bf16 output rounding only; tensor-core kernel: P and dS are bf16 MMA operands and
    # D comes from the bf16 forward output
    tol = 1e-2 if use_o else 6e-3
    assert rel_l2(got_dq, flat(qh.grad, Lq)) < tol
    assert rel_l2(got_dk, flat(kh.grad, Lk)) < tol
    assert rel_l2(dv, flat(vh.grad, Lk)) < tol


@pytest.mark.parametrize("B,Lq,Lk,pad", [(2, 300, 300, True), (3, 100, 300, True), (1, 850, 850, False), (2, 800, 800, True), (2, 100, 1216, True),
                                        (1, 1200, 1200, False), (1, 513, 64, False), (1, 1280, 1280, True)])
def test_attention_bwd_masks_and_long_sequences(B, Lq, Lk, pad):
    """gwd_attention_bwd with a key-padding mask (padded batches: src_key_padding_mask of nn.MultiheadAttention) and on sequences
    beyond 512 tokens (two launches: per query block with K, V resident, per key block with Q, dO resident) vs torch.autograd"""
    ops = _ops()
    heads, hd = 8, 32
    E = heads * hd
    g = _g(Lq * 3 + Lk + int(pad))
    scale = hd ** -0.5
    q = torch.randn(B * Lq, E, generator=g).bfloat16().cuda()
    k = torch.randn(B * Lk, E, generator=g).bfloat16().cuda()
    v = torch.randn(B * Lk, E, generator=g).bfloat16().cuda()
    d_o = torch.randn(B * Lq, E, generator=g).bfloat16().cuda()
    mask = torch.zeros(B, Lk, dtype=torch.bool)
    if pad:          # ragged tails + a hole, different per image
        for b in range(B):
            mask[b, Lk - 17 - 40 * b:] = True
            mask[b, 5 + b:9 + b] = True

    def heads_of(t, L):
        return t.float().view(B, L, heads, hd).permute(0, 2, 1, 3).clone().requires_grad_(True)
    qh, kh, vh = heads_of(q, Lq), heads_of(k, Lk), heads_of(v, Lk)
    sc = qh @ kh.transpose(-1, -2) * scale
    sc = sc.masked_fill(mask.cuda()[:, None, None, :], float("-inf"))
    o = torch.softmax(sc, dim=-1) @ vh
    o.backward(d_o.float().view(B, Lq, heads, hd).permute(0, 2, 1, 3))
    flat = lambda t, L: t.permute(0, 2, 1, 3).reshape(B * L, E)
    dq = torch.empty(B * Lq, E, dtype=torch.bfloat16, device="cuda")
    dk = torch.full((B * Lk, E), 7.0, dtype=torch.bfloat16, device="cuda")
    dv = torch.full((B * Lk, E), 7.0, dtype=torch.bfloat16, device="cuda")
    o_fwd = flat(o.detach(), Lq).bfloat16().contiguous()
    ops.attention_bwd(q, k, v, d_o, dq, dk, dv, items=B, heads=heads, Lq=Lq, Lk=Lk, hd=hd, q_strides=(Lq * E, E), k_strides=(Lk * E, E),
                      v_strides=(Lk * E, E), do_strides=(Lq * E, E), dq_strides=(Lq * E, E), dk_strides=(Lk * E, E), dv_strides=(Lk * E, E),
                      scale=scale, o=o_fwd, o_strides=(Lq * E, E), key_padding=mask.cuda() if pad else None)
    assert rel_l2(dq, flat(qh.grad, Lq)) < 1e-2
    assert rel_l2(dk, flat(kh.grad, Lk)) < 1e-2
    assert rel_l2(dv, flat(vh.grad, Lk)) < 1e-2
    if pad:          # padded keys get exact zeros
        m = mask.view(-1).cuda()
        assert (dk[m] == 0).all() and (dv[m] == 0).all()


@pytest.mark.parametrize("max_norm,world", [(0.1, 1), (0.0, 1), (0.1, 2)])
def test_fused_clip_adamw_matches_torch(max_norm, world):
    ops = _ops()
    g = _g(5)
    n = 100_003
    p0 = torch.randn(n, generator=g)
    P = p0.clone().cuda()
    M, V = torch.zeros_like(P), torch.zeros_like(P)
    mirror = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ref = torch.nn.Parameter(p0.clone().cuda())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    ssq = torch.zeros(1, dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, generator=g).cuda() * (0.5 if step == 2 else 1e-4)     # step 2 clips, the others do not
        ref.grad = grad.clone()
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_([ref], max_norm)
        opt.step()
        ssq.zero_()
        G = grad * world                                # what a sum all-reduce over `world` equal ranks leaves behind
        ops.sumsq(G, ssq)
        ops.adamw_step(P, G, M, V, mirror, lr=1e-3, weight_decay=1e-2, step=step, max_norm=max_norm, grad_scale=1.0 / world,
                       sumsq_buf=ssq)
        assert abs(float(ssq.sqrt()) - float(G.double().norm())) < 1e-9 * float(G.double().norm()) + 1e-12
        assert rel_l2(P, ref.data) < 2e-6
        assert torch.equal(mirror, P.bfloat16())


def test_flat_adamw_is_a_drop_in_for_clip_plus_torch_adamw():
    """optim.FlatAdamW on a real module with the reference's two parameter groups (src/main_glassrgbd.py:59-66) and
    autograd gradients: same trajectory as clip_grad_norm_ + torch.optim.AdamW"""
    import copy
    import gwdepth_b200  # noqa: F401
    from gwdepth_b200 import optim
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.LayerNorm(64),
                              torch.nn.Linear(64, 3)).cuda()
    net[2].weight.requires_grad_(False)                                  # a frozen tensor, as the backbone stem is
    ref = copy.deepcopy(net)

    def groups(m):
        return [{"params": [p for n, p in m.named_parameters() if not n.startswith("0.") and p.requires_grad]},
                {"params": [p for n, p in m.named_parameters() if n.startswith("0.") and p.requires_grad], "lr": 1e-3}]
    opt = optim.FlatAdamW(groups(net), lr=1e-2, weight_decay=1e-4, max_norm=0.1)
    ref_opt = torch.optim.AdamW(groups(ref), lr=1e-2, weight_decay=1e-4)
    x = torch.randn(50, 37, device="cuda")
    y = torch.randn(50, 3, device="cuda")
    for it in range(5):
        for m, o in ((net, opt), (ref, ref_opt)):
            o.zero_grad()
            torch.nn.functional.mse_loss(m(x), y).mul(10.0 if it == 2 else 1e-3).backward()     # step 2 is clipped
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
        ref_opt.step()
        opt.step()
        for (n, a), b in zip(net.named_parameters(), ref.parameters()):
            assert rel_l2(a, b) < 5e-6, (it, n)
    assert torch.equal(net[2].weight, ref[2].weight)                     # frozen tensors untouched
    assert all(p.data_ptr() >= g["P"].data_ptr() for g in opt.groups for p in g["params"])
    # it is a torch Optimizer: the reference's StepLR (src/main_glassrgbd.py:67) drives its learning rates, and a resumed run
    # (state_dict -> load_state_dict, :160-163) continues on the same trajectory
    assert isinstance(opt, torch.optim.Optimizer)
    sched, ref_sched = torch.optim.lr_scheduler.StepLR(opt, 1), torch.optim.lr_scheduler.StepLR(ref_opt, 1)
    sched.step(); ref_sched.step()
    assert abs(opt.param_groups[0]["lr"] - 1e-3) < 1e-12 and abs(opt.param_groups[1]["lr"] - 1e-4) < 1e-12
    state = opt.state_dict()
    net2 = copy.deepcopy(ref)
    for a, b in zip(net2.parameters(), net.parameters()):
        a.data.copy_(b.data)
    opt2 = optim.FlatAdamW(groups(net2), lr=1e-2, weight_decay=1e-4, max_norm=0.1)
    opt2.load_state_dict(state)
    for m, o in ((net, opt), (net2, opt2), (ref, ref_opt)):
        o.zero_grad()
        torch.nn.functional.mse_loss(m(x), y).mul(1e-3).backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.1)
    ref_opt.step(); opt.step(); opt2.step()
    for (n, a), b, c in zip(net.named_parameters(), ref.parameters(), net2.parameters()):
        assert rel_l2(a, b) < 5e-6 and torch.equal(a, c), n


# ------------------------------------------------------------------------------------------ the branch
_cache = {}


def setup(B=2, H=224, W=320):
    key = (B, H, W)
    if key not in _cache:
        import gwdepth_b200  # noqa: F401
        from gwdepth_b200 import model as M, train
        net, criterions, _ = M.build_model(M.default_args(device="cuda", dropout=0.0))
        net.load_state_dict(synth_weights())
        net.cuda().eval()
        images, targets, _, _ = synth.synth_batch(B, H, W, seed=3)
        with torch.no_grad():
            c5 = net.plan().backbone(images.cuda().float().contiguous())[3].contiguous()
        targets = [{k: v.cuda() for k, v in t.items()} for t in targets]
        _cache[key] = (net, criterions[0], train, c5, targets)
    return _cache[key]


def oracle_branch(sd, c5_nchw, cfg):
    """the oracle's line branch with autograd enabled (oracle.forward wraps it in no_grad)"""
    p = oracle.P(sd)
    B, _, h, w = c5_nchw.shape
    m5 = torch.zeros(B, h, w, dtype=torch.bool)
    pos5 = oracle.sine_position(m5, cfg["hidden_dim"] // 2, True)
    src = F.conv2d(c5_nchw, sd["input_proj.weight"], sd["input_proj.bias"])
    hs, _ = oracle.detr_transformer(src, m5, sd["query_embed.weight"], pos5, p.sub("transformer"), cfg)
    logits = oracle.linear(hs, p, "class_embed")
    t = hs
    for i in range(3):
        t = oracle.linear(t, p, "lines_embed.layers.%d" % i)
        if i < 2:
            t = F.relu(t)
    return logits, t.sigmoid()


# (encoder layers, decoder layers, tolerance on the global relative L2 error of the gradient).  Linear paths agree to ~1 %
# (class_embed, lines_embed.layers.2); behind every ReLU the derivative is evaluated at bf16 activations that differ
# slightly from the oracle's fp32 ones (a unit that changes sign on 0.1 % of its inputs is a 3 % L2 difference), which
# saturates at ~6 % after two ReLU layers and does NOT grow with depth (measured: 6.1 % at 1+1 layers, 5.8 % at 6+6).
@pytest.mark.parametrize("enc,dec,tol", [(1, 1, 8e-2), (6, 6, 8e-2)])
def test_line_branch_gradients_match_oracle_autograd(enc, dec, tol):
    net, criterion, train, c5, targets = setup()
    cfg_mine = dict(net.cfg, enc_layers=enc, dec_layers=dec)
    lb = train.LineBranch(synth_weights(), cfg_mine)
    logits, lines = lb.forward(c5)
    # the oracle on the same C5 map, fp32, autograd on every branch parameter and on C5
    cfg = dict(oracle.DEFAULT_CFG, enc_layers=enc, dec_layers=dec)
    sd = {k: v.clone().float() for k, v in synth_weights().items() if v.is_floating_point()}
    names = list(lb.index)
    for k in names:
        sd[k].requires_grad_(True)
    c5_ref = c5.float().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    lo, li = oracle_branch(sd, c5_ref, cfg)
    assert rel_l2(logits, lo) < 4e-2 and rel_l2(lines, li) < 2e-2
    # same assignments on both sides (the ones the CUDA path's matcher makes), same loss weights
    total, losses, dc5 = lb.loss_and_grads(c5, targets, criterion)
    tl = [t["lines"].cpu() for t in targets]
    num_items = max(float(sum(len(t) for t in tl)), 1.0)
    ref_total = 0.0
    for s in range(lo.shape[0]):
        idx = criterion.matcher({"pred_logits": logits[s], "pred_lines": lines[s]}, targets)
        ce, l1 = oracle.set_losses(lo[s], li[s], tl, idx, num_items, 0.1)
        ref_total = ref_total + ce * 1.0 + l1 * 5.0
    assert abs(float(total) - float(ref_total)) < 2e-2 * abs(float(ref_total))
    # The L1 line loss has a discontinuous gradient (sign(pred - target)): the ~0.5 % forward difference between the bf16
    # path and the fp32 oracle flips the sign of ~0.6 % of the matched coordinates, which alone is a 15-20 % L2 difference
    # in EVERY downstream gradient.  The backward kernels are therefore checked as a vector-Jacobian product: the oracle
    # is differentiated with the SAME output cotangents (dL/dlogits, dL/dlines of the CUDA path's criterion).
    dlogits, dlines = lb.last_cotangents
    ref_grads = torch.autograd.grad([lo, li], [sd[k] for k in names] + [c5_ref], [dlogits.cpu(), dlines.cpu()], allow_unused=True)
    ref_grads = [torch.zeros_like(t) if g is None else g for g, t in zip(ref_grads, [sd[k] for k in names] + [c5_ref])]
    got = lb.grads()
    num = den = 0.0
    worst = (0.0, "")
    for k, gr in zip(names, ref_grads[:-1]):
        gg = got[k].double().cpu().reshape(-1)
        gr = gr.double().reshape(-1)
        num += float((gg - gr).pow(2).sum())
        den += float(gr.pow(2).sum())
        if float(gr.norm()) > 1e-3 * (den ** 0.5):       # parameters that carry a non-negligible share of the gradient
            cos = float(gg @ gr / (gg.norm() * gr.norm()).clamp_min(1e-30))
            worst = max(worst, (1.0 - cos, k))
    print("line-branch gradient: global rel-L2 %.4f, worst 1-cos %.4f at %s" % ((num / den) ** 0.5, worst[0], worst[1]))
    assert (num / den) ** 0.5 < tol, "global relative L2 error of the branch gradient %.4f" % (num / den) ** 0.5
    assert worst[0] < 1.5e-2, "direction of d%s off: 1 - cos = %.4f" % (worst[1], worst[0])
    ref_dc5 = ref_grads[-1].permute(0, 2, 3, 1).reshape(dc5.shape)
    assert rel_l2(dc5, ref_dc5) < 1.5 * tol
    # padded rows of the flat gradient stay exactly zero (so AdamW keeps the pads at zero)
    ce_w = lb.view(lb.G, "class_embed.weight")
    assert float(ce_w[2:].abs().max()) == 0.0


def test_line_branch_training_lowers_the_loss_and_keeps_mirror_in_sync():
    net, criterion, train, c5, targets = setup()
    # Adam moves every weight by ~lr per step whatever the gradient scale; on this random-init, high-gain network the
    # reference's lr (1e-4, meant for a pretrained DETR) overshoots, so the descent check uses a small step
    lb = train.LineBranch(synth_weights(), net.cfg, lr=3e-6)
    first = None
    for it in range(6):
        total, _ = lb.train_step(c5, targets, criterion)
        first = float(total) if first is None else first
        assert torch.isfinite(total)
    assert float(total) < first, "loss did not go down: %.4f -> %.4f" % (first, float(total))
    assert torch.equal(lb.Wb, lb.P.bfloat16())
    assert lb.t == 6


def test_stacked_criterion_equals_stage_by_stage_criterion():
    """SetCriterion.forward_stacked (one matching launch, batched losses) == SetCriterion.forward (the reference's
    stage-by-stage structure): same assignments bit for bit, same 12 losses, same gradients"""
    net, criterion, train, c5, targets = setup()
    lb = train.LineBranch(synth_weights(), net.cfg)
    logits, lines = lb.forward(c5)
    a_lo, a_li = logits.detach().clone().requires_grad_(True), lines.detach().clone().requires_grad_(True)
    b_lo, b_li = logits.detach().clone().requires_grad_(True), lines.detach().clone().requires_grad_(True)
    la = criterion.forward_stacked(a_lo, a_li, targets)
    out = {"pred_logits": b_lo[-1], "pred_lines": b_li[-1],
           "aux_outputs": [{"pred_logits": x, "pred_lines": y} for x, y in zip(b_lo[:-1], b_li[:-1])]}
    lb_ = criterion(out, targets)
    assert set(la) == set(lb_)
    for s, stage in enumerate(criterion.indices_of_last_call()):
        ref = criterion.matcher({"pred_logits": logits[s], "pred_lines": lines[s]}, targets)
        for (i, j), (ri, rj) in zip(stage, ref):
            assert torch.equal(i, ri) and torch.equal(j, rj)
    for k in la:
        assert abs(float(la[k]) - float(lb_[k])) <= 1e-5 * max(1.0, abs(float(lb_[k]))), k
    wd = criterion.weight_dict
    sum(la[k] * wd[k] for k in la).backward()
    sum(lb_[k] * wd[k] for k in lb_).backward()
    assert rel_l2(a_lo.grad, b_lo.grad) < 1e-5 and rel_l2(a_li.grad, b_li.grad) < 1e-5
    # the fused kernel (forward + backward of the criterion in one launch, no autograd) against both
    lc, dlo, dli = criterion.forward_backward_stacked(logits, lines, targets)
    assert set(lc) == set(lb_)
    for k in lc:
        assert abs(float(lc[k]) - float(lb_[k])) <= 2e-6 * max(1.0, abs(float(lb_[k]))), k
    assert rel_l2(dlo, b_lo.grad) < 2e-6 and rel_l2(dli, b_li.grad) < 2e-6
    assert torch.equal(dli != 0, b_li.grad != 0)


@pytest.mark.parametrize("case", ["one_image_without_lines", "no_lines_at_all", "more_targets_than_queries"])
def test_criterion_edge_cases(case):
    """images without any target line (the reference matches nothing and normalises by max(num_items, 1)) and an image
    with more targets than queries (only Q of them can be matched): the three criterion paths agree"""
    net, criterion, train, c5, targets = setup()
    g = _g(17)
    S, B, Q = 6, 2, 100
    logits = torch.randn(S, B, Q, 2, generator=g).cuda()
    lines = torch.rand(S, B, Q, 6, generator=g).cuda()
    n = {"one_image_without_lines": (13, 0), "no_lines_at_all": (0, 0), "more_targets_than_queries": (120, 5)}[case]
    tg = [{"lines": torch.rand(k, 6, generator=g).cuda(), "labels": torch.zeros(k, dtype=torch.int64).cuda()} for k in n]
    a_lo, a_li = logits.clone().requires_grad_(True), lines.clone().requires_grad_(True)
    out = {"pred_logits": a_lo[-1], "pred_lines": a_li[-1],
           "aux_outputs": [{"pred_logits": x, "pred_lines": y} for x, y in zip(a_lo[:-1], a_li[:-1])]}
    ref = criterion(out, tg)
    wd = criterion.weight_dict
    sum(ref[k] * wd[k] for k in ref).backward()
    stacked = criterion.forward_stacked(logits, lines, tg)
    fused, dlo, dli = criterion.forward_backward_stacked(logits, lines, tg)
    for k in ref:
        assert abs(float(stacked[k]) - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), k
        assert abs(float(fused[k]) - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), k
    assert rel_l2(dlo, a_lo.grad) < 1e-5
    if sum(n) > 0:
        assert rel_l2(dli, a_li.grad) < 1e-5
    else:
        assert float(dli.abs().max()) == 0.0 and float(a_li.grad.abs().max()) == 0.0
    matched = sum(len(i) for i, _ in criterion.indices_of_last_call()[0])
    assert matched == sum(min(k, Q) for k in n)


def test_graphed_and_eager_training_steps_agree():
    """the CUDA-graph replay (forward graph + backward graph) must produce the same gradients as kernel-by-kernel launches"""
    net, criterion, train, c5, targets = setup()
    a, b = train.LineBranch(synth_weights(), net.cfg), train.LineBranch(synth_weights(), net.cfg)
    a.use_cuda_graph, b.use_cuda_graph = True, False
    for it in range(2):         # second iteration: replay of an existing capture after an optimizer step
        ta, _, dca = a.loss_and_grads(c5, targets, criterion)
        tb, _, dcb = b.loss_and_grads(c5, targets, criterion)
        if it == 0:
            # identical weights: identical forward, assignments and data gradients; weight / bias / LayerNorm gradients are
            # reduced with fp32 atomics (order-dependent in the last bits)
            assert abs(float(ta) - float(tb)) <= 1e-6 * abs(float(tb))
            assert rel_l2(a.G, b.G) < 1e-5 and torch.equal(dca, dcb)
        else:
            # after a step the two weight sets differ in the last bits, and the L1 loss / Hungarian matching are
            # discontinuous in them: only the loss is compared
            assert abs(float(ta) - float(tb)) <= 1e-3 * abs(float(tb))
        a.step()
        b.step()
    assert rel_l2(a.P, b.P) < 1e-4
