"""Micro-benchmarks of the non-GEMM kernels at model sizes (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000


def main(which):
    dev = "cuda"
    if "diffuse" in which:
        B, heads, P, R = 16, 16, 441, 40
        a0 = torch.randn(B, heads, P, R, device=dev)
        a1, raw = torch.empty_like(a0), torch.empty_like(a0)
        st = torch.empty(B * heads * 2, dtype=torch.float64, device=dev)
        w, b = (torch.randn(16, 16, 3, 3) * 0.1).contiguous(), torch.zeros(16)
        print("ref_diffuse (conv+norm)  %8.1f us" % timeit(lambda: ops.ref_diffuse(a0, a1, w, b, raw, st, B, heads, P, R)))
    if "ln" in which:
        x = torch.randn(16 * 120 * 160, 320, device=dev).bfloat16()
        g, bt = torch.ones(320, device=dev), torch.zeros(320, device=dev)
        out = torch.empty_like(x)
        us = timeit(lambda: ops.layernorm(x, g, bt, act=ops.ACT_GELU, out=out))
        print("layernorm+gelu 307200x320 %8.1f us  %.2f TB/s" % (us, 2 * x.numel() * 2 / us / 1e6))
        x = torch.randn(16 * 120 * 160, 64, device=dev).bfloat16()
        g, bt = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        us = timeit(lambda: ops.window_gather(x.view(16, 120, 160, 64), 16, 120, 160, 7, 3, g, bt))
        print("window_gather 1/4 C=64    %8.1f us  %.2f TB/s" % (us, 2 * x.numel() * 2 / us / 1e6))
        win = ops.window_gather(x.view(16, 120, 160, 64), 16, 120, 160, 7, 3, g, bt)
        us = timeit(lambda: ops.window_merge(win, x, 16, 120, 160, 7, 3, g, bt, want_ln=True))
        print("window_merge+LN 1/4 C=64  %8.1f us  %.2f TB/s" % (us, 4 * x.numel() * 2 / us / 1e6))
    if "pyr" in which:
        B, H, W, C = 16, 120, 160, 160
        cat = torch.randn(B, H, W, 5 * C, device=dev).bfloat16()
        us = timeit(lambda: [ops.avgpool(cat, k, C=C) for k in (16, 8, 4, 2)])
        print("avgpool x4 (16,8,4,2)      %8.1f us" % us)
        us = timeit(lambda: ops.avgpool_pyramid(cat, C=C))
        print("avgpool_pyramid            %8.1f us  %.2f TB/s" % (us, B * H * W * C * 2 / us / 1e6))
        pyr = ops.avgpool_pyramid(cat, C=C)
        us = timeit(lambda: [ops.bilinear_up_into(p_, cat, j * C, H, W) for j, p_ in enumerate(pyr, start=1)])
        print("bilinear_up x4             %8.1f us" % us)
        us = timeit(lambda: ops.bilinear_up4_into(pyr, cat, C, H, W))
        print("bilinear_up4               %8.1f us  %.2f TB/s" % (us, B * H * W * 4 * C * 2 / us / 1e6))
    if "attn" in which:
        B, L, E = 16, 300, 256
        qk = torch.randn(B * L, 2 * E, device=dev).bfloat16()
        v = torch.randn(B * L, E, device=dev).bfloat16()
        o = torch.empty_like(v)
        us = timeit(lambda: ops.attention(qk, qk[:, E:], v, o, items=B, heads=8, Lq=L, Lk=L, hd=32, q_strides=(L * 2 * E, 2 * E),
                                          k_strides=(L * 2 * E, 2 * E), v_strides=(L * E, E), o_strides=(L * E, E)))
        print("DETR encoder attention    %8.1f us  %.1f TFLOP/s" % (us, 4 * B * 8 * L * L * 32 / us / 1e6))


if __name__ == "__main__":
    main(sys.argv[1:] or ["diffuse", "ln", "pyr", "attn"])
