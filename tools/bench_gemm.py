"""Micro-benchmark of gwd_conv_gemm on the model's dominant shapes (CUDA events, L2-sized rotation of inputs).
usage: python tools/bench_gemm.py [case ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import ops  # noqa: E402

CASES = {
    # name: (B, H, W, Cin, Cout, taps, ln, act)
    "pyr160_plain": (16, 120, 160, 160, 160, 9, False, 0),
    "pyr160_ln_gelu": (16, 120, 160, 160, 160, 9, True, 2),
    "pyr160_gelu": (16, 120, 160, 160, 160, 9, False, 2),
    "last800_320": (16, 120, 160, 800, 320, 9, False, 0),
    "head64_64_elu": (16, 240, 320, 64, 64, 9, False, 3),
    "head64_32_elu": (16, 480, 640, 64, 32, 9, False, 3),
    "head32_32_elu": (16, 480, 640, 32, 32, 9, False, 3),
    "head32_1_sigmoid": (16, 480, 640, 32, 1, 9, False, 4),
    "head64_64_ln": (16, 240, 320, 64, 64, 9, True, 0),
    "ffn_256_2048": (1, 1, 4800, 256, 2048, 1, False, 1),
    "ffn_2048_256_ln": (1, 1, 4800, 2048, 256, 1, True, 0),
    "bb_64_256_relu": (1, 1, 307200, 64, 256, 1, False, 1),
    "bb_64_256_res": (1, 1, 307200, 64, 256, 1, "res", 1),
    "bb_256_1024_res": (1, 1, 19200, 256, 1024, 1, "res", 1),
    "bb_256_64_relu": (1, 1, 307200, 256, 64, 1, False, 1),
    "bb_512_128": (1, 1, 76800, 512, 128, 1, False, 1),
    "lin64_192": (1, 1, 16 * 414 * 49, 64, 192, 1, False, 0),
    "lin192_384": (1, 1, 16 * 414 * 49, 192, 384, 1, False, 0),
}


def run(name):
    B, H, W, C, N, taps, ln, act = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    xs = [torch.randn(B, H, W, C, device="cuda", generator=g).bfloat16() for _ in range(3)]
    if taps == 9:
        pw = ops.pack_conv3x3(torch.randn(N, C, 3, 3, device="cuda", generator=g) * (9 * C) ** -0.5, torch.zeros(N, device="cuda"))
    else:
        pw = ops.pack_linear(torch.randn(N, C, device="cuda", generator=g) * C ** -0.5, torch.zeros(N, device="cuda"))
    res = None
    if ln == "res":
        ln = False
        res = torch.randn(B, H, W, pw.n_pad, device="cuda", generator=g).bfloat16()
    lnp = (torch.ones(pw.n_pad, device="cuda"), torch.zeros(pw.n_pad, device="cuda")) if ln else None
    out = torch.empty(B, H, W, pw.n_pad, device="cuda", dtype=torch.bfloat16)
    for i in range(3):
        ops.conv_gemm(xs[i % 3], pw, ln=lnp, post_act=act, out=out, res=res, res_mode=1 if res is not None else 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for i in range(iters):
        ops.conv_gemm(xs[i % 3], pw, ln=lnp, post_act=act, out=out, res=res, res_mode=1 if res is not None else 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * B * H * W * C * N * taps
    print("%-18s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1000, flop / ms / 1e9), "%6.0f GB/s" % ((xs[0].numel() + out.numel()) * 2 / ms / 1e6))


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        run(n)
