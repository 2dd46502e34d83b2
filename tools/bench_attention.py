"""Micro-benchmark of gwd_attention (the tcgen05 DETR attention kernel, gwd_attn_tc.cu) at the model's sizes: encoder
self-attention at B=16 / L=300 (480x640) and at B=64 / L=1200 (960x1280, the corner SURVEY section 7 names for the tensor-pipe
measurement), decoder cross-attention (Lq=100).  CUDA events over rotating inputs; FLOPs = 4 * Lq * Lk * 256 per image.
usage: python tools/bench_attention.py [case ...]      (for ncu: -k regex:gwd_attention_tc_kernel)"""
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import gwdepth_b200  # noqa: F401,E402
from gwdepth_b200 import ops  # noqa: E402

CASES = {"enc_b16_l300": (16, 300, 300), "enc_b64_l1200": (64, 1200, 1200), "enc_b16_l1200": (16, 1200, 1200),
         "cross_b16_q100_l300": (16, 100, 300), "cross_b64_q100_l1200": (64, 100, 1200), "self_b16_q100": (16, 100, 100)}
E, NH, HD = 256, 8, 32


def run(name):
    B, Lq, Lk = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(0)
    sets = [(torch.randn(B * Lq, E, device="cuda", generator=g).bfloat16() * HD ** -0.25,
             torch.randn(B * Lk, E, device="cuda", generator=g).bfloat16() * HD ** -0.25,
             torch.randn(B * Lk, E, device="cuda", generator=g).bfloat16()) for _ in range(3)]
    o = torch.empty(B * Lq, E, device="cuda", dtype=torch.bfloat16)

    def call(i):
        q, k, v = sets[i % 3]
        ops.attention(q, k, v, o, items=B, heads=NH, Lq=Lq, Lk=Lk, hd=HD, q_strides=(Lq * E, E), k_strides=(Lk * E, E),
                      v_strides=(Lk * E, E), o_strides=(Lq * E, E), scale=1.0)
    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for i in range(iters):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 4.0 * B * Lq * Lk * E
    # parity of the last call against torch (fp32 soft-max on the same bf16 operands)
    q, k, v = sets[(iters - 1) % 3]
    h = lambda t, L: t.float().view(B, L, NH, HD).permute(0, 2, 1, 3)  # noqa: E731
    want = (torch.softmax(h(q, Lq) @ h(k, Lk).transpose(-1, -2), -1) @ h(v, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, E)
    err = float((o.float() - want).abs().max())
    print("%-22s %8.1f us  %7.1f TFLOP/s  (%.1f %% of 1393.9)   max abs err vs torch %.4f" % (name, ms * 1000, flop / ms / 1e9, flop / ms / 1e9 / 13.939, err))


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        run(n)
