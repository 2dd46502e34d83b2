// gwd_attn_tc.cu -- fused multi-head attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for the DETR
// encoder / decoder attention of the line branch (head_dim 32).  Two kernels: Lk <= 480 -- the whole key axis fits one TMEM
// pass (below) --, longer key axes -- key tiles + online soft-max (gwd_attention_flash_tc_kernel, further down).
//
//   CTA = one (batch item, head, 128-query tile), 5 warps:
//     warp 4 (one lane) : TMA loads of the Q tile, all K rows and all V rows of the (item, head); then
//                         S[128, Lk] = Q K^T        (tcgen05.mma, K-major Q/K, 64-byte swizzle, fp32 in TMEM)
//                         O[128, 32] = P V          (P from shared memory as the A operand, V as an MN-major B operand:
//                                                    V is stored [key][d], i.e. d-contiguous, so no transpose is needed)
//     warps 0..3        : softmax, one thread per query row = one TMEM lane: row max and exp straight out of TMEM,
//                         P written to shared memory as bf16 in the 128-byte-swizzled K-major UMMA layout;
//                         finally O / rowsum -> bf16 -> global
// Replaces the bmm / masked_fill / softmax / bmm chain of src/models/multi_head_attention.py:317-372 (the q scaling of
// :276 is folded into the projection weights; the head-averaged weights of :375-378 are dead and not produced).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kHD = 32;
constexpr int kQTile = 128;
constexpr int kMaxLk = 480;

struct TcAttnParams {
  int items, heads, Lq, Lk, Lk_pad;
  bf16* o;
  int64_t o_is, o_rs;
  const uint8_t* kpm;
  float scale;
  uint32_t tmem_cols;
  uint32_t idesc_s0, idesc_s1, idesc_o;   // S MMA (two N halves), O MMA
  int n0, n1;                             // N of the two S MMAs (n1 may be 0)
  int q_coff, k_coff, v_coff;             // channel offset of head 0 inside the q / k / v buffers
  const uint32_t* drop_seed;              // train-mode dropout of the probabilities (nullptr: none)
  uint32_t drop_site, drop_thr;
  float drop_scale;                       // 1 / (1 - p)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > (1u << 24)) __trap();   // a protocol bug must fault, never hang the box
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 columns without the wait (the caller issues tcgen05.wait::ld once per batch of loads)
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// UMMA shared-memory descriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version 1 @46 | layout @61 (2=SW128, 4=SW64)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(160, 1)
gwd_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, const __grid_constant__ TcAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Lk_pad = p.Lk_pad;
  uint8_t* sQ = smem;                                        // [128][64 B]            SW64
  uint8_t* sK = sQ + kQTile * 64;                            // [Lk_pad][64 B]         SW64
  uint8_t* sV = sK + ((Lk_pad * 64 + 1023) & ~1023);         // [Lk_pad][64 B]         SW64 (MN-major B operand)
  uint8_t* sP = sV + ((Lk_pad * 64 + 1023) & ~1023);         // [Lk_pad/64][128][128 B] SW128
  const int nchunk = (Lk_pad + 63) / 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + static_cast<size_t>(nchunk) * kQTile * 128);
  uint64_t* ld_bar = bars;          // TMA loads landed
  uint64_t* s_bar = bars + 1;       // S = QK^T complete
  uint64_t* p_bar = bars + 2;       // P written by the 128 softmax threads
  uint64_t* o_bar = bars + 3;       // O = PV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kQTile;

  if (threadIdx.x == 0) {
    mbar_init(ld_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, 128);
    mbar_init(o_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_o = tmem_base + Lk_pad;   // O accumulator after the S columns
  // launched programmatically (gwd_launch): the prologue above may have overlapped the tail of the kernel in front; global memory
  // is touched from here on.  The trigger comes after the TMEM allocation: a dependent that took the columns first would wait for
  // this kernel while holding them
  gwd_pdl_wait();
  gwd_pdl_trigger();

  if (warp == 4) {
    if (elect_one()) {
      // ---- loads
      const uint32_t bytes = kQTile * 64 + 2u * Lk_pad * 64;
      mbar_expect_tx(ld_bar, bytes);
      tma_load_2d(sQ, &map_q, ld_bar, p.q_coff + head * kHD, item * p.Lq + q0);
      for (int r0 = 0; r0 < Lk_pad; r0 += 160) {   // boxes of 160 rows (Lk_pad is a multiple of 32, <= 480)
        tma_load_2d(sK + r0 * 64, &map_k, ld_bar, p.k_coff + head * kHD, item * p.Lk + r0);
        tma_load_2d(sV + r0 * 64, &map_v, ld_bar, p.v_coff + head * kHD, item * p.Lk + r0);
      }
      mbar_wait(ld_bar, 0);
      fence_after();
      // ---- S = Q K^T : K-major operands, 64-byte rows, two K steps of 16
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
      for (int k = 0; k < 2; ++k) {
        umma(tmem_base, make_desc(qa + k * 32, 512, 4), make_desc(ka + k * 32, 512, 4), p.idesc_s0, k);
        if (p.n1 > 0)
          umma(tmem_base + p.n0, make_desc(qa + k * 32, 512, 4), make_desc(ka + p.n0 * 64 + k * 32, 512, 4), p.idesc_s1, k);
      }
      umma_commit(s_bar);
      // ---- O = P V : A = P (K-major, 128-byte swizzle, 64-key chunks), B = V (MN-major, 64-byte rows)
      mbar_wait(p_bar, 0);
      fence_after();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
      const int ksteps = Lk_pad / 16;
      for (int k = 0; k < ksteps; ++k) {
        uint64_t ad = make_desc(pa + (k >> 2) * (kQTile * 128) + (k & 3) * 32, 1024, 2);
        uint64_t bd = make_desc(va + k * (16 * 64), 512, 4);
        umma(tmem_o, ad, bd, p.idesc_o, k);
      }
      umma_commit(o_bar);
    }
  } else {
    // ---- softmax: thread <-> query row <-> TMEM lane
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint8_t* kp = p.kpm ? p.kpm + static_cast<int64_t>(item) * p.Lk : nullptr;
    mbar_wait(s_bar, 0);
    fence_after();
    // full 16-key chunks without key padding take the check-free path: 1 FMNMX per score in the first pass and
    // FFMA + MUFU.EX2 + FADD in the second (the per-element validity tests made this 4-warp soft-max the critical path)
    // (four independent partial maxima / sums: one serial chain over Lk keys costs ~4 clocks per key)
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int c = 0; c < Lk_pad; c += 16) {
      uint32_t r[16];
      tmem_ld16(t_row + c, r);
      if (kp == nullptr && c + 16 <= p.Lk) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          bool ok = (c + i < p.Lk) && !(kp && kp[c + i]);
          if (ok) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        }
      }
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
    const float l2e = 1.4426950408889634f;
    const float mxs = mx * l2e;
    const uint32_t drop_key = p.drop_seed ? gwd_drop_key(*p.drop_seed, p.drop_site, static_cast<uint32_t>(item * p.heads + head)) : 0u;
    for (int c = 0; c < Lk_pad; c += 16) {
      uint32_t r[16];
      tmem_ld16(t_row + c, r);
      float e[16];
      if (kp == nullptr && c + 16 <= p.Lk) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          e[i] = ex2_fast(fmaf(__uint_as_float(r[i]), l2e, -mxs));
          sum4[i & 3] += e[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          bool ok = (c + i < p.Lk) && !(kp && kp[c + i]);
          e[i] = ok ? ex2_fast(fmaf(__uint_as_float(r[i]), l2e, -mxs)) : 0.f;
          sum4[i & 3] += e[i];
        }
      }
      if (p.drop_seed) {        // the row sum keeps every key; only the P that meets V is dropped (and re-scaled)
        const uint32_t base = static_cast<uint32_t>(q0 + row) * static_cast<uint32_t>(p.Lk) + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = gwd_drop_keep(drop_key, base + i, p.drop_thr) ? e[i] * p.drop_scale : 0.f;
      }
      // 16 keys = two 16-byte units of the 64-key chunk (c / 64), row `row`, 128-byte swizzle
      uint8_t* chunk = sP + static_cast<size_t>(c >> 6) * (kQTile * 128) + row * 128;
      const int u0 = (c & 63) >> 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 w;
        w.x = gwd_pack_bf16x2(e[8 * h + 0], e[8 * h + 1]); w.y = gwd_pack_bf16x2(e[8 * h + 2], e[8 * h + 3]);
        w.z = gwd_pack_bf16x2(e[8 * h + 4], e[8 * h + 5]); w.w = gwd_pack_bf16x2(e[8 * h + 6], e[8 * h + 7]);
        *reinterpret_cast<uint4*>(chunk + (((u0 + h) ^ (row & 7)) << 4)) = w;
      }
    }
    const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
    // make the generic-proxy writes of P visible to the tensor core (async proxy), then signal
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    mbar_arrive(p_bar);
    mbar_wait(o_bar, 0);
    fence_after();
    const int qi = q0 + row;
    uint32_t r0[16], r1[16];
    tmem_ld16(t_row + Lk_pad, r0);
    tmem_ld16(t_row + Lk_pad + 16, r1);
    if (qi < p.Lq) {
      const float inv = 1.f / sum;
      bf16* orow = p.o + item * p.o_is + static_cast<int64_t>(qi) * p.o_rs + head * kHD;
      uint4 w[4];
      uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[i] = gwd_pack_bf16x2(__uint_as_float(r0[2 * i]) * inv, __uint_as_float(r0[2 * i + 1]) * inv);
#pragma unroll
      for (int i = 0; i < 8; ++i) wp[8 + i] = gwd_pack_bf16x2(__uint_as_float(r1[2 * i]) * inv, __uint_as_float(r1[2 * i + 1]) * inv);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(orow)[i] = w[i];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Long key axes (Lk > 480: the DETR encoder at 960x1280 has L = 1 200 tokens): the same CTA organisation with the keys
// walked in TILES of 128 and an ONLINE soft-max (running row max m and row sum l; flash-attention recurrence):
//     per tile t:  S_t = Q K_t^T (tcgen05, TMEM)  ->  m' = max(m, rowmax S_t), alpha = 2^((m - m') log2 e)
//                  P_t = 2^((S_t - m') log2 e) -> bf16 -> shared memory;  l = l alpha + rowsum P_t;  O = O alpha
//                  PV_t = P_t V_t (tcgen05, a fresh 32-column TMEM accumulator)  ->  O += PV_t
// O lives in REGISTERS (one query row = one thread = 32 fp32 values), so the rescale costs 32 FMULs per tile and TMEM never
// has to be read-modified-written.  K_t / V_t are double-buffered: the TMA of tile t+1 is issued before the soft-max of tile
// t.  Inside a CTA the steps of one tile are serial (MMA -> soft-max -> MMA); the kernel needs 72 KB of shared memory and 256
// TMEM columns, so up to THREE CTAs share an SM and one CTA's soft-max overlaps another's MMAs.  With head dim 32 the kernel
// is bound by the exponentials, not by the tensor pipe: a 128 x 128 tile needs 16 384 ex2 (MUFU: 16 per clock and SM = 1 024
// clocks) for 2 MFLOP of MMA (256 clocks at the dense bf16 rate), i.e. the tensor pipe cannot exceed ~25 % with MUFU exponentials.
// ------------------------------------------------------------------------------------------------------------------
#ifndef GWD_FLASH_KT
#define GWD_FLASH_KT 64
#endif
constexpr uint32_t kFlashTmem = GWD_FLASH_KT + 32 <= 128 ? 128u : 256u;
constexpr int kPChunks = (GWD_FLASH_KT + 63) / 64;      // 64-key chunks of the P tile in shared memory
constexpr int kKT = GWD_FLASH_KT;    // keys per tile: S (KT) + PV (32) TMEM columns.  Measured at B = 64, L = 1 200: 128 keys (2 CTAs per SM) 548 us,
                                     // 96 keys (3 CTAs) 417 us, 64 keys (4 CTAs) 368 us: the soft-max latency chains want co-resident CTAs

__global__ void __launch_bounds__(160, 4)
gwd_attention_flash_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                              const __grid_constant__ CUtensorMap map_v, const __grid_constant__ TcAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                        // [128][64 B]          SW64
  uint8_t* sK = sQ + kQTile * 64;                            // 2 x [128][64 B]      SW64
  uint8_t* sV = sK + 2 * kKT * 64;                           // 2 x [128][64 B]      SW64 (MN-major B operand)
  uint8_t* sP = sV + 2 * kKT * 64;                           // [2][128][128 B]      SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPChunks * kQTile * 128);
  uint64_t* ld_bar = bars;          // [2] K_t / V_t landed (Q rides on buffer 0's first phase)
  uint64_t* s_bar = bars + 2;       // S_t complete
  uint64_t* p_bar = bars + 3;       // P_t written by the 128 soft-max threads
  uint64_t* o_bar = bars + 4;       // PV_t complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * kQTile;
  const int T = (p.Lk + kKT - 1) / kKT;

  if (threadIdx.x == 0) {
    mbar_init(ld_bar, 1);
    mbar_init(ld_bar + 1, 1);
    mbar_init(s_bar, 1);
    mbar_init(p_bar, 128);
    mbar_init(o_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kFlashTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_s = *tmem_ptr;
  const uint32_t tmem_pv = tmem_s + kKT;
  gwd_pdl_wait();      // launched programmatically (gwd_launch): global memory is touched from here on
  gwd_pdl_trigger();   // (after the TMEM allocation: a dependent that got the columns first would wait for this kernel while holding them)

  if (warp == 4) {
    if (elect_one()) {
      const uint32_t qa = smem_u32(sQ), pa = smem_u32(sP);
      mbar_expect_tx(ld_bar, kQTile * 64 + 2u * kKT * 64);
      tma_load_2d(sQ, &map_q, ld_bar, head * kHD, item * p.Lq + q0);
      tma_load_2d(sK, &map_k, ld_bar, head * kHD, item * p.Lk);
      tma_load_2d(sV, &map_v, ld_bar, head * kHD, item * p.Lk);
      for (int t = 0; t < T; ++t) {
        const int b = t & 1;
        if (t + 1 < T) {          // prefetch tile t+1 into the other buffer: PV_{t-1} (its last reader) must have completed
          if (t >= 1) mbar_wait(o_bar, (t - 1) & 1);
          mbar_expect_tx(ld_bar + (b ^ 1), 2u * kKT * 64);
          tma_load_2d(sK + (b ^ 1) * kKT * 64, &map_k, ld_bar + (b ^ 1), head * kHD, item * p.Lk + (t + 1) * kKT);
          tma_load_2d(sV + (b ^ 1) * kKT * 64, &map_v, ld_bar + (b ^ 1), head * kHD, item * p.Lk + (t + 1) * kKT);
        }
        mbar_wait(ld_bar + b, (t >> 1) & 1);
        fence_after();
        const uint32_t ka = smem_u32(sK + b * kKT * 64), va = smem_u32(sV + b * kKT * 64);
        for (int k = 0; k < 2; ++k) umma(tmem_s, make_desc(qa + k * 32, 512, 4), make_desc(ka + k * 32, 512, 4), p.idesc_s0, k);
        umma_commit(s_bar);
        mbar_wait(p_bar, t & 1);
        fence_after();
        for (int k = 0; k < kKT / 16; ++k)
          umma(tmem_pv, make_desc(pa + (k >> 2) * (kQTile * 128) + (k & 3) * 32, 1024, 2), make_desc(va + k * (16 * 64), 512, 4),
               p.idesc_o, k);
        umma_commit(o_bar);
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t t_row = tmem_s + (static_cast<uint32_t>(warp * 32) << 16);
    const uint8_t* kp = p.kpm ? p.kpm + static_cast<int64_t>(item) * p.Lk : nullptr;
    const float l2e = 1.4426950408889634f;
    const uint32_t drop_key = p.drop_seed ? gwd_drop_key(*p.drop_seed, p.drop_site, static_cast<uint32_t>(item * p.heads + head)) : 0u;
    float m = -INFINITY, l = 0.f, O[kHD];
#pragma unroll
    for (int i = 0; i < kHD; ++i) O[i] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int k0 = t * kKT;
      const bool plain = kp == nullptr && k0 + kKT <= p.Lk;      // no masked key in this tile: check-free path
      mbar_wait(s_bar, t & 1);
      fence_after();
      // (four independent partial maxima / sums: a serial chain over the tile's keys would cost ~4 clocks per key)
      float mx4[4] = {m, m, m, m};
      for (int c = 0; c < kKT; c += 32) {
        uint32_t r[32];
        tmem_ld32_nowait(t_row + c, r);
        tmem_ld_wait();
        if (plain) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kk = k0 + c + i;
            if (kk < p.Lk && !(kp && kp[kk])) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float mxs = mx == -INFINITY ? 0.f : mx * l2e;          // every key so far masked: keep the exponent finite
      const float alpha = m == -INFINITY ? 0.f : ex2_fast(m * l2e - mxs);
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < kKT; c += 32) {
        uint32_t r[32];
        tmem_ld32_nowait(t_row + c, r);
        tmem_ld_wait();
        // 32 keys = four 16-byte units of the 64-key chunk (c / 64), row `row`, 128-byte swizzle
        uint8_t* chunk = sP + static_cast<size_t>(c >> 6) * (kQTile * 128) + row * 128;
        const int u0 = (c & 63) >> 3;
        float e[32];
        if (plain) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            e[i] = ex2_fast(fmaf(__uint_as_float(r[i]), l2e, -mxs));
            sum4[i & 3] += e[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kk = k0 + c + i;
            const bool ok = kk < p.Lk && !(kp && kp[kk]);
            e[i] = ok ? ex2_fast(fmaf(__uint_as_float(r[i]), l2e, -mxs)) : 0.f;
            sum4[i & 3] += e[i];
          }
        }
        if (p.drop_seed) {
          const uint32_t base = static_cast<uint32_t>(q0 + row) * static_cast<uint32_t>(p.Lk) + k0 + c;
#pragma unroll
          for (int i = 0; i < 32; ++i) e[i] = gwd_drop_keep(drop_key, base + i, p.drop_thr) ? e[i] * p.drop_scale : 0.f;
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint4 w;
          w.x = gwd_pack_bf16x2(e[8 * h + 0], e[8 * h + 1]); w.y = gwd_pack_bf16x2(e[8 * h + 2], e[8 * h + 3]);
          w.z = gwd_pack_bf16x2(e[8 * h + 4], e[8 * h + 5]); w.w = gwd_pack_bf16x2(e[8 * h + 6], e[8 * h + 7]);
          *reinterpret_cast<uint4*>(chunk + (((u0 + h) ^ (row & 7)) << 4)) = w;
        }
      }
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      l = fmaf(l, alpha, sum);
      m = mx;
#pragma unroll
      for (int i = 0; i < kHD; ++i) O[i] *= alpha;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fence_before();
      mbar_arrive(p_bar);
      mbar_wait(o_bar, t & 1);
      fence_after();
      uint32_t pv[32];
      tmem_ld32_nowait(t_row + kKT, pv);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < kHD; ++i) O[i] += __uint_as_float(pv[i]);
      fence_before();             // the next tile's MMAs overwrite these TMEM columns: order the loads ahead of them
    }
    const int qi = q0 + row;
    if (qi < p.Lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* orow = p.o + item * p.o_is + static_cast<int64_t>(qi) * p.o_rs + head * kHD;
      uint4 w[4];
      uint32_t* wp = reinterpret_cast<uint32_t*>(w);
#pragma unroll
      for (int i = 0; i < 16; ++i) wp[i] = gwd_pack_bf16x2(O[2 * i] * inv, O[2 * i + 1] * inv);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(orow)[i] = w[i];
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "r"(kFlashTmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2D map over the [rows, heads*32] head-channel slice of a bf16 matrix whose rows are row_stride elements apart (base
// points at the first head's channel 0), box = 32 channels x box_rows rows, 64-byte swizzle
int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t row_stride, int cols, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(row_stride) * 2};
  cuuint32_t box[2] = {kHD, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

uint32_t idesc_bf16(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}

}  // namespace

// returns 1 when the problem is not eligible for the tensor-core path (caller falls back to the CUDA-core kernel)
int gwd_attention_tc_try(const gwd_attn_desc* d, cudaStream_t stream) {
  if (d->hd != kHD || d->bias || d->mask || d->Lk < 16) return 1;
  // key tiles + online soft-max: always for long key axes; also for self-attention over a few hundred tokens, where its
  // four co-resident CTAs per SM beat the one-CTA-per-SM single pass (encoder, B = 16, L = 300: 19.3 vs 28.9 us), but not
  // for the decoder's 100 queries (one query tile per (image, head): 14.9 vs 12.7 us)
  static const int flash_min = [] { const char* e = getenv("GWD_FLASH_MIN_LK"); return e ? atoi(e) : 0; }();
  const bool flash = d->Lk > kMaxLk || (flash_min > 0 ? d->Lk >= flash_min : (d->Lq >= 256 && d->Lk >= 192));
  // tensors must be plain [items*L, row_stride] matrices (item stride = L * row stride) with 16-byte aligned rows
  if (d->q_item_stride != static_cast<int64_t>(d->Lq) * d->q_row_stride ||
      d->k_item_stride != static_cast<int64_t>(d->Lk) * d->k_row_stride ||
      d->v_item_stride != static_cast<int64_t>(d->Lk) * d->v_row_stride)
    return 1;
  if (d->q_row_stride % 8 || d->k_row_stride % 8 || d->v_row_stride % 8 || d->o_row_stride % 8 || d->o_item_stride % 8)
    return 1;
  if ((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) |
       reinterpret_cast<uintptr_t>(d->o)) & 15)
    return 1;
  if (!encode_fn()) return 1;
  TcAttnParams p;
  memset(&p, 0, sizeof(p));
  p.items = d->items; p.heads = d->heads; p.Lq = d->Lq; p.Lk = d->Lk;
  p.Lk_pad = (d->Lk + 31) & ~31;
  p.o = static_cast<bf16*>(d->o); p.o_is = d->o_item_stride; p.o_rs = d->o_row_stride;
  p.kpm = d->key_padding; p.scale = d->scale;
  if (d->scale != 1.0f) return 1;    // the model folds the q scaling into the projection weights
  if (d->dropout_seed != nullptr && d->dropout_p > 0.f) {
    if (static_cast<int64_t>(d->Lq) * d->Lk >= (1ll << 32)) return 1;
    p.drop_seed = d->dropout_seed; p.drop_site = d->dropout_site;
    p.drop_thr = static_cast<uint32_t>(static_cast<double>(d->dropout_p) * 4294967296.0);
    p.drop_scale = 1.f / (1.f - d->dropout_p);
  }
  // the channel slice may start anywhere inside the row: express it as a column offset of an aligned base
  // (pointers are already 16-byte aligned, so the maps can simply start at the slice)
  p.q_coff = p.k_coff = p.v_coff = 0;
  CUtensorMap mq, mk, mv;
  const int kbox = flash ? kKT : (160 < p.Lk_pad ? 160 : p.Lk_pad);
  if (!flash && p.Lk_pad % kbox) return 1;      // Lk_pad is 32..160, 320 or 480
  const int cols = d->heads * kHD;
  if (make_map(&mq, d->q, static_cast<int64_t>(d->items) * d->Lq, d->q_row_stride, cols, kQTile) ||
      make_map(&mk, d->k, static_cast<int64_t>(d->items) * d->Lk, d->k_row_stride, cols, kbox) ||
      make_map(&mv, d->v, static_cast<int64_t>(d->items) * d->Lk, d->v_row_stride, cols, kbox)) {
    gwd_set_error("gwd_attention: cuTensorMapEncodeTiled failed");
    return GWD_ERR_CUDA;
  }
  if (flash) {
    p.idesc_s0 = idesc_bf16(kKT, false);
    p.idesc_o = idesc_bf16(kHD, true);
    const size_t smem_f = 1024 + kQTile * 64 + 4 * static_cast<size_t>(kKT) * 64 + kPChunks * kQTile * 128 + 64;
    static bool configured_f = false;
    if (!configured_f) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_attention_flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_f)));
      configured_f = true;
    }
    dim3 gridf(static_cast<unsigned>(gwd_ceil_div(d->Lq, kQTile)), d->heads, d->items);
    GWD_CUDA(gwd_launch(gwd_attention_flash_tc_kernel, gridf, dim3(160), smem_f, stream, 1, mq, mk, mv, p));
    GWD_LAUNCHED();
    return GWD_OK;
  }
  if (p.Lk_pad <= 256) { p.n0 = p.Lk_pad; p.n1 = 0; }
  else { p.n0 = p.Lk_pad / 2; p.n1 = p.Lk_pad - p.n0; if (p.n0 % 16 || p.n1 % 16) return 1; }
  p.idesc_s0 = idesc_bf16(p.n0, false);
  p.idesc_s1 = p.n1 ? idesc_bf16(p.n1, false) : 0;
  p.idesc_o = idesc_bf16(kHD, true);
  uint32_t need = p.Lk_pad + kHD, tcols = 32;
  while (tcols < need) tcols <<= 1;
  p.tmem_cols = tcols;
  const int nchunk = (p.Lk_pad + 63) / 64;
  size_t smem = 1024 + kQTile * 64 + 2 * static_cast<size_t>((p.Lk_pad * 64 + 1023) & ~1023) +
                static_cast<size_t>(nchunk) * kQTile * 128 + 64;
  static bool configured = false;
  if (!configured) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(d->Lq, kQTile)), d->heads, d->items);
  GWD_CUDA(gwd_launch(gwd_attention_tc_kernel, grid, dim3(160), smem, stream, 1, mq, mk, mv, p));
  GWD_LAUNCHED();
  return GWD_OK;
}
