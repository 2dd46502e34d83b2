// gwd_gemm.cu -- persistent, warp-specialised implicit-GEMM for sm_100a.
//
//   D[128 pixels, Nt] (fp32, TMEM) += A_tap[128, BK] (bf16, smem, K-major, swizzled) * W_tap[Nt, BK]^T
//
// One CTA per SM, looping over output tiles.  Roles:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 4D for activations, 3D for packed weights)
//   warp 1      : MMA issuer    (tcgen05.mma.cta_group::1.kind::f16, one elected lane) + TMEM alloc
//   warps 2..9  : epilogue      (tcgen05.ld -> bias / act / residual / LayerNorm / act -> global)
// The fp32 accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main
// loop of tile i+1.
//
// CTA-pair variant (template flag CTA2, clusters of two CTAs): streamed-weight 3x3 convolutions run
// tcgen05.mma.cta_group::2 -- one M = 256 MMA over both SMs, each CTA holding its 128 rows of A and
// half of the weight tile -- see the wrappers below and the kernel's header comment.
//
// 3x3 convolutions (stride 1, zero pad 1) are tap-GEMMs: for each channel chunk and each dx the
// producer loads ONE activation box of (TH+2) x TW pixels (TMA out-of-bounds zero fill provides
// the padding) and the three dy taps are MMAs whose A descriptor start address is shifted by
// dy*TW rows -- a multiple of 8 rows, so the shared-memory swizzle phase is preserved.  The
// weights for (dx, dy=0..2) arrive as one 3D box.  Linear layers are the taps==1 special case
// of the same kernel (W = rows, H = B = 1).
//
// Reference call sites replaced: see include/gwd_b200.h (gwd_conv_gemm).
#include <cuda.h>
#include <stdlib.h>
#include "gwd_common.cuh"

namespace {

constexpr int kMaxEpilogueWarps = 16;   // 8 (two column groups) or 16 (four column groups) epilogue warps per CTA
constexpr int kNumThreads = (2 + kMaxEpilogueWarps) * 32;
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 8;
constexpr int kTileM = 128;

struct GemmParams {
  int B, H, W;
  int TW, TH, pad;
  int tiles_x, tiles_y, m_tiles;
  int n_tiles, Nt;
  int kchunks, BK, nsub, ndx;
  int x_coff;
  int w_img_stride;  // taps when the weights are per image ([B][taps][n_pad][cin]), else 0
  int stages;
  uint32_t a_stage_bytes, b_stage_bytes;  // 1024-aligned slot sizes
  uint32_t a_tx_bytes, b_tx_bytes;        // bytes TMA actually delivers per stage
  uint32_t a_off[9], b_off[9];            // byte offsets of the A / B operand of every MMA sub-step inside a stage
  uint32_t a_tab[36], b_tab[36];          // descriptor start-address increments (16-byte units) of every MMA of a stage
  int resident;                           // 1: all weights live in shared memory for the whole kernel
  uint32_t w_kc_bytes, w_total_bytes;     // resident weights: bytes per channel chunk / in total
  uint32_t layout_type;                   // UMMA smem descriptor layout: 2=SW128 4=SW64 6=SW32
  uint32_t sbo_a, sbo_b;                  // byte stride between 8-row groups of the A / B operand
  uint32_t idesc;
  uint32_t acc_stride;                    // TMEM columns between consecutive accumulators
  uint32_t tmem_cols;
  uint32_t stage_off;                     // TMAEP kernels: byte offset of the per-warp epilogue staging tiles
  int nacc;                               // TMEM accumulators in flight (2, 4 or 8; power of two)
  int tile_split;   // 1: each 4-warp epilogue group drains whole tiles (tile j -> group j % groups); 0: groups split columns
  int n, n_pad, store_n;
  const float* bias;
  const float* ln_g;
  const float* ln_b;
  float ln_eps;
  int pre_act, post_act;
  float out_scale;
  const __nv_bfloat16* res;
  int res_cstride, res_coff, res_mode;
  void* y;
  int y_cstride, y_coff, y_f32;
  __nv_bfloat16* y_raw;
  int yraw_cstride, yraw_coff;
  int ag_act, ag_from_input;      // res_mode == GWD_RES_MUL_ACTGRAD: activation whose derivative (at the saved tensor `res`) scales the result
  float ag_y_mul, ag_scale;
  int phase_n;      // > 0: fused nearest x2 up-sampling, output channels are 4 phase groups of phase_n channels
  int ln_groups;    // LayerNorm is applied per group of n / ln_groups channels (1 = whole row)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    // watchdog: a protocol bug must fault, never hang the box (try_wait itself sleeps in HW)
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- CTA-pair (cta_group::2) forms: the two CTAs of a cluster run ONE M = 256 MMA, each holding its own 128 rows of A and HALF
// of the B tile in shared memory, so every SM reads A + B/2 per MMA instead of A + B (the wide 3x3 convolutions are bound by
// shared-memory bandwidth, not by the tensor pipe).  Only the leader (cluster rank 0) issues MMAs; both CTAs issue TMA loads that
// complete on the LEADER's full barrier; MMA completion is multicast to both CTAs' empty / accumulator-full barriers.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta), as for the local arrive: the TMEM reads are ordered by tcgen05.fence::before_thread_sync; a
  // .release.cluster here compiles to MEMBAR.ALL.GPU, i.e. waits for the tile's global stores (16 % of the kernel's stall samples)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int N, bool CTA2, typename P>
__device__ __forceinline__ void issue_mmas(const P& p, uint32_t d_tmem, uint64_t adesc0, uint64_t bdesc0,
                                           uint32_t accumulate) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (CTA2) umma_bf16_2sm(d_tmem, adesc0 + p.a_tab[i], bdesc0 + p.b_tab[i], p.idesc, i > 0 ? 1u : accumulate);
    else umma_bf16(d_tmem, adesc0 + p.a_tab[i], bdesc0 + p.b_tab[i], p.idesc, i > 0 ? 1u : accumulate);
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, swizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 | [46,48) version=1 |
// [61,64) layout type
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}

struct TileCoord {
  int b, y0, x0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  int per_img = p.tiles_x * p.tiles_y;
  t.b = mt / per_img;
  int r = mt - t.b * per_img;
  int ty = r / p.tiles_x;
  int tx = r - ty * p.tiles_x;
  t.y0 = ty * p.TH;
  t.x0 = tx * p.TW;
  t.n0 = nt * p.Nt;
  return t;
}

// ---------------------------------------------------------------------------------------------
// epilogue helpers: one thread owns one output pixel (one TMEM lane)
// ---------------------------------------------------------------------------------------------
struct RowCtx {
  bool valid;
  int64_t pix;  // linear pixel index (b*H + y)*W + x
  int b, py, px;
};

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if (ACT == GWD_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == GWD_ACT_GELU) return gwd_gelu(v);
  // branch-free ELU: exp(min(v,0)) - 1 with the fast exponential (abs error < 1e-7, far below bf16 output rounding);
  // libm expm1f made the epilogue of the dense-head convs the bottleneck (profiles/README.md)
  if (ACT == GWD_ACT_ELU) return fmaxf(v, 0.f) + (__expf(fminf(v, 0.f)) - 1.f);
  if (ACT == GWD_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

// loads 16 accumulator columns starting at chunk c0 (relative to tile) and applies bias/pre_act/res(before)
__device__ __forceinline__ void add_bf16x8(float* v, const uint4& u) {
  float2 f0 = gwd_unpack_bf16x2(u.x), f1 = gwd_unpack_bf16x2(u.y), f2 = gwd_unpack_bf16x2(u.z), f3 = gwd_unpack_bf16x2(u.w);
  v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y;
  v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
}

// v *= act'(saved activation) * scale  (GWD_RES_MUL_ACTGRAD)
__device__ __forceinline__ void mul_actgrad_bf16x8(float* v, const uint4& u, const GemmParams& p) {
  const float2 f0 = gwd_unpack_bf16x2(u.x), f1 = gwd_unpack_bf16x2(u.y), f2 = gwd_unpack_bf16x2(u.z), f3 = gwd_unpack_bf16x2(u.w);
  const float y[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= gwd_act_factor(y[i] * p.ag_y_mul, p.ag_act, p.ag_from_input) * p.ag_scale;
}

// RES_BY_CALLER: the caller adds the (software-pipelined) residual itself
template <int PRE_ACT, bool RES_BY_CALLER = false>
__device__ __forceinline__ void load_chunk(const GemmParams& p, uint32_t taddr, int n_base, const RowCtx& rc,
                                           float (&v)[16]) {
  uint32_t r[16];
  tmem_ld16(taddr, r);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  if (p.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + n_base);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 b4 = __ldg(bp + i);
      v[4 * i + 0] += b4.x;
      v[4 * i + 1] += b4.y;
      v[4 * i + 2] += b4.z;
      v[4 * i + 3] += b4.w;
    }
  }
  if (PRE_ACT != GWD_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = act_fn<PRE_ACT>(v[i]);
  }
  if (!RES_BY_CALLER && p.res_mode == GWD_RES_BEFORE_NORM && rc.valid) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.res + rc.pix * p.res_cstride + p.res_coff + n_base);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 u = __ldg(rp + h);
      float2 f0 = gwd_unpack_bf16x2(u.x), f1 = gwd_unpack_bf16x2(u.y), f2 = gwd_unpack_bf16x2(u.z),
             f3 = gwd_unpack_bf16x2(u.w);
      v[8 * h + 0] += f0.x; v[8 * h + 1] += f0.y; v[8 * h + 2] += f1.x; v[8 * h + 3] += f1.y;
      v[8 * h + 4] += f2.x; v[8 * h + 5] += f2.y; v[8 * h + 6] += f3.x; v[8 * h + 7] += f3.y;
    }
  }
}

__device__ __forceinline__ void store_bf16_16(__nv_bfloat16* dst, const float (&v)[16], int n_base, int store_n) {
  // dst points at channel n_base of this pixel; channel counts are multiples of 8
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (n_base + 8 * h + 8 <= store_n) {
      uint4 u;
      u.x = gwd_pack_bf16x2(v[8 * h + 0], v[8 * h + 1]);
      u.y = gwd_pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
      u.z = gwd_pack_bf16x2(v[8 * h + 4], v[8 * h + 5]);
      u.w = gwd_pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
      *reinterpret_cast<uint4*>(dst + 8 * h) = u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// PRE_ACT / POST_ACT / HAS_LN are compile-time so that each instantiation carries one activation body instead of a
// 5-way switch unrolled 16x at three sites (that version was instruction-fetch bound: ~100 KB of SASS per kernel)
// TMAEP (plain [rows, N] Linears with 64 output columns per epilogue warp, bf16 out): the epilogue never touches global
// memory with the LSU.  Each epilogue warp owns a 32-row x 128-byte staging tile (SWIZZLE_128B): the residual tile
// arrives in it by TMA (requested before the accumulator is waited for), the thread-per-row math reads / overwrites it
// in place (conflict-free: chunk ^ (row & 7)), and the result leaves by TMA store.  Ragged row / column tails are
// clipped by the tensor maps.  A separate instantiation: the other epilogues do not carry this code (the kernel is
// sensitive to its instruction footprint).
// CTA2: the CTA-pair variant (see the cta_group::2 wrappers above).  The pair owns two consecutive M tiles of the same N tile (rank r
// = M tile 2q + r; an odd tail gives rank 1 a ghost tile: its loads are zero-filled out-of-bounds boxes and its stores are masked).
// ACTGRAD (training, data-gradient GEMMs): the saved activation tensor travels the residual path (software-pipelined loads, or the TMA
// staging tile) and the result is MULTIPLIED by the activation's derivative there -- the separate gwd_act_bwd pass over the gradient
// (read dy, read y, write) disappears.  Its own instantiations: the other epilogues do not carry the code.
template <int PRE_ACT, int POST_ACT, bool HAS_LN, bool TMAEP = false, bool CTA2 = false, bool ACTGRAD = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gwd_tapgemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap map_res,
                   const __grid_constant__ CUtensorMap map_y) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B; do not trust the declared alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + static_cast<size_t>(p.stages) * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      smem_b + (p.resident ? ((p.w_total_bytes + 1023u) & ~1023u) : static_cast<size_t>(p.stages) * p.b_stage_bytes));
  uint64_t* full_bar = bars;                      // [stages]
  uint64_t* empty_bar = bars + kMaxStages;        // [stages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;              // [nacc]
  uint64_t* tmem_empty = bars + 2 * kMaxStages + kMaxAcc;   // [nacc]
  uint64_t* w_bar = bars + 2 * kMaxStages + 2 * kMaxAcc;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAcc + 1);

  float2* ln_stats = reinterpret_cast<float2*>(tmem_ptr + 4);
  float* ln_gs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ln_stats + 2 * 4 * 128) + 15) & ~uintptr_t(15));   // HAS_LN: [256] scale, [256] shift (16-byte aligned)
  float* ln_bs = ln_gs + 256;
  uint64_t* ep_bar = reinterpret_cast<uint64_t*>(smem + p.stage_off);   // TMAEP: [16] one barrier per epilogue warp
  uint8_t* ep_stage = smem + p.stage_off + 1024;                        // TMAEP: [16][32 rows][128 B]   // [2][4 column groups][128 rows] partial (sum, sum of squares)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  // work items: tiles, or (M-tile pair, N tile) for the CTA pair
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const int q_first = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int q_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_q = CTA2 ? ((p.m_tiles + 1) >> 1) * p.n_tiles : total_tiles;
  auto tile_of = [&](int q) -> int {
    return CTA2 ? (2 * (q / p.n_tiles) + static_cast<int>(rank)) * p.n_tiles + q % p.n_tiles : q;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < p.nacc; ++a) {
      mbar_init(&tmem_full[a], 1);
      // CTA pair: the epilogue warps of BOTH CTAs release the leader's barrier
      mbar_init(&tmem_empty[a], (p.tile_split ? 4 : (blockDim.x >> 5) - 2) * (CTA2 ? 2 : 1));
    }
    mbar_init(w_bar, 1);
    if (TMAEP)
      for (int w = 0; w < kMaxEpilogueWarps; ++w) mbar_init(&ep_bar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    if (CTA2) {   // the same warp of both CTAs allocates the pair's columns
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // everything above ran on kernel parameters, shared memory and TMEM only: with a programmatic launch it overlapped the tail of
  // the kernel in front.  From here on global memory is touched (gwd_common.cuh, PDL)
  gwd_pdl_wait();
  if (HAS_LN) {
    const int cnt = p.phase_n > 0 ? p.phase_n : p.Nt;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      ln_gs[i] = __ldg(p.ln_g + i);
      ln_bs[i] = __ldg(p.ln_b + i);
    }
  }
  tcgen05_fence_before();
  if (CTA2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // only now may the next kernel's CTAs come aboard: this CTA (pair) holds its TMEM columns.  A dependent that became resident
  // earlier could take the columns first and then sit in its griddepcontrol.wait for THIS kernel to finish: a deadlock
  gwd_pdl_trigger();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.resident) {  // small filters: fetch every weight once, they stay in shared memory for all tiles of this CTA
        mbar_arrive_expect_tx(w_bar, p.w_total_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_3d(smem_b + static_cast<size_t>(kc) * p.w_kc_bytes, &map_b, w_bar, kc * p.BK, 0, 0);
      }
      for (int q = q_first; q < total_q; q += q_step) {
        TileCoord t = decode_tile(p, tile_of(q));
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int dx = 0; dx < p.ndx; ++dx) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (CTA2) {
              // both CTAs' boxes complete on the leader's barrier, which expects the bytes of both; each CTA fetches its own
              // 128 rows of A and its half of the B tile's rows
              const uint32_t full_addr = mapa_u32(smem_u32(&full_bar[stage]), 0u);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (p.a_tx_bytes + p.b_tx_bytes));
              tma_load_4d_2sm(smem_a + static_cast<size_t>(stage) * p.a_stage_bytes, &map_a, full_addr,
                              p.x_coff + kc * p.BK, t.x0 + dx - p.pad, t.y0 - p.pad, t.b);
              tma_load_3d_2sm(smem_b + static_cast<size_t>(stage) * p.b_stage_bytes, &map_b, full_addr,
                              kc * p.BK, t.n0 + static_cast<int>(rank) * (p.Nt >> 1), dx * p.nsub);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
              continue;
            }
            mbar_arrive_expect_tx(&full_bar[stage], p.a_tx_bytes + p.b_tx_bytes);
            tma_load_4d(smem_a + static_cast<size_t>(stage) * p.a_stage_bytes, &map_a, &full_bar[stage],
                        p.x_coff + kc * p.BK, t.x0 + dx - p.pad, t.y0 - p.pad, t.b);
            if (!p.resident)
              tma_load_3d(smem_b + static_cast<size_t>(stage) * p.b_stage_bytes, &map_b, &full_bar[stage],
                          kc * p.BK, t.n0, dx * p.nsub + t.b * p.w_img_stride);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    // elect.sync (not `lane == 0`): ptxas then knows exactly one thread runs this region and keeps the descriptors on
    // the uniform datapath; with a lane test every tcgen05.mma was wrapped in an R2UR waterfall loop (~100 clk each)
    if ((!CTA2 || rank == 0) && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int nmma = p.nsub * (p.BK / 16);
      if (p.resident) {
        mbar_wait(w_bar, 0);
        tcgen05_fence_after();
      }
      for (int q = q_first; q < total_q; q += q_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * p.acc_stride;
        uint32_t accumulate = 0;
        int kc = 0, dx = 0;
        for (int it = 0; it < p.kchunks * p.ndx; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_base = smem_u32(smem_a + static_cast<size_t>(stage) * p.a_stage_bytes);
          const uint32_t b_base = p.resident ? smem_u32(smem_b) + static_cast<uint32_t>(kc) * p.w_kc_bytes
                                             : smem_u32(smem_b) + static_cast<uint32_t>(stage) * p.b_stage_bytes;
          if (++dx == p.ndx) {
            dx = 0;
            ++kc;
          }
          // the single issuing thread is the critical path of small tiles: descriptors are one add per operand
          const uint64_t adesc0 = make_smem_desc(a_base, p.sbo_a, p.layout_type);
          const uint64_t bdesc0 = make_smem_desc(b_base, p.sbo_b, p.layout_type);
          // fully unrolled with constant indices: every table entry is a constant-bank operand that goes straight to
          // the uniform datapath (a table in shared memory cost a generic load + R2UR pair per MMA on this one thread,
          // which is the critical path of small-K tiles)
          // and branch-free: a branch between two MMAs makes ptxas re-materialise every uniform register
          switch (nmma) {
            case 1: issue_mmas<1, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 2: issue_mmas<2, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 3: issue_mmas<3, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 4: issue_mmas<4, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 6: issue_mmas<6, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 9: issue_mmas<9, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 12: issue_mmas<12, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            case 18: issue_mmas<18, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
            default: issue_mmas<36, CTA2>(p, d_tmem, adesc0, bdesc0, accumulate); break;
          }
          accumulate = 1;
          if (CTA2) umma_commit_2sm(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (CTA2) umma_commit_2sm(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == p.nacc) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lane quarter this warp may touch
    // 4 warps (one per TMEM lane quarter) form a group.  Column split: the groups share every tile, each storing its
    // own column range.  Tile split: group g drains tiles g, g + groups, ... of this CTA on its own accumulators, so
    // several tiles' epilogues (TMEM load -> math -> store latency chains) are in flight at once.
    const int wgroups = ((blockDim.x >> 5) - 2) >> 2;
    const int ngrp = p.tile_split ? 1 : wgroups;       // column groups
    const int half = p.tile_split ? 0 : ew >> 2;       // which column group of the tile this warp stores
    const int j_step = p.tile_split ? wgroups : 1;     // CTA-local tile counter stride
    const int acc_mask = p.nacc - 1, acc_shift = 31 - __clz(p.nacc);
    const int row = quad * 32 + lane;
    const int nchunks = p.Nt / 16;
    const int c_begin = (nchunks * half + ngrp - 1) / ngrp;
    const int c_end = (nchunks * (half + 1) + ngrp - 1) / ngrp;
    int j = p.tile_split ? ew >> 2 : 0;
    for (int q = q_first + j * q_step; q < total_q; q += j_step * q_step, j += j_step) {
      const int acc = j & acc_mask;
      const uint32_t acc_phase = (j >> acc_shift) & 1;
      TileCoord t = decode_tile(p, tile_of(q));
      RowCtx rc;
      {
        int py = t.y0 + row / p.TW;
        int px = t.x0 + row % p.TW;
        rc.valid = (py < p.H) && (px < p.W) && (!CTA2 || t.b < p.B);   // t.b >= B: the ghost tile of an odd pair tail
        rc.pix = (static_cast<int64_t>(t.b) * p.H + py) * p.W + px;
        rc.b = t.b; rc.py = py; rc.px = px;
      }
      // Plain epilogues: the residual of the first chunk is requested before the accumulator is waited for and the
      // residual of chunk c+1 while chunk c is processed (the profile showed the epilogue warps parked on one
      // dependent DRAM-latency load per 16-column chunk).
      uint8_t* my_stage = ep_stage + ew * 4096;
      uint8_t* my_row = my_stage + lane * 128;
      const int ep_col0 = t.n0 + c_begin * 16, ep_row0 = t.x0 + quad * 32;
      if (TMAEP) {
        if (lane == 0) {
          // the previous tile's store must have finished reading the staging tile before it is overwritten
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (p.res_mode != GWD_RES_NONE) {
            mbar_arrive_expect_tx(&ep_bar[ew], 4096);
            tma_load_2d(my_stage, &map_res, &ep_bar[ew], ep_col0, ep_row0);
          }
        }
        __syncwarp();
      }
      uint4 rpre0 = make_uint4(0u, 0u, 0u, 0u), rpre1 = rpre0;
      // (LayerNorm epilogues: only the residual that is added AFTER the norm; the one added before it enters the statistics pass
      // through load_chunk)
      const bool res_pf = !TMAEP && p.res_mode != GWD_RES_NONE && rc.valid && (!HAS_LN || p.res_mode == GWD_RES_AFTER);
      const __nv_bfloat16* res_row = p.res + rc.pix * p.res_cstride + p.res_coff + t.n0;
      if (res_pf && c_begin < c_end) {
        const bool after = p.res_mode != GWD_RES_BEFORE_NORM;
        const int nb = t.n0 + c_begin * 16;
        if (!after || nb + 8 <= p.store_n) rpre0 = __ldg(reinterpret_cast<const uint4*>(res_row + c_begin * 16));
        if (!after || nb + 16 <= p.store_n) rpre1 = __ldg(reinterpret_cast<const uint4*>(res_row + c_begin * 16) + 1);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + static_cast<uint32_t>(acc) * p.acc_stride +
                             (static_cast<uint32_t>(quad * 32) << 16);
      if (TMAEP && p.res_mode != GWD_RES_NONE) mbar_wait(&ep_bar[ew], static_cast<uint32_t>(j) & 1u);   // one phase per tile

      // A "segment" is a chunk range whose LayerNorm statistics are taken together.  Whole-row LayerNorm: both column
      // halves read the full row for the statistics and store their own half.  Grouped LayerNorm (fused up-sampling:
      // one group per output phase): each half owns whole groups.
      const int groups = HAS_LN ? p.ln_groups : 1;
      const int cpg = nchunks / groups;
      const int seg_first = groups > 1 ? half * (groups / ngrp) : 0;
      const int seg_last = groups > 1 ? (half + 1) * (groups / ngrp) : 1;
      const int n_group = p.n / groups;
      for (int seg = seg_first; seg < seg_last; ++seg) {
        const int st_begin = groups > 1 ? seg * cpg : 0;
        const int st_end = groups > 1 ? (seg + 1) * cpg : nchunks;
        const int my_begin = groups > 1 ? st_begin : c_begin;
        const int my_end = groups > 1 ? st_end : c_end;
        float mean = 0.f, rstd = 1.f;
        if (HAS_LN) {
          // whole-row LayerNorm under column split: every group sums its own columns, the partial sums of a row meet
          // in shared memory (one named barrier per TMEM lane quarter), so no group re-reads the whole row
          const bool exchange = groups == 1 && ngrp > 1;
          float s = 0.f, ss = 0.f;
          for (int c = exchange ? my_begin : st_begin; c < (exchange ? my_end : st_end); ++c) {
            float v[16];
            load_chunk<PRE_ACT>(p, t_row + c * 16, t.n0 + c * 16, rc, v);
            if (groups > 1 || t.n0 + c * 16 + 16 <= p.n) {   // only the last chunk of a padded row needs the column test
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                s += v[i];
                ss = fmaf(v[i], v[i], ss);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (t.n0 + c * 16 + i < p.n) {
                  s += v[i];
                  ss = fmaf(v[i], v[i], ss);
                }
              }
            }
          }
          if (exchange) {
            float2* slot = ln_stats + (j & 1) * 512 + row;   // double buffered by tile parity
            slot[half * 128] = make_float2(s, ss);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "r"(ngrp * 32) : "memory");
            s = 0.f;
            ss = 0.f;
            for (int g = 0; g < ngrp; ++g) {
              const float2 part = slot[g * 128];
              s += part.x;
              ss += part.y;
            }
          }
          mean = s / static_cast<float>(n_group);
          float var = fmaxf(ss / static_cast<float>(n_group) - mean * mean, 0.f);
          rstd = rsqrtf(var + p.ln_eps);
        }
        for (int c = my_begin; c < my_end; ++c) {
          const int n_base = t.n0 + c * 16;
          float v[16];
          load_chunk<PRE_ACT, !HAS_LN>(p, t_row + c * 16, n_base, rc, v);
          const uint4 rcur0 = rpre0, rcur1 = rpre1;
          if (res_pf && c + 1 < my_end) {
            const bool after = p.res_mode != GWD_RES_BEFORE_NORM;
            rpre0 = rpre1 = make_uint4(0u, 0u, 0u, 0u);
            if (!after || n_base + 24 <= p.store_n) rpre0 = __ldg(reinterpret_cast<const uint4*>(res_row + (c + 1) * 16));
            if (!after || n_base + 32 <= p.store_n) rpre1 = __ldg(reinterpret_cast<const uint4*>(res_row + (c + 1) * 16) + 1);
          }
          if (!HAS_LN && res_pf && p.res_mode == GWD_RES_BEFORE_NORM) {
            add_bf16x8(v, rcur0);
            add_bf16x8(v + 8, rcur1);
          }
          uint4* ep_p0 = reinterpret_cast<uint4*>(my_row + ((((c - c_begin) * 2) ^ (lane & 7)) << 4));
          uint4* ep_p1 = reinterpret_cast<uint4*>(my_row + ((((c - c_begin) * 2 + 1) ^ (lane & 7)) << 4));
          if (TMAEP && p.res_mode == GWD_RES_BEFORE_NORM) {
            add_bf16x8(v, *ep_p0);
            add_bf16x8(v + 8, *ep_p1);
          }
          if (p.y_raw != nullptr && rc.valid) {
            store_bf16_16(p.y_raw + rc.pix * p.yraw_cstride + p.yraw_coff + n_base, v, n_base, p.store_n);
          }
          // output location: plain, or one of the four x2 up-sampling phases (channel group -> (oy, ox))
          int64_t opix = rc.pix;
          int och = n_base, olimit = p.store_n;
          if (p.phase_n > 0) {
            const int phase = n_base / p.phase_n;
            och = n_base - phase * p.phase_n;
            olimit = p.phase_n;
            opix = (static_cast<int64_t>(rc.b) * (2 * p.H) + 2 * rc.py + (phase >> 1)) * (2 * p.W) + 2 * rc.px + (phase & 1);
          }
          if (HAS_LN) {
            const float shift = -mean * rstd;   // (v - mean) * rstd * g + b as two FMAs
            // scale / shift from the shared-memory copy made at kernel start (broadcast reads; the global loads here were the top
            // stall of the LayerNorm epilogues)
            const float4* gp = reinterpret_cast<const float4*>(ln_gs + och);
            const float4* bp = reinterpret_cast<const float4*>(ln_bs + och);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 g4 = gp[i], b4 = bp[i];
              v[4 * i + 0] = fmaf(fmaf(v[4 * i + 0], rstd, shift), g4.x, b4.x);
              v[4 * i + 1] = fmaf(fmaf(v[4 * i + 1], rstd, shift), g4.y, b4.y);
              v[4 * i + 2] = fmaf(fmaf(v[4 * i + 2], rstd, shift), g4.z, b4.z);
              v[4 * i + 3] = fmaf(fmaf(v[4 * i + 3], rstd, shift), g4.w, b4.w);
            }
          }
          if (POST_ACT != GWD_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = act_fn<POST_ACT>(v[i]);
          }
          if (p.out_scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= p.out_scale;
          }
          if (!HAS_LN) {
            if (ACTGRAD) {
              if (TMAEP) {
                mul_actgrad_bf16x8(v, *ep_p0, p);
                mul_actgrad_bf16x8(v + 8, *ep_p1, p);
              } else if (res_pf) {
                mul_actgrad_bf16x8(v, rcur0, p);
                mul_actgrad_bf16x8(v + 8, rcur1, p);
              }
            } else {
              if (res_pf && p.res_mode == GWD_RES_AFTER) {
                add_bf16x8(v, rcur0);
                add_bf16x8(v + 8, rcur1);
              }
              if (TMAEP && p.res_mode == GWD_RES_AFTER) {
                add_bf16x8(v, *ep_p0);
                add_bf16x8(v + 8, *ep_p1);
              }
            }
          } else if (res_pf) {   // LayerNorm, residual added after it: the pieces were prefetched (zeros beyond store_n)
            add_bf16x8(v, rcur0);
            add_bf16x8(v + 8, rcur1);
          }
          // channels beyond the logical width are padding: keep them exactly zero
          if (n_base + 16 > p.n) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n_base + i >= p.n) v[i] = 0.f;
          }
          if (TMAEP) {   // the result overwrites the residual pieces in place
            uint4 u0, u1;
            u0.x = gwd_pack_bf16x2(v[0], v[1]); u0.y = gwd_pack_bf16x2(v[2], v[3]);
            u0.z = gwd_pack_bf16x2(v[4], v[5]); u0.w = gwd_pack_bf16x2(v[6], v[7]);
            u1.x = gwd_pack_bf16x2(v[8], v[9]); u1.y = gwd_pack_bf16x2(v[10], v[11]);
            u1.z = gwd_pack_bf16x2(v[12], v[13]); u1.w = gwd_pack_bf16x2(v[14], v[15]);
            *ep_p0 = u0;
            *ep_p1 = u1;
          } else if (rc.valid) {
            if (p.y_f32) {
              float* dst = reinterpret_cast<float*>(p.y) + opix * p.y_cstride + p.y_coff + och;
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (och + i < olimit) dst[i] = v[i];
            } else {
              store_bf16_16(reinterpret_cast<__nv_bfloat16*>(p.y) + opix * p.y_cstride + p.y_coff + och, v, och, olimit);
            }
          }
        }
      }
      // release the accumulator (all TMEM reads of this tile are done)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0u));
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (TMAEP) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_y, my_stage, ep_col0, ep_row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (TMAEP && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
  }

  tcgen05_fence_before();
  if (CTA2) cluster_sync_all();   // neither CTA leaves (or frees the pair's columns) while the other still works
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                   : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

uint32_t pow2_at_least(uint32_t v, uint32_t lo) {
  uint32_t r = lo;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace

extern "C" int gwd_conv_gemm(const gwd_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(d != nullptr && d->x && d->w && d->y, "gwd_conv_gemm: null pointer");
  GWD_CHECK_ARG(d->taps == 1 || d->taps == 9, "gwd_conv_gemm: taps must be 1 or 9 (got %d)", d->taps);
  GWD_CHECK_ARG((d->x_wstride == 0 && d->x_hstride == 0 && d->x_bstride == 0) ||
                    (d->taps == 1 && d->x_wstride > 0 && d->x_hstride > 0 && d->x_bstride > 0),
                "gwd_conv_gemm: strided pixel views need taps == 1 and three positive strides");
  GWD_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0, "gwd_conv_gemm: empty input");
  GWD_CHECK_ARG(d->cin > 0 && d->cin % 16 == 0, "gwd_conv_gemm: cin %% 16 != 0 (%d)", d->cin);
  GWD_CHECK_ARG(d->n_pad > 0 && d->n_pad % 16 == 0 && d->n > 0 && d->n <= d->n_pad, "gwd_conv_gemm: bad n/n_pad");
  GWD_CHECK_ARG(d->x_cstride % 8 == 0 && d->x_coff % 8 == 0 && d->x_coff + d->cin <= d->x_cstride,
                "gwd_conv_gemm: bad x channel slice");
  GWD_CHECK_ARG((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
                "gwd_conv_gemm: x/w must be 16-byte aligned");
  int store_n = d->store_n > 0 ? d->store_n : d->n;
  GWD_CHECK_ARG(store_n <= d->n_pad, "gwd_conv_gemm: store_n > n_pad");
  if (!d->y_f32)
    GWD_CHECK_ARG(d->y_cstride % 8 == 0 && d->y_coff % 8 == 0 && store_n % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(d->y) & 15) == 0,
                  "gwd_conv_gemm: bf16 output needs 8-channel alignment");
  if (d->res_mode != GWD_RES_NONE)
    GWD_CHECK_ARG(d->res && d->res_cstride % 8 == 0 && d->res_coff % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(d->res) & 15) == 0 && store_n % 8 == 0,
                  "gwd_conv_gemm: bad residual");
  if (d->y_raw)
    GWD_CHECK_ARG(d->yraw_cstride % 8 == 0 && d->yraw_coff % 8 == 0 && store_n % 8 == 0, "gwd_conv_gemm: bad y_raw");
  const bool actgrad = d->res_mode == GWD_RES_MUL_ACTGRAD;
  if (actgrad)
    GWD_CHECK_ARG(d->ln_g == nullptr && d->pre_act == GWD_ACT_NONE && d->post_act == GWD_ACT_NONE && !d->y_f32 && d->y_raw == nullptr &&
                      !d->upsample2 && !d->w_per_image && (d->out_scale == 1.f || d->out_scale == 0.f) &&
                      (d->ag_act == GWD_ACT_RELU || d->ag_act == GWD_ACT_ELU || d->ag_act == GWD_ACT_SIGMOID ||
                       (d->ag_act == GWD_ACT_GELU && d->ag_from_input)),
                  "gwd_conv_gemm: GWD_RES_MUL_ACTGRAD needs a plain bf16 epilogue and ReLU / ELU / sigmoid (or GELU from its input)");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W;
  const bool conv = d->taps == 9;
  p.pad = conv ? 1 : 0;
  p.nsub = conv ? 3 : 1;
  p.ndx = conv ? 3 : 1;
  // N tiling
  int n_tiles = 1;
  while (!(d->n_pad % n_tiles == 0 && (d->n_pad / n_tiles) % 16 == 0 && d->n_pad / n_tiles <= 256)) {
    ++n_tiles;
    GWD_CHECK_ARG(n_tiles <= d->n_pad, "gwd_conv_gemm: cannot tile n_pad=%d", d->n_pad);
  }
  // Small problems (a few M tiles, e.g. the 1 600- and 4 800-row Linears of the DETR transformer): every CTA streams its
  // whole [Nt x K] weight slab through one SM (~100 GB/s), so narrower N tiles on more SMs cut the time almost
  // linearly.  Split until about half the SMs have a tile (LayerNorm / fused up-sampling epilogues need the whole row).
  if (d->ln_g == nullptr && !d->upsample2 && !d->w_per_image && d->taps == 1) {
    const int64_t m_tiles_est = gwd_ceil_div(static_cast<int64_t>(d->B) * d->H * d->W, kTileM);
    static const bool split_enabled = []() { const char* e = getenv("GWD_GEMM_NSPLIT"); return !(e && e[0] == '0'); }();
    while (split_enabled && m_tiles_est * n_tiles * 2 <= gwd_num_sms() && (d->n_pad / n_tiles) % 2 == 0 &&
           ((d->n_pad / n_tiles) / 2) % 16 == 0 && (d->n_pad / n_tiles) / 2 >= 64)
      n_tiles *= 2;
  }
  p.n_tiles = n_tiles;
  p.Nt = d->n_pad / n_tiles;
  GWD_CHECK_ARG(d->ln_g == nullptr || n_tiles == 1, "gwd_conv_gemm: LayerNorm epilogue needs n_pad <= 256");
  if (d->upsample2)
    GWD_CHECK_ARG(d->taps == 9 && n_tiles == 1 && d->n == d->n_pad && (d->n_pad / 4) % 16 == 0 && d->res_mode == GWD_RES_NONE &&
                      d->y_raw == nullptr && (p.Nt / 16) % 4 == 0,
                  "gwd_conv_gemm: fused up-sampling needs a 3x3 filter with 4 x (multiple of 16) unpadded channels, no residual");
  // Resident-weight mode: when the whole packed filter fits in ~96 KB of shared memory it is fetched once per CTA
  // and only activations stream through the TMA ring.  For 3x3 convs this mode also loads ONE (TH+2)x(TW+2) halo box
  // per channel chunk and forms all nine taps from it: tap (dy,dx) is the same box read at a start address shifted
  // by (dy*(TW+2)+dx) pixel rows, with TW = 8 so that each 8-row MMA group is one image row and the group stride
  // (SBO) is uniform.  (The 128/64/32-byte swizzle is a function of the absolute shared-memory address, as the
  // K-advance of +32 B inside a row already relies on, so a start address that is not 1024-byte aligned is fine.)
  const uint64_t w_bytes = static_cast<uint64_t>(d->taps) * p.Nt * d->cin * 2;
  p.resident = (!d->w_per_image && n_tiles == 1 && w_bytes <= 96 * 1024 && (!conv || d->H >= 12)) ? 1 : 0;
  int box_w, box_h;
  // M tiling: TW x TH = 128 pixels, TW a multiple of 8
  if (!conv) {
    p.TW = 128; p.TH = 1;
    box_w = 128; box_h = 1;
  } else if (p.resident) {
    p.TW = 8; p.TH = 16;
    box_w = p.TW + 2; box_h = p.TH + 2;
    p.nsub = 9; p.ndx = 1;
  } else {
    const int cand[5][2] = {{16, 8}, {8, 16}, {32, 4}, {64, 2}, {128, 1}};
    int64_t best = -1;
    for (int i = 0; i < 5; ++i) {
      // padded pixels covered by the tiling, plus the halo rows every tile re-reads
      int64_t cover = gwd_ceil_div(d->W, cand[i][0]) * cand[i][0] * gwd_ceil_div(d->H, cand[i][1]) * (cand[i][1] + 2);
      if (best < 0 || cover < best) {
        best = cover; p.TW = cand[i][0]; p.TH = cand[i][1];
      }
    }
    box_w = p.TW; box_h = p.TH + 2;
  }
  p.tiles_x = static_cast<int>(gwd_ceil_div(d->W, p.TW));
  p.tiles_y = static_cast<int>(gwd_ceil_div(d->H, p.TH));
  p.m_tiles = p.tiles_x * p.tiles_y * d->B;
  // K chunking
  p.BK = (d->cin % 64 == 0) ? 64 : (d->cin % 32 == 0) ? 32 : 16;
  const int a_rows = box_w * box_h;
  auto round1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  // Two co-resident CTAs per SM for small filters: the producer thread, the MMA-issuing thread and each epilogue warp
  // are serial latency chains (about 0.4-0.9 us per tile each), so a second independent CTA on the SM nearly doubles
  // the tile rate of the small-K convolutions whose tiles carry little tensor work.  Needs: resident weights, half the
  // shared memory, at most 256 TMEM columns, 8 epilogue warps (96 registers x 320 threads x 2 fits the register file).
  static const int forced_ctas = []() { const char* e = getenv("GWD_GEMM_CTAS"); return e ? atoi(e) : 0; }();
  int ctas = 1;
  if (p.resident && forced_ctas != 1 && pow2_at_least(static_cast<uint32_t>(p.Nt), 32) <= 128 &&
      p.m_tiles >= 4 * gwd_num_sms()) {
    // at least three ring slots per CTA; a 64-channel chunk may be halved to get there
    for (int bk = p.BK; bk >= 32 && ctas == 1; bk >>= 1) {
      if (round1k(static_cast<uint32_t>(w_bytes)) + 3 * round1k(static_cast<uint32_t>(a_rows) * bk * 2) <= 104 * 1024) {
        ctas = 2;
        p.BK = bk;
      }
    }
  }
  // CTA pair (cta_group::2, see the kernel): streamed-weight 3x3 convolutions with enough M tiles.  Per stage an SM then holds its
  // own activation box and HALF of the weight box, and every MMA reads A + B/2 from shared memory instead of A + B: at N = 160,
  // K chunk 32 that is 133 instead of 195 bytes per clock of shared-memory traffic (TMA writes + MMA reads) against 128 B/clk.
  static const bool cta2_enabled = []() { const char* e = getenv("GWD_GEMM_CTA2"); return !(e && e[0] == '0'); }();
  // (measured on 16 x 120 x 160 maps: 800->320 1 217 -> 1 014 us = 1.40 PFLOP/s; 160->160 unchanged; 64->128 and 64->256, whose tiles
  // carry few MMAs per epilogue, 20-25 % SLOWER: the pair couples the two CTAs' epilogues) -> long reductions only
  static const int cta2_min_k = []() { const char* e = getenv("GWD_GEMM_CTA2_MIN_K"); return e ? atoi(e) : 0; }();
  const bool cta2 = cta2_enabled && conv && !p.resident && !d->w_per_image && ctas == 1 && p.m_tiles >= 2 * gwd_num_sms() &&
                    (p.Nt / 2) % 8 == 0 && p.Nt >= 64 && d->taps * d->cin >= cta2_min_k;
  const int Nb = cta2 ? p.Nt / 2 : p.Nt;      // rows of the B tile one CTA holds
  const uint32_t budget = ctas == 2 ? 104 * 1024 : 200 * 1024;
  uint32_t w_region = 0;
  if (p.resident) {
    p.a_tx_bytes = static_cast<uint32_t>(a_rows) * p.BK * 2;
    p.b_tx_bytes = 0;
    p.a_stage_bytes = round1k(p.a_tx_bytes);
    p.b_stage_bytes = 0;
    p.w_kc_bytes = static_cast<uint32_t>(d->taps) * p.Nt * p.BK * 2;
    p.w_total_bytes = static_cast<uint32_t>(w_bytes);
    w_region = round1k(p.w_total_bytes);
    p.kchunks = d->cin / p.BK;
    p.stages = static_cast<int>((budget - w_region) / p.a_stage_bytes);
  } else {
    while (true) {
      p.a_tx_bytes = static_cast<uint32_t>(a_rows) * p.BK * 2;
      p.b_tx_bytes = static_cast<uint32_t>(p.nsub) * Nb * p.BK * 2;
      p.a_stage_bytes = round1k(p.a_tx_bytes);
      p.b_stage_bytes = round1k(p.b_tx_bytes);
      if ((p.a_stage_bytes + p.b_stage_bytes) * 3 <= budget || p.BK == 16) break;
      p.BK /= 2;  // keep at least 3 stages in flight
    }
    p.kchunks = d->cin / p.BK;
    p.stages = static_cast<int>(budget / (p.a_stage_bytes + p.b_stage_bytes));
  }
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  GWD_CHECK_ARG(p.stages >= 2, "gwd_conv_gemm: tile does not fit shared memory");
  const uint32_t row_bytes = p.BK * 2;
  p.layout_type = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  p.sbo_b = 8 * row_bytes;
  p.sbo_a = (conv && p.resident) ? static_cast<uint32_t>(box_w) * row_bytes : 8 * row_bytes;
  for (int sub = 0; sub < p.nsub; ++sub) {
    if (conv && p.resident) {   // sub = dx*3 + dy (the packed tap order)
      int dx = sub / 3, dy = sub % 3;
      p.a_off[sub] = static_cast<uint32_t>(dy * box_w + dx) * row_bytes;
    } else {
      p.a_off[sub] = static_cast<uint32_t>(sub) * p.TW * row_bytes;   // dy shift of TW rows
    }
    p.b_off[sub] = static_cast<uint32_t>(sub) * Nb * row_bytes;
  }
  {
    const int ks = p.BK / 16;
    for (int i = 0; i < p.nsub * ks; ++i) {
      p.a_tab[i] = (p.a_off[i / ks] + (i % ks) * 32) >> 4;
      p.b_tab[i] = (p.b_off[i / ks] + (i % ks) * 32) >> 4;
    }
  }
  // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both, N>>3 @17, M>>4 @24
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.Nt >> 3) << 17) |
            (static_cast<uint32_t>((cta2 ? 2 * kTileM : kTileM) >> 4) << 24);
  // Epilogue organisation.  Column split (the 4-warp groups share each tile): lowest latency for one tile, right for
  // wide plain tiles and for launches with about one tile per CTA.  Tile split (each group drains whole tiles on its
  // own TMEM accumulators): narrow tiles (no columns to share) and LayerNorm epilogues (whose statistics pass would be
  // repeated by every column group) when the CTA has a queue of tiles to overlap.
  // (staging the bf16 tile in shared memory for row-contiguous stores was measured: -13% on 192-column Linears, but
  // +20..40% on narrow / LayerNorm tiles and -2% on the whole step, so stores stay thread-per-row)
  static const int forced_warps = []() { const char* e = getenv("GWD_GEMM_EPI_WARPS"); return e ? atoi(e) : 0; }();
  static const int forced_split = []() { const char* e = getenv("GWD_GEMM_TILE_SPLIT"); return e ? atoi(e) : -1; }();
  const bool has_ln = d->ln_g != nullptr;
  const int64_t tiles_all = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
  p.acc_stride = pow2_at_least(static_cast<uint32_t>(p.Nt), 32);
  const int acc_fit = static_cast<int>(512u / p.acc_stride);   // accumulators that fit the 512 TMEM columns
  p.tile_split = ((has_ln && p.Nt <= 128) || p.Nt < 64) && tiles_all >= 3 * static_cast<int64_t>(gwd_num_sms()) ? 1 : 0;
  if (forced_split >= 0) p.tile_split = forced_split;
  int epi_warps;
  if (ctas == 2) {
    const int fit = static_cast<int>(256u / p.acc_stride);
    p.nacc = fit >= 4 ? 4 : 2;
    epi_warps = 8;
    p.tile_split = forced_split >= 0 ? forced_split : ((has_ln || p.Nt < 64) ? 1 : 0);
  } else if (p.tile_split) {
    p.nacc = acc_fit >= 8 ? 8 : acc_fit >= 4 ? 4 : 2;
    epi_warps = p.nacc >= 4 ? 16 : 8;
    if (forced_warps == 8) epi_warps = 8;
  } else {
    p.nacc = 2;
    epi_warps = p.Nt >= 64 ? 16 : 8;
    if (forced_warps == 8 || (forced_warps == 16 && epi_warps == 16)) epi_warps = forced_warps;
  }
  p.tmem_cols = pow2_at_least(static_cast<uint32_t>(p.nacc) * p.acc_stride, 32);
  int threads = (2 + epi_warps) * 32;
  p.x_coff = d->x_coff;
  p.w_img_stride = d->w_per_image ? d->taps : 0;
  p.n = d->n; p.n_pad = d->n_pad; p.store_n = store_n;
  p.bias = d->bias; p.ln_g = d->ln_g; p.ln_b = d->ln_b; p.ln_eps = d->ln_eps;
  p.pre_act = d->pre_act; p.post_act = d->post_act; p.out_scale = d->out_scale;
  p.res = static_cast<const __nv_bfloat16*>(d->res);
  p.res_cstride = d->res_cstride; p.res_coff = d->res_coff; p.res_mode = d->res_mode;
  p.y = d->y; p.y_cstride = d->y_cstride; p.y_coff = d->y_coff; p.y_f32 = d->y_f32;
  p.y_raw = static_cast<__nv_bfloat16*>(d->y_raw);
  p.yraw_cstride = d->yraw_cstride; p.yraw_coff = d->yraw_coff;
  p.ag_act = d->ag_act; p.ag_from_input = d->ag_from_input;
  p.ag_y_mul = d->ag_y_mul != 0.f ? d->ag_y_mul : 1.f;
  p.ag_scale = d->ag_scale != 0.f ? d->ag_scale : 1.f;
  p.phase_n = d->upsample2 ? d->n_pad / 4 : 0;
  p.ln_groups = d->upsample2 ? 4 : 1;

  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) {
    gwd_set_error("cuTensorMapEncodeTiled entry point not available");
    return GWD_ERR_CUDA;
  }
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                   : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap map_a, map_b;
  {
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(d->x_cstride), static_cast<cuuint64_t>(d->W),
                          static_cast<cuuint64_t>(d->H), static_cast<cuuint64_t>(d->B)};
    const bool strided = d->x_wstride != 0 || d->x_hstride != 0 || d->x_bstride != 0;
    const cuuint64_t ws = strided ? d->x_wstride : 1, hs = strided ? d->x_hstride : d->W,
                     bs = strided ? d->x_bstride : static_cast<cuuint64_t>(d->H) * d->W;
    cuuint64_t gstr[3] = {ws * d->x_cstride * 2, hs * d->x_cstride * 2, bs * d->x_cstride * 2};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(p.BK), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      gwd_set_error("cuTensorMapEncodeTiled(activation) failed: %d", static_cast<int>(r));
      return GWD_ERR_CUDA;
    }
  }
  {
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->n_pad),
                          static_cast<cuuint64_t>(d->taps) * (d->w_per_image ? d->B : 1)};
    cuuint64_t gstr[2] = {static_cast<cuuint64_t>(d->cin) * 2, static_cast<cuuint64_t>(d->n_pad) * d->cin * 2};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(p.BK), static_cast<cuuint32_t>(Nb),
                         static_cast<cuuint32_t>(p.resident ? d->taps : p.nsub)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      gwd_set_error("cuTensorMapEncodeTiled(weights) failed: %d", static_cast<int>(r));
      return GWD_ERR_CUDA;
    }
  }

  const size_t bar_bytes =
      ((2 * kMaxStages + 2 * kMaxAcc + 1) * sizeof(uint64_t) + 16 + (has_ln ? 2 * 4 * 128 * sizeof(float2) + 512 * sizeof(float) + 16 : 0) + 127) & ~size_t(127);
  // TMA epilogue (see the kernel): plain dense [rows, N] Linears whose epilogue warps own 64 columns each
  CUtensorMap map_res, map_y;
  memset(&map_res, 0, sizeof(map_res));
  memset(&map_y, 0, sizeof(map_y));
  bool tma_ep = false;
  {
    static const bool enabled = []() { const char* e = getenv("GWD_GEMM_TMAEP"); return !(e && e[0] == '0'); }();
    const bool strided_x = d->x_wstride != 0;
    const size_t ep_bytes = 1024 + static_cast<size_t>(kMaxEpilogueWarps) * 4096;
    // one 4-warp group per 64 output columns: N tiles of 128 / 192 / 256 columns -> 8 / 12 / 16 epilogue warps
    if (enabled && d->taps == 1 && d->B == 1 && d->H == 1 && !strided_x && ctas == 1 && !p.tile_split && !has_ln && !d->y_f32 &&
        d->y_raw == nullptr && !d->upsample2 && !d->w_per_image && d->pre_act == GWD_ACT_NONE && d->out_scale == 1.f &&
        (d->post_act == GWD_ACT_NONE || d->post_act == GWD_ACT_RELU || d->post_act == GWD_ACT_GELU) &&
        (p.Nt == 128 || p.Nt == 192 || p.Nt == 256) &&
        store_n % 8 == 0 && d->y_cstride % 8 == 0 && d->y_coff % 8 == 0 &&
        (d->res == nullptr || (d->res_cstride % 8 == 0 && d->res_coff % 8 == 0)) &&
        static_cast<int64_t>(p.m_tiles) * p.n_tiles >= gwd_num_sms()) {
      int stages = p.stages;
      const size_t slot = p.a_stage_bytes + p.b_stage_bytes;
      auto total = [&](int st) { return 1024 + st * slot + w_region + ((bar_bytes + 1023) & ~size_t(1023)) + ep_bytes; };
      while (stages > 3 && total(stages) > 226 * 1024) --stages;
      if (total(stages) <= 226 * 1024) {
        p.stages = stages;
        tma_ep = true;
        epi_warps = p.Nt / 16;
        threads = (2 + epi_warps) * 32;
      }
    }
  }
  const size_t ring_bytes = static_cast<size_t>(p.stages) * (p.a_stage_bytes + p.b_stage_bytes) + w_region;
  size_t smem_bytes = 1024 + ring_bytes + bar_bytes;
  if (tma_ep) {
    p.stage_off = static_cast<uint32_t>((ring_bytes + bar_bytes + 1023) & ~size_t(1023));   // staging tiles 1024-aligned (SW128)
    smem_bytes = 1024 + p.stage_off + 1024 + static_cast<size_t>(kMaxEpilogueWarps) * 4096;
    const cuuint64_t rows = static_cast<cuuint64_t>(d->W);
    cuuint32_t box[2] = {64, 32};
    cuuint32_t estr[2] = {1, 1};
    {
      cuuint64_t gdim[2] = {static_cast<cuuint64_t>(store_n), rows};
      cuuint64_t gstr[1] = {static_cast<cuuint64_t>(d->y_cstride) * 2};
      void* base = static_cast<__nv_bfloat16*>(d->y) + d->y_coff;
      CUresult r = encode(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        gwd_set_error("cuTensorMapEncodeTiled(output) failed: %d", static_cast<int>(r));
        return GWD_ERR_CUDA;
      }
    }
    if (d->res != nullptr) {
      cuuint64_t gdim[2] = {static_cast<cuuint64_t>(store_n), rows};
      cuuint64_t gstr[1] = {static_cast<cuuint64_t>(d->res_cstride) * 2};
      void* base = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(d->res)) + d->res_coff;
      CUresult r = encode(&map_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        gwd_set_error("cuTensorMapEncodeTiled(residual) failed: %d", static_cast<int>(r));
        return GWD_ERR_CUDA;
      }
    }
  }
  const int total_tiles = p.m_tiles * p.n_tiles;
  int grid = ctas * gwd_num_sms();
  if (grid > total_tiles) grid = total_tiles;
  if (cta2) {   // one CTA pair per two SMs
    const int pairs_all = ((p.m_tiles + 1) / 2) * p.n_tiles;
    int pairs = gwd_num_sms() / 2;
    if (pairs > pairs_all) pairs = pairs_all;
    grid = 2 * pairs;
  }
#define GWD_GEMM_CASE_TMAEP(POST)                                                                             \
  if (tma_ep && d->post_act == POST) {                                                                        \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      GWD_CUDA(cudaFuncSetAttribute(gwd_tapgemm_kernel<GWD_ACT_NONE, POST, false, true>,                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                \
      attr_set = true;                                                                                        \
    }                                                                                                         \
    GWD_CUDA(gwd_launch(gwd_tapgemm_kernel<GWD_ACT_NONE, POST, false, true>, dim3(grid), dim3(threads), smem_bytes, stream, 1, \
                        map_a, map_b, p, map_res, map_y));                                                    \
    launched = true;                                                                                          \
  }
#define GWD_GEMM_CASE(PRE, POST, LN)                                                                          \
  if (d->pre_act == PRE && d->post_act == POST && has_ln == LN && cta2) {                                     \
    auto kfn = gwd_tapgemm_kernel<PRE, POST, LN, false, true>;                                                \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      GWD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));           \
      attr_set = true;                                                                                        \
    }                                                                                                         \
    GWD_CUDA(gwd_launch(kfn, dim3(grid), dim3(threads), smem_bytes, stream, 2, map_a, map_b, p, map_res, map_y)); \
    launched = true;                                                                                          \
  } else if (d->pre_act == PRE && d->post_act == POST && has_ln == LN) {                                      \
    static bool attr_set = false;                                                                             \
    if (!attr_set) {                                                                                          \
      GWD_CUDA(cudaFuncSetAttribute(gwd_tapgemm_kernel<PRE, POST, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    227 * 1024));                                                             \
      attr_set = true;                                                                                        \
    }                                                                                                         \
    GWD_CUDA(gwd_launch(gwd_tapgemm_kernel<PRE, POST, LN>, dim3(grid), dim3(threads), smem_bytes, stream, 1,   \
                        map_a, map_b, p, map_res, map_y));                                                    \
    launched = true;                                                                                          \
  }
  bool launched = false;
  if (actgrad) {
#define GWD_GEMM_AG(TM, C2)                                                                                          \
  {                                                                                                                  \
    auto kfn = gwd_tapgemm_kernel<GWD_ACT_NONE, GWD_ACT_NONE, false, TM, C2, true>;                                  \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      GWD_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));                  \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    GWD_CUDA(gwd_launch(kfn, dim3(grid), dim3(threads), smem_bytes, stream, C2 ? 2 : 1, map_a, map_b, p, map_res, map_y)); \
  }
    if (tma_ep) GWD_GEMM_AG(true, false)
    else if (cta2) GWD_GEMM_AG(false, true)
    else GWD_GEMM_AG(false, false)
#undef GWD_GEMM_AG
    GWD_LAUNCHED();
    return GWD_OK;
  }
  GWD_GEMM_CASE_TMAEP(GWD_ACT_NONE)
  else GWD_GEMM_CASE_TMAEP(GWD_ACT_RELU)
  else GWD_GEMM_CASE_TMAEP(GWD_ACT_GELU)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_NONE, false)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_RELU, false)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_GELU, false)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_ELU, false)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_SIGMOID, false)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_NONE, true)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_GELU, true)
  else GWD_GEMM_CASE(GWD_ACT_NONE, GWD_ACT_RELU, true)
  else GWD_GEMM_CASE(GWD_ACT_ELU, GWD_ACT_NONE, true)
#undef GWD_GEMM_CASE
#undef GWD_GEMM_CASE_TMAEP
  if (!launched) {
    gwd_set_error("gwd_conv_gemm: epilogue combination pre_act=%d post_act=%d ln=%d is not instantiated", d->pre_act,
                  d->post_act, static_cast<int>(has_ln));
    return GWD_ERR_ARG;
  }
  GWD_LAUNCHED();
  return GWD_OK;
}
