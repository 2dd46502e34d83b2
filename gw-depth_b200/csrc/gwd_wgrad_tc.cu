// gwd_wgrad_tc.cu -- 3x3 convolution WEIGHT GRADIENT on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for the wide
// convolutions of the dense branch (PyramidLayer 160->160 / 800->320 at 1/4 scale: src/models/points/points_sample.py:12-43,
// 106-125 under torch.autograd):
//
//     dW[tap][n][c] += sum_{b,y,x} dY[b,y,x,n] * X[b, y+dy-1, x+dx-1, c]            tap = dx*3 + dy
//
// Per tap this is a GEMM D[M, N] = A^T B whose contraction runs over the PIXELS, i.e. over the slow axis of both
// channels-last operands.  On tcgen05 that needs no transposed copies: both operands are fed as MN-MAJOR shared-memory
// tiles (64-channel x 64-pixel atoms, 128-byte swizzle, exactly what one TMA box of the channels-last map produces);
// the tap shift is a coordinate offset of the X box and the zero padding of the convolution is the TMA out-of-bounds fill.
//
//   CTA = (tap, one or TWO 128-row tiles of the M-side channels -- two accumulators in TMEM share every N-side tile load --,
//          <=256-column tile of the N-side channels, pixel split), 6 warps:
//     warp 4 (one lane) : TMA producer -- per 64-pixel chunk (16 x 4 pixel block of one image) the M-side and N-side atoms
//                         into a 3-stage ring of 64 KB (full / empty mbarriers)
//     warp 5 (one lane) : 4 (x 2) x tcgen05.mma (M = 128, K = 16 pixels) per chunk, fp32 accumulators [128 x N] in TMEM,
//                         tcgen05.commit releases the stage; the last commit signals the epilogue
//     warps 0..3        : TMEM -> registers -> vector atomics into the flat fp32 gradient buffer (the pixel splits and the
//                         optimizer's accumulate-into-G contract both want +=)
//   Which operand plays M is chosen per shape to minimise padded work (dW^T is accumulated when X plays M).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kStages = 3;
constexpr int kTW = 16, kTH = 4;                 // pixel block of one chunk (64 pixels = 4 MMA K-steps)
constexpr int kAtomBytes = 64 * 128;             // 64 pixels x 64 channels bf16
constexpr int kMaxNAtoms = 4;                    // N tile <= 256 channels
constexpr int kMaxMAtoms = 4;                    // up to TWO 128-row M tiles per CTA (they share every N-side tile load)

struct WgTcParams {
  float* dw;                 // [9][N][C] (tap-major), += semantics
  int N, C;                  // dY channels, X channels (logical extents of dW)
  int m_is_x;                // 1: the M side is X (D = dW^T tile)
  int m_cnt, n_cnt;          // channel counts of the M-side / N-side operand
  int n_tile;                // N-side tile width (multiple of 16, <= 256)
  int m_tiles, n_tiles, splits;
  int m_pair, m_units;       // 128-row M tiles per CTA (1 or 2) and CTAs along M
  uint32_t acc_stride;       // TMEM columns between the two accumulators
  int B, H, W, tiles_x, tiles_y, chunks;
  int taps;                  // 9: 3x3 convolution (tap = blockIdx-derived shift of the X box); 1: Linear (no shift)
  int tw, th;                // pixel block of one 64-pixel chunk: 16 x 4 (maps) or 64 x 1 (plain row matrices)
  int64_t dw_rs;             // row stride of dW (floats)
  uint32_t tmem_cols;
  float* db;                 // optional (Linear, dY on the M side): db[n] += sum_r dY[r][n] from one extra N = 16 MMA per K step
  uint32_t ones_col;         // against a shared-memory tile of ones; its accumulators live at this TMEM column (+16: second M tile)
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
    if (++spins > (1u << 26)) __trap();   // a protocol bug must fault, never hang the box
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// UMMA shared-memory descriptor of an MN-major, 128-byte-swizzled operand: in 16-byte units the canonical layout is
// ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- 8 units = 64 channels contiguous, the next 64-channel atom LBO bytes on, one
// K row (pixel) every 128 bytes, groups of 8 pixels SBO = 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(192, 1)
gwd_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_n,
                    const __grid_constant__ WgTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // work unit
  int u = blockIdx.x;
  const int split = u % p.splits; u /= p.splits;
  const int nt = u % p.n_tiles; u /= p.n_tiles;
  const int mt = u % p.m_units; u /= p.m_units;
  const int tap = u;                                   // dx * 3 + dy
  const int sx = p.taps == 9 ? tap / 3 - 1 : 0, sy = p.taps == 9 ? tap % 3 - 1 : 0;
  const int per = (p.chunks + p.splits - 1) / p.splits;
  const int q_begin = split * per, q_end = min(p.chunks, q_begin + per);
  const int iters = q_end - q_begin;
  if (iters <= 0) return;
  const int m0 = mt * p.m_pair * 128, n0 = nt * p.n_tile;
  const int m_valid = min(128, p.m_cnt - m0), n_valid = min(p.n_tile, p.n_cnt - n0);
  const int m_atoms = (m_valid + 63) >> 6;             // atoms that hold any valid channel (the rest are never read out)
  const int m_valid_b = p.m_pair == 2 ? max(0, min(128, p.m_cnt - m0 - 128)) : 0;     // second M tile of this CTA (may be absent)
  const int m_atoms_b = (m_valid_b + 63) >> 6;
  const int n_mma = (n_valid + 15) & ~15;
  const int n_atoms = (n_mma + 63) >> 6;
  const int stage_bytes = (kMaxMAtoms + kMaxNAtoms) * kAtomBytes;

  // bias gradient: only the CTAs of the first N tile, which see every row of dY exactly once per M tile
  const bool do_db = p.db != nullptr && nt == 0;
  uint8_t* ones = smem + kStages * stage_bytes;        // one 8 KB atom of bf16 ones (any swizzle of all-ones is all-ones)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + kAtomBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* done = bars + 2 * kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (do_db) {
    for (int i = threadIdx.x; i < kAtomBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0 x 2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core's reads
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
  // launched programmatically (gwd_launch): everything above used kernel parameters, shared memory and TMEM only and may have
  // overlapped the tail of the kernel in front; global memory is touched from here on.  The trigger comes after the TMEM
  // allocation (a dependent that took the columns first would sit in its wait for this kernel while holding them)
  gwd_pdl_wait();
  gwd_pdl_trigger();

  if (warp == 4) {
    if (elect_one()) {
      // the shifted operand is X: dY[pixel] meets X[pixel + (sy, sx)]
      const int msx = p.m_is_x ? sx : 0, msy = p.m_is_x ? sy : 0;
      const int nsx = p.m_is_x ? 0 : sx, nsy = p.m_is_x ? 0 : sy;
      const uint32_t bytes = static_cast<uint32_t>(m_atoms + m_atoms_b + n_atoms) * kAtomBytes;
      for (int it = 0; it < iters; ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(empty + s, ((it / kStages) - 1) & 1);
        int q = q_begin + it;
        const int tx = q % p.tiles_x; q /= p.tiles_x;
        const int ty = q % p.tiles_y; q /= p.tiles_y;
        const int x0 = tx * p.tw, y0 = ty * p.th, b = q;
        const uint32_t base = smem_u32(smem + s * stage_bytes);
        mbar_expect_tx(full + s, bytes);
        for (int a = 0; a < m_atoms; ++a)
          tma_load_4d(base + a * kAtomBytes, &map_m, full + s, m0 + a * 64, x0 + msx, y0 + msy, b);
        for (int a = 0; a < m_atoms_b; ++a)
          tma_load_4d(base + (2 + a) * kAtomBytes, &map_m, full + s, m0 + 128 + a * 64, x0 + msx, y0 + msy, b);
        for (int a = 0; a < n_atoms; ++a)
          tma_load_4d(base + (kMaxMAtoms + a) * kAtomBytes, &map_n, full + s, n0 + a * 64, x0 + nsx, y0 + nsy, b);
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      // instruction descriptor: D fp32, A / B bf16, both MN-major, N, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             (static_cast<uint32_t>(n_mma >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
      const uint32_t idesc_ones = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(16 >> 3) << 17) |
                                  (static_cast<uint32_t>(128 >> 4) << 24);
      for (int it = 0; it < iters; ++it) {
        const int s = it % kStages;
        mbar_wait(full + s, (it / kStages) & 1);
        fence_after();
        const uint32_t base = smem_u32(smem + s * stage_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {    // 16 pixels = 16 rows of 128 bytes per K step
          const uint64_t bdesc = make_desc_mn(base + kMaxMAtoms * kAtomBytes + k * 2048, kAtomBytes);
          umma(tmem_base, make_desc_mn(base + k * 2048, kAtomBytes), bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          if (m_valid_b > 0)
            umma(tmem_base + p.acc_stride, make_desc_mn(base + 2 * kAtomBytes + k * 2048, kAtomBytes), bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          if (do_db) {   // D1[m][0..15] += sum over the 16 pixels of dY[pixel][m] * 1
            const uint64_t odesc = make_desc_mn(smem_u32(ones), kAtomBytes);
            umma(tmem_base + p.ones_col, make_desc_mn(base + k * 2048, kAtomBytes), odesc, idesc_ones, (it > 0 || k > 0) ? 1u : 0u);
            if (m_valid_b > 0)
              umma(tmem_base + p.ones_col + 16, make_desc_mn(base + 2 * kAtomBytes + k * 2048, kAtomBytes), odesc, idesc_ones,
                   (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty + s);
      }
      umma_commit(done);
    }
  } else {
    // epilogue: thread <-> M row <-> TMEM lane
    mbar_wait(done, 0);
    fence_after();
    const int m = warp * 32 + lane;
    float* dw_tap = p.dw + static_cast<int64_t>(tap) * p.N * p.dw_rs;
    for (int tile = 0; tile < 2; ++tile) {
      const int mv = tile == 0 ? m_valid : m_valid_b;
      if (mv <= 0) break;
      const int mb = m0 + tile * 128;
      const bool row_ok = m < mv;
      const uint32_t t_row = tmem_base + tile * p.acc_stride + (static_cast<uint32_t>(warp * 32) << 16);
      if (do_db) {
        uint32_t r[16];
        tmem_ld16(tmem_base + p.ones_col + tile * 16 + (static_cast<uint32_t>(warp * 32) << 16), r);
        if (row_ok) atomicAdd(p.db + mb + m, __uint_as_float(r[0]));
      }
      for (int c = 0; c < n_mma; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_row + c, r);
        if (!row_ok) continue;
        if (!p.m_is_x) {          // D[m][j] = dW[n = mb + m][c = n0 + c + j]: 16 consecutive floats of one row
          float* dst = dw_tap + static_cast<int64_t>(mb + m) * p.dw_rs + n0 + c;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (c + j < n_valid)    // channel counts are multiples of 4 here (checked on the host)
              atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                         __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        } else {                  // D[m][j] = dW[n = n0 + c + j][c = mb + m]: lanes of a warp are contiguous in c
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < n_valid) atomicAdd(dw_tap + static_cast<int64_t>(n0 + c + j) * p.dw_rs + mb + m, __uint_as_float(r[j]));
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 4) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
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

// 4D map over a channels-last bf16 map [B,H,W,cs] restricted to its first `ch` channels:
// box = 64 channels x 16 x 4 pixels of one image, 128-byte swizzle, zero fill outside the map
int make_map(CUtensorMap* m, const void* base, int B, int H, int W, int64_t cs, int ch, int tw = kTW, int th = kTH) {
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(ch), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(cs) * 2, static_cast<cuuint64_t>(cs) * 2 * W, static_cast<cuuint64_t>(cs) * 2 * W * H};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// padded MMA work of a role assignment (M side padded to 128, N side tiled to <= 256 in multiples of 16)
int64_t padded_work(int m_cnt, int n_cnt) {
  const int n_tiles = (n_cnt + 255) / 256;
  const int n_tile = ((n_cnt + n_tiles - 1) / n_tiles + 15) & ~15;
  return static_cast<int64_t>((m_cnt + 127) / 128) * 128 * n_tiles * n_tile;
}

}  // namespace

// returns 1 when the problem is not eligible for the tensor-core path (the caller falls back to the mma.sync kernels),
// 0 on success, a negative gwd error code otherwise
static int launch_tc(WgTcParams& p, const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, cudaStream_t stream) {
  const int N = p.N, C = p.C;
  p.m_is_x = padded_work(C, N) < padded_work(N, C) ? 1 : 0;
  p.m_cnt = p.m_is_x ? C : N;
  p.n_cnt = p.m_is_x ? N : C;
  p.m_tiles = (p.m_cnt + 127) / 128;
  p.n_tiles = (p.n_cnt + 255) / 256;
  p.n_tile = ((p.n_cnt + p.n_tiles - 1) / p.n_tiles + 15) & ~15;
  p.tiles_x = (p.W + p.tw - 1) / p.tw;
  p.tiles_y = (p.H + p.th - 1) / p.th;
  p.chunks = p.B * p.tiles_x * p.tiles_y;
  p.m_pair = p.m_tiles >= 2 ? 2 : 1;
  p.m_units = (p.m_tiles + p.m_pair - 1) / p.m_pair;
  p.acc_stride = static_cast<uint32_t>((p.n_tile + 31) & ~31);
  const int units = p.taps * p.m_units * p.n_tiles;
  int splits = (gwd_num_sms() + units / 2) / units;
  splits = max(1, min(splits, p.chunks / 8));
  p.splits = splits;
  uint32_t need = p.acc_stride * (p.m_pair - 1) + static_cast<uint32_t>(p.n_tile);
  if (p.db != nullptr) {     // bias gradient on the tensor core: needs dY on the M side and 32 spare TMEM columns
    p.ones_col = (need + 15u) & ~15u;
    if (p.m_is_x || p.taps != 1 || p.ones_col + 32 > 512) p.db = nullptr;
    else need = p.ones_col + 32;
  }
  uint32_t cols = 32;
  while (cols < need) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap map_dy, map_x;
  if (make_map(&map_dy, dy, p.B, p.H, p.W, dy_cs, N, p.tw, p.th) || make_map(&map_x, x, p.B, p.H, p.W, x_cs, C, p.tw, p.th)) return 1;
  const size_t smem = static_cast<size_t>(kStages) * (kMaxMAtoms + kMaxNAtoms) * kAtomBytes + kAtomBytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>(units * splits);
  GWD_CUDA(gwd_launch(gwd_wgrad_tc_kernel, dim3(grid), dim3(192), smem, stream, 1, p.m_is_x ? map_x : map_dy, p.m_is_x ? map_dy : map_x, p));
  GWD_LAUNCHED();
  return GWD_OK;
}

// returns 1 when the problem is not eligible for the tensor-core path (the caller falls back to the mma.sync kernels),
// 0 on success, a negative gwd error code otherwise
int gwd_conv3x3_wgrad_tc_try(const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, int B, int H, int W, int N, int C,
                             float* dw, cudaStream_t stream) {
  if (N % 16 || C % 16 || (N <= 64 && C <= 64) || N > 1024 || C > 1024) return 1;
  if (H < kTH || W < kTW || static_cast<int64_t>(B) * H * W < 4096) return 1;
  if (dy_cs % 8 || x_cs % 8 || dy_cs < N || x_cs < C) return 1;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dw)) & 15) return 1;
  if (!encode_fn()) return 1;
  WgTcParams p;
  memset(&p, 0, sizeof(p));
  p.dw = dw; p.N = N; p.C = C; p.dw_rs = C;
  p.B = B; p.H = H; p.W = W;
  p.taps = 9; p.tw = kTW; p.th = kTH;
  return launch_tc(p, dy, dy_cs, x, x_cs, stream);
}

// Linear weight gradient dW[n][k] += sum_r dY[r][n] X[r][k] over many rows (the 1/4- and 1/8-scale Swin stages: 85 k - 325 k window
// tokens): the same kernel with one "tap", the row matrices seen as a [C, rows] map walked in 64-row boxes
// db (optional): the bias gradient; *db_done tells the caller whether it was accumulated here (else: a column-sum pass)
int gwd_linear_wgrad_tc_try(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int N, int K, float* dw,
                            int64_t dw_rs, float* db, int* db_done, cudaStream_t stream) {
  if (db_done) *db_done = 0;
  if (N % 16 || K % 16 || N > 1024 || K > 1024 || rows < 32768 || rows >= (1ll << 31)) return 1;
  if (dy_rs % 8 || x_rs % 8 || dy_rs < N || x_rs < K || dw_rs % 4 || dw_rs < K) return 1;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dw)) & 15) return 1;
  if (!encode_fn()) return 1;
  WgTcParams p;
  memset(&p, 0, sizeof(p));
  p.dw = dw; p.N = N; p.C = K; p.dw_rs = dw_rs;
  p.B = 1; p.H = 1; p.W = static_cast<int>(rows);
  p.taps = 1; p.tw = 64; p.th = 1;
  static const bool fuse_db = [] { const char* e = getenv("GWD_WGRAD_DB"); return !(e && e[0] == '0'); }();
  p.db = fuse_db ? db : nullptr;
  const int rc = launch_tc(p, dy, dy_rs, x, x_rs, stream);
  if (rc == 0 && db_done) *db_done = p.db != nullptr;
  return rc;
}
