// gwd_attn_win.cu -- window attention with relative-position bias (+ shifted-window mask), persistent CTAs.
//
// The (shifted-)window attention cores of src/models/multiscale_transformerr.py:311-328 (WindowAttention) and :539-556
// (WindowClassAttention): windows of N = ws*ws <= 64 tokens, head dims 4..32, a [heads, N, N] bias table and an
// optional [nW, N, N] 0 / -100 mask.  At the finest level this is 16 images x 414 windows x 16 heads of 49x49 scores
// with head_dim 4: the matrix products are tiny, the cost is the 254 M soft-max elements, and a thread-per-query
// kernel spends its time broadcasting K/V rows out of shared memory.  Design:
//   * one CTA owns a group of heads (64 channels) and walks many windows: the bias of its heads sits in shared memory
//     (bf16, pre-multiplied by log2 e) for the CTA's whole life;
//   * per (head, 16-query tile) one warp runs S = Q K^T and O = P V on warp-level mma.sync m16n8k16 (bf16 in, fp32
//     accumulate; head dims below 16 are zero-padded in the fragments).  A 49-token window is 1/8 of the smallest
//     tcgen05 tile, so the 5th-generation path has nothing to offer here; the soft-max runs on the accumulator
//     fragments in registers and P is re-used as the A operand of P V without leaving the register file;
//   * the shifted-window mask of a window becomes one 64-bit word per query (warp ballots), and unmasked rows skip it;
//   * output tiles are staged so that rows leave as 8-byte pieces.
#include <stdlib.h>
#include <string.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

struct WinParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  int items, heads, N, hg, nW;
  int64_t q_is, q_rs, k_is, k_rs, v_is, v_rs, o_is, o_rs;
  const float* bias;   // [heads, N, N]
  const float* mask;   // [nW, N, N] entries 0 or one negative constant, or null
  float scale;
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 256;
// kD = channels (heads of the group x head_dim) a CTA works on: 32 for head dims 4 / 8, else 64, so that the bias of
// the group (hg x N x N bf16) plus the tiles leave room for at least two CTAs per SM.
// kRowWords = shared-memory row stride of the Q / K tiles in 32-bit words: kD/2 + 4, so that the 8 rows x 4 words of
// a fragment load fall into 32 different banks.  V^T rows hold 64 keys (+ the same 4 words of padding).
constexpr int kVtWords = 36;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// d = a * b (no accumulator input: the zero C operand costs no register initialisation)
__device__ __forceinline__ void mma_16816_first(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

// N (tokens per window) is a compile-time constant: tile counts, the key-padding selects and all shared-memory strides
// fold into immediates (the run-time-N version of this kernel executed 2.2x the instructions).
template <int HD, int kD, int N>
__global__ void __launch_bounds__(kThreads, 2) gwd_window_attention_kernel(const WinParams p) {
  gwd_pdl_trigger();   // a programmatically launched dependent may start its prologue (gwd_common.cuh, PDL)
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int kRowWords = kD / 2 + 4;
  constexpr int KS = (HD + 15) / 16;     // k-steps of Q K^T
  constexpr int DN = (HD + 7) / 8;       // 8-wide output column tiles of P V
  constexpr int NT = (N + 7) / 8;        // 8-key column tiles that hold real keys
  constexpr int MT = (N + 15) / 16;      // 16-query row tiles
  constexpr int NP = NT * 8;             // bias row stride (padding columns are zero)
  constexpr int HG = kD / HD;            // heads per CTA
  uint32_t* Qs = reinterpret_cast<uint32_t*>(smraw);          // [64 queries][kRowWords]   rows >= N stay zero
  uint32_t* Ks = Qs + 64 * kRowWords;                         // [64 keys][kRowWords]
  uint32_t* Vt = Ks + 64 * kRowWords;                         // [kD + 8 channels][kVtWords]  V transposed: [channel][key]
  bf16* Os = reinterpret_cast<bf16*>(Vt + (kD + 8) * kVtWords);    // [64][kD]
  bf16* bias_s = Os + 64 * kD;                                // [HG][N][NP], pre-multiplied by log2(e)
  unsigned long long* mask_s = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(bias_s + HG * N * NP) + 7) & ~uintptr_t(7));   // [64]
  float* mask_val_s = reinterpret_cast<float*>(mask_s + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kThreads / 32;
  const int g = lane >> 2, t = lane & 3;
  const int h0 = blockIdx.y * HG;

  for (int i = tid; i < 128 * kRowWords + (kD + 8) * kVtWords; i += kThreads) Qs[i] = 0u;   // zero padding rows / columns
  for (int idx = tid; idx < HG * N * NP; idx += kThreads) {
    const int row = idx / NP, j = idx - row * NP;   // row = h * N + i
    bias_s[idx] = __float2bfloat16(j < N ? __ldg(p.bias + static_cast<int64_t>(h0) * N * N + row * N + j) * kLog2e : 0.f);
  }
  for (int i = tid; i < 64; i += kThreads) mask_s[i] = 0ull;
  const float qscale = p.scale * kLog2e;
  __syncthreads();

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    // ---- stage Q, K (row major) and V (transposed) of the group's channels; build the mask words ----
    const bf16* qb = p.q + item * p.q_is + h0 * HD;
    const bf16* kb = p.k + item * p.k_is + h0 * HD;
    const bf16* vb = p.v + item * p.v_is + h0 * HD;
    bf16* Vt16 = reinterpret_cast<bf16*>(Vt);
    constexpr int kWpr = kD / 2;                      // 32-bit words per staged row
    for (int idx = tid; idx < N * kWpr; idx += kThreads) {
      const int j = idx / kWpr, w = idx - j * kWpr;
      const uint32_t qq = __ldg(reinterpret_cast<const uint32_t*>(qb + j * p.q_rs) + w);
      const uint32_t kk = __ldg(reinterpret_cast<const uint32_t*>(kb + j * p.k_rs) + w);
      const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(vb + j * p.v_rs) + w);
      Qs[j * kRowWords + w] = qq;
      Ks[j * kRowWords + w] = kk;
      const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(&vv);
      Vt16[(2 * w) * (2 * kVtWords) + j] = v2.x;
      Vt16[(2 * w + 1) * (2 * kVtWords) + j] = v2.y;
    }
    if (p.mask != nullptr) {
      const float* mw = p.mask + static_cast<int64_t>(item % p.nW) * N * N;
      for (int i = warp; i < N; i += nwarps) {
        const float m0 = lane < N ? __ldg(mw + i * N + lane) : 0.f;
        const float m1 = lane + 32 < N ? __ldg(mw + i * N + lane + 32) : 0.f;
        const unsigned lo = __ballot_sync(0xffffffffu, m0 != 0.f), hi = __ballot_sync(0xffffffffu, m1 != 0.f);
        if (lane == 0) mask_s[i] = (static_cast<unsigned long long>(hi) << 32) | lo;
        if (m0 != 0.f) *mask_val_s = m0;   // all non-zero entries carry the same value
        if (m1 != 0.f) *mask_val_s = m1;
      }
    }
    __syncthreads();   // tiles + mask ready (and the previous window's output rows have left Os)

    for (int it = warp; it < HG * MT; it += nwarps) {
      const int hl = it / MT, mt = it - hl * MT;
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      const bool hi_rows = 16 * mt + 8 < N;   // warp-uniform: the last row tile of a 49-token window holds one row
      const int cw = (hl * HD) >> 1;          // first 32-bit word of this head inside a row
      // ---- S = Q K^T ----
      uint32_t a[KS][4];
      const uint32_t* q0 = Qs + r0 * kRowWords + cw + t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const bool lo_ok = 16 * ks + 2 * t < HD, hi_ok = 16 * ks + 8 + 2 * t < HD;
        a[ks][0] = lo_ok ? q0[8 * ks] : 0u;
        a[ks][1] = lo_ok ? q0[8 * kRowWords + 8 * ks] : 0u;
        a[ks][2] = hi_ok ? q0[8 * ks + 4] : 0u;
        a[ks][3] = hi_ok ? q0[8 * kRowWords + 8 * ks + 4] : 0u;
      }
      float s[NT][4];
      const uint32_t* k0 = Ks + g * kRowWords + cw + t;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const bool lo_ok = 16 * ks + 2 * t < HD, hi_ok = 16 * ks + 8 + 2 * t < HD;
          const uint32_t b0 = lo_ok ? k0[8 * nt * kRowWords + 8 * ks] : 0u;
          const uint32_t b1 = hi_ok ? k0[8 * nt * kRowWords + 8 * ks + 4] : 0u;
          if (ks == 0) mma_16816_first(s[nt], a[ks], b0, b1);
          else mma_16816(s[nt], a[ks], b0, b1);
        }
      }
      // ---- scores = scale * S + bias (+ mask), all in the log2 domain; row maxima ----
      const uint32_t* b_r0 = reinterpret_cast<const uint32_t*>(bias_s + (hl * N + (r0 < N ? r0 : 0)) * NP) + t;
      const uint32_t* b_r1 = reinterpret_cast<const uint32_t*>(bias_s + (hl * N + (r1 < N ? r1 : 0)) * NP) + t;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 bb0 = gwd_unpack_bf16x2(b_r0[4 * nt]);
        const float2 bb1 = gwd_unpack_bf16x2(b_r1[4 * nt]);
        s[nt][0] = fmaf(s[nt][0], qscale, bb0.x);
        s[nt][1] = fmaf(s[nt][1], qscale, bb0.y);
        s[nt][2] = fmaf(s[nt][2], qscale, bb1.x);
        s[nt][3] = fmaf(s[nt][3], qscale, bb1.y);
        if (8 * nt + 8 > N) {   // compile-time: only the last column tile holds padding keys
          const int j0 = 8 * nt + 2 * t;
          if (j0 >= N) s[nt][0] = s[nt][2] = -INFINITY;
          if (j0 + 1 >= N) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      const unsigned long long mb0 = mask_s[r0], mb1 = mask_s[r1];   // zero without a mask and for rows >= N
      if ((mb0 | mb1) != 0ull) {
        const float mval = *mask_val_s * kLog2e;
        const uint32_t lo0 = static_cast<uint32_t>(mb0) >> (2 * t), hi0 = static_cast<uint32_t>(mb0 >> 32) >> (2 * t);
        const uint32_t lo1 = static_cast<uint32_t>(mb1) >> (2 * t), hi1 = static_cast<uint32_t>(mb1 >> 32) >> (2 * t);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint32_t w0 = (nt < 4 ? lo0 : hi0) >> (8 * (nt & 3)), w1 = (nt < 4 ? lo1 : hi1) >> (8 * (nt & 3));
          if (w0 & 1u) s[nt][0] += mval;
          if (w0 & 2u) s[nt][1] += mval;
          if (w1 & 1u) s[nt][2] += mval;
          if (w1 & 2u) s[nt][3] += mval;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      // ---- P = exp2(scores - max) packed straight into the A fragments of P V ----
      float l0 = 0.f, l1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) pa[kk][0] = pa[kk][1] = pa[kk][2] = pa[kk][3] = 0u;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float p0 = ex2_approx(s[nt][0] - mx0), p1 = ex2_approx(s[nt][1] - mx0);
        l0 += p0 + p1;
        pa[nt >> 1][(nt & 1) * 2] = gwd_pack_bf16x2(p0, p1);
        if (hi_rows) {
          const float p2 = ex2_approx(s[nt][2] - mx1), p3 = ex2_approx(s[nt][3] - mx1);
          l1 += p2 + p3;
          pa[nt >> 1][(nt & 1) * 2 + 1] = gwd_pack_bf16x2(p2, p3);
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float inv0 = rcp_approx(l0), inv1 = rcp_approx(l1);
      // ---- O = P V ----
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        float o[4];
        const uint32_t* vrow = Vt + (hl * HD + 8 * dn + g) * kVtWords + t;   // channel (output column) g of this tile
        mma_16816_first(o, pa[0], vrow[0], vrow[4]);
#pragma unroll
        for (int kk = 1; kk < (NT + 1) / 2; ++kk) mma_16816(o, pa[kk], vrow[8 * kk], vrow[8 * kk + 4]);
        if (8 * dn + 2 * t < HD) {
          bf16* orow = Os + r0 * kD + hl * HD + 8 * dn + 2 * t;
          if (r0 < N) *reinterpret_cast<uint32_t*>(orow) = gwd_pack_bf16x2(o[0] * inv0, o[1] * inv0);
          if (r1 < N) *reinterpret_cast<uint32_t*>(orow + 8 * kD) = gwd_pack_bf16x2(o[2] * inv1, o[3] * inv1);
        }
      }
    }
    __syncthreads();   // output tile complete; Q / K / V free for the next window

    // ---- rows of the output tile leave as 8-byte pieces ----
    bf16* ob = p.o + item * p.o_is + h0 * HD;
    constexpr int kPpr = kD / 4;
    for (int idx = tid; idx < N * kPpr; idx += kThreads) {
      const int i = idx / kPpr, w = idx - i * kPpr;
      *reinterpret_cast<uint2*>(ob + i * p.o_rs + 4 * w) = *reinterpret_cast<const uint2*>(Os + i * kD + 4 * w);
    }
  }
}

template <int HD, int kD>
int launch_window(const WinParams& p, cudaStream_t stream) {
  constexpr int N = 49, NP = 56;
  const size_t smem = static_cast<size_t>(128 * (kD / 2 + 4) + (kD + 8) * kVtWords) * 4 + 64 * kD * 2 +
                      static_cast<size_t>(p.hg) * N * NP * 2 + 8 + 64 * 8 + 16;
  static int occ = 0;
  static size_t occ_smem = 0;
  if (occ == 0 || occ_smem != smem) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_window_attention_kernel<HD, kD, 49>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int nb = 0;
    GWD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gwd_window_attention_kernel<HD, kD, 49>, kThreads, smem));
    GWD_CHECK_ARG(nb > 0, "gwd_attention(window): configuration does not fit an SM");
    occ = nb; occ_smem = smem;
  }
  const int groups = p.heads / p.hg;
  int per_group = (occ * gwd_num_sms() + groups - 1) / groups;
  if (per_group > p.items) per_group = p.items;
  dim3 grid(static_cast<unsigned>(per_group), static_cast<unsigned>(groups));
  gwd_window_attention_kernel<HD, kD, 49><<<grid, kThreads, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

}  // namespace

// returns 0 = launched, < 0 = error, 1 = not a window-attention problem (caller falls through to the generic kernel)
int gwd_attention_window_try(const gwd_attn_desc* d, cudaStream_t stream) {
  if (d->bias == nullptr || d->key_padding != nullptr || d->Lq != d->Lk || d->Lq != 49) return 1;   // 7x7 windows
  const int D = d->hd <= 8 ? 32 : 64;
  if (d->heads * d->hd < D || D % d->hd != 0) return 1;
  const int hg = D / d->hd;
  if (d->heads % hg != 0) return 1;
  // 8-byte output pieces / 4-byte input words
  if (d->o_row_stride % 4 != 0 || d->o_item_stride % 4 != 0 || (reinterpret_cast<uintptr_t>(d->o) & 7) != 0) return 1;
  if (d->q_row_stride % 2 != 0 || d->q_item_stride % 2 != 0 || (reinterpret_cast<uintptr_t>(d->q) & 3) != 0) return 1;
  WinParams p;
  p.q = static_cast<const bf16*>(d->q); p.k = static_cast<const bf16*>(d->k); p.v = static_cast<const bf16*>(d->v);
  p.o = static_cast<bf16*>(d->o);
  p.items = d->items; p.heads = d->heads; p.N = d->Lq; p.hg = hg;
  p.nW = d->mask_windows > 0 ? d->mask_windows : 1;
  p.q_is = d->q_item_stride; p.q_rs = d->q_row_stride; p.k_is = d->k_item_stride; p.k_rs = d->k_row_stride;
  p.v_is = d->v_item_stride; p.v_rs = d->v_row_stride; p.o_is = d->o_item_stride; p.o_rs = d->o_row_stride;
  p.bias = d->bias; p.mask = d->mask; p.scale = d->scale;
  switch (d->hd) {
    case 4: return launch_window<4, 32>(p, stream);
    case 8: return launch_window<8, 32>(p, stream);
    case 16: return launch_window<16, 64>(p, stream);
    case 32: return launch_window<32, 64>(p, stream);
    default: return 1;
  }
}

// =====================================================================================================================
// class-token CHANNEL attention of WindowClassAttention (multiscale_transformerr.py:561-578) on mma.sync
// =====================================================================================================================
// Per window and head: A[r][c] = softmax_c(scale * sum_n tq[n][r] tk[n][c]) with r = (depth | seg) x 4 token channels and
// c = tc key channels, then out[n][r] = sum_c A[r][c] tv[n][c].  Both products contract over an axis that is the ROW
// axis of the row-major operands in shared memory, which is exactly what ldmatrix.trans delivers, so the window is
// staged once (8-byte copies, each head's tc channels padded to a multiple of 8 so that every 8x8 block is 16-byte
// aligned) and each warp runs two heads: 4 k-steps of m16n8k16 for the scores, a 12..24-wide soft-max on the
// accumulator fragments, and one MMA per 16 tokens for the output, whose B operand is the soft-max result still in
// registers.  The thread-per-output kernel in gwd_attn.cu (730 us on 6624 windows) stays as the generic fallback.
namespace {

struct TokMmaParams {
  const bf16* dq; const bf16* sq; const bf16* tk; const bf16* tv;
  bf16* dout; bf16* sout;
  int items, N, heads, tc;
  int64_t q_rs, k_rs, v_rs, o_rs;
  float scale;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// NTC = 8-wide key-channel tiles per head (tc padded to 8 * NTC)
template <int NTC>
__global__ void __launch_bounds__(256, 2) gwd_token_attention_mma_kernel(const TokMmaParams p) {
  gwd_pdl_trigger();   // a programmatically launched dependent may start its prologue (gwd_common.cuh, PDL)
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int TCP = 8 * NTC;
  const int N = p.N, heads = p.heads, tc = p.tc;
  const int SQ = heads * 16 + 16;             // bytes per staged query row: [head][depth x4 | seg x4] + 16 pad
  const int SK = heads * TCP * 2 + 16;        // bytes per staged key / value row: [head][TCP] + 16 pad
  const int SO = 256 + 8;                     // bytes per staged output row: depth 64 | seg 64 (+ pad)
  uint8_t* TQ = smraw;                        // [64][SQ]   rows >= N stay zero (they are the k padding of the scores)
  uint8_t* TK = TQ + 64 * SQ;                 // [64][SK]
  uint8_t* TV = TK + 64 * SK;                 // [64][SK] (+ 64 bytes slack)
  uint8_t* OS = TV + 64 * SK + 64;            // [64][SO]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  {
    const int words = (64 * SQ + 128 * SK + 64 + 64 * SO) >> 2;
    for (int i = tid; i < words; i += 256) reinterpret_cast<uint32_t*>(smraw)[i] = 0u;
  }
  __syncthreads();
  const int kp = tc >> 2;                     // 8-byte pieces per head in a key / value row
  const int m_tiles = (N + 15) >> 4;
  const float sc = p.scale * kLog2e;

  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    // ---- stage the window ----
    const int64_t row0 = static_cast<int64_t>(item) * N;
    for (int idx = tid; idx < N * heads; idx += 256) {          // queries: one 8-byte piece per (token, head, tensor)
      const int n = idx / heads, h = idx - n * heads;
      const uint2 d = __ldg(reinterpret_cast<const uint2*>(p.dq + (row0 + n) * p.q_rs) + h);
      const uint2 s = __ldg(reinterpret_cast<const uint2*>(p.sq + (row0 + n) * p.q_rs) + h);
      *reinterpret_cast<uint4*>(TQ + n * SQ + h * 16) = make_uint4(d.x, d.y, s.x, s.y);
    }
    const int kv_pieces = heads * kp;
    for (int idx = tid; idx < N * kv_pieces; idx += 256) {
      const int n = idx / kv_pieces, r = idx - n * kv_pieces;
      const int h = r / kp, w = r - h * kp;
      const uint2 k = __ldg(reinterpret_cast<const uint2*>(p.tk + (row0 + n) * p.k_rs) + r);
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(p.tv + (row0 + n) * p.v_rs) + r);
      *reinterpret_cast<uint2*>(TK + n * SK + h * TCP * 2 + w * 8) = k;
      *reinterpret_cast<uint2*>(TV + n * SK + h * TCP * 2 + w * 8) = v;
    }
    __syncthreads();

    for (int h = warp; h < heads; h += 8) {
      // ---- scores[r][c] = sum_n tq[n][r] tk[n][c]  (M = 8 real rows, N = TCP, K = 64 zero-padded tokens) ----
      float s[NTC][4];
      const uint32_t q_addr = smem_addr(TQ + (lane & 15) * SQ + h * 16);
      const uint32_t k_addr = smem_addr(TK + (lane & 15) * SK + h * TCP * 2);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        ldsm_x2_trans(a[0], a[2], q_addr + kk * 16 * SQ);
        a[1] = 0u;
        a[3] = 0u;
#pragma unroll
        for (int nt = 0; nt < NTC; ++nt) {
          uint32_t b0, b1;
          ldsm_x2_trans(b0, b1, k_addr + kk * 16 * SK + nt * 16);
          if (kk == 0) mma_16816_first(s[nt], a, b0, b1);
          else mma_16816(s[nt], a, b0, b1);
        }
      }
      // ---- soft-max over the tc key channels of every row g (accumulator rows g + 8 are padding) ----
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        const int c0 = 8 * nt + 2 * t;
        s[nt][0] = c0 < tc ? s[nt][0] * sc : -INFINITY;
        s[nt][1] = c0 + 1 < tc ? s[nt][1] * sc : -INFINITY;
        mx = fmaxf(mx, fmaxf(s[nt][0], s[nt][1]));
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float l = 0.f;
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) {
        s[nt][0] = ex2_approx(s[nt][0] - mx);
        s[nt][1] = ex2_approx(s[nt][1] - mx);
        l += s[nt][0] + s[nt][1];
      }
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = rcp_approx(l);
      uint32_t pb[4] = {0u, 0u, 0u, 0u};      // B fragments of the output product: B[k = c][col = r] = A[r][c]
#pragma unroll
      for (int nt = 0; nt < NTC; ++nt) pb[nt] = gwd_pack_bf16x2(s[nt][0] * inv, s[nt][1] * inv);
      // ---- out[n][r] = sum_c tv[n][c] A[r][c]  (M = tokens, N = 8, K = TCP padded to 16 / 32) ----
      const uint32_t v_addr = smem_addr(TV + (lane & 15) * SK + h * TCP * 2 + (lane >> 4) * 16);
      for (int mt = 0; mt < m_tiles; ++mt) {
        float o[4];
        uint32_t a[4];
        ldsm_x4(a, v_addr + mt * 16 * SK);
        mma_16816_first(o, a, pb[0], pb[1]);
        if (NTC > 2) {
          ldsm_x4(a, v_addr + mt * 16 * SK + 32);
          mma_16816(o, a, pb[2], pb[3]);
        }
        uint8_t* orow = OS + (16 * mt + g) * SO + (t >> 1) * 128 + h * 8 + (t & 1) * 4;
        *reinterpret_cast<uint32_t*>(orow) = gwd_pack_bf16x2(o[0], o[1]);
        *reinterpret_cast<uint32_t*>(orow + 8 * SO) = gwd_pack_bf16x2(o[2], o[3]);
      }
    }
    __syncthreads();

    // ---- rows leave as 16-byte pieces: depth 64 channels | seg 64 channels ----
    const int ppr = heads >> 1;                // 16-byte pieces per tensor row (heads * 4 channels * 2 bytes / 16)
    for (int idx = tid; idx < N * 2 * ppr; idx += 256) {
      const int n = idx / (2 * ppr), r = idx - n * 2 * ppr;
      const int which = r / ppr, w = r - which * ppr;
      const uint2 lo = *reinterpret_cast<const uint2*>(OS + n * SO + which * 128 + w * 16);
      const uint2 hi = *reinterpret_cast<const uint2*>(OS + n * SO + which * 128 + w * 16 + 8);
      bf16* dst = (which == 0 ? p.dout : p.sout) + (row0 + n) * p.o_rs + w * 8;
      *reinterpret_cast<uint4*>(dst) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
  }
}

template <int NTC>
int launch_token_mma(const TokMmaParams& p, cudaStream_t stream) {
  const size_t SQ = p.heads * 16 + 16, SK = static_cast<size_t>(p.heads) * 8 * NTC * 2 + 16, SO = 264;
  const size_t smem = 64 * SQ + 128 * SK + 64 + 64 * SO;
  static int occ = 0;
  static size_t occ_smem = 0;
  if (occ == 0 || occ_smem != smem) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_token_attention_mma_kernel<NTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int nb = 0;
    GWD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gwd_token_attention_mma_kernel<NTC>, 256, smem));
    GWD_CHECK_ARG(nb > 0, "gwd_token_attention: window does not fit shared memory");
    occ = nb; occ_smem = smem;
  }
  int grid = occ * gwd_num_sms();
  if (grid > p.items) grid = p.items;
  gwd_token_attention_mma_kernel<NTC><<<grid, 256, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

}  // namespace

// returns 0 = launched, < 0 = error, 1 = shape not covered (caller uses the generic kernel)
int gwd_token_attention_mma_try(const void* dq, const void* sq, const void* tk, const void* tv, void* dout, void* sout,
                                int items, int N, int heads, int td, int tc, int64_t q_rs, int64_t k_rs, int64_t v_rs,
                                int64_t o_rs, float scale, cudaStream_t stream) {
  if (td != 4 || N > 64 || N < 1 || heads < 2 || heads > 16 || heads % 2 != 0 || tc % 4 != 0 || tc < 4 || tc > 32) return 1;
  if (q_rs % 4 != 0 || k_rs % 4 != 0 || v_rs % 4 != 0 || o_rs % 8 != 0) return 1;
  const uintptr_t align8 = reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(sq) | reinterpret_cast<uintptr_t>(tk) |
                           reinterpret_cast<uintptr_t>(tv);
  if ((align8 & 7) != 0 || ((reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(sout)) & 15) != 0) return 1;
  TokMmaParams p;
  p.dq = static_cast<const bf16*>(dq); p.sq = static_cast<const bf16*>(sq);
  p.tk = static_cast<const bf16*>(tk); p.tv = static_cast<const bf16*>(tv);
  p.dout = static_cast<bf16*>(dout); p.sout = static_cast<bf16*>(sout);
  p.items = items; p.N = N; p.heads = heads; p.tc = tc;
  p.q_rs = q_rs; p.k_rs = k_rs; p.v_rs = v_rs; p.o_rs = o_rs; p.scale = scale;
  const int ntc = (tc + 7) / 8;
  switch (ntc) {
    case 1: case 2: return launch_token_mma<2>(p, stream);
    case 3: return launch_token_mma<3>(p, stream);
    default: return launch_token_mma<4>(p, stream);
  }
}
