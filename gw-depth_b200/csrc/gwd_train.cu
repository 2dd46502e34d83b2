// gwd_train.cu -- backward and optimizer kernels of the line branch (DETR encoder / decoder, line heads):
//   gwd_layernorm_bwd   : dz, dgamma, dbeta of y = LN(z)                         (HBM bound, warp per row)
//   gwd_act_bwd         : dy * act'(y) for ReLU / sigmoid from the OUTPUT, with dtype conversion + channel padding
//   gwd_transpose       : bf16 [rows, C] -> [C, rows_pad] (+ fp32 column sums = bias gradients); the weight-gradient
//                         GEMMs dW = dY^T X then run on gwd_conv_gemm with dY^T as the row operand and X^T as the filter
//   gwd_attention_bwd   : dQ, dK, dV of O = softmax(scale Q K^T) V, soft-max recomputed (two passes, no atomics)
//   gwd_sumsq / gwd_adamw_step : global gradient norm and fused clip + AdamW over a flat parameter segment, writing
//                         the bf16 mirror the forward kernels read
// Data-gradient GEMMs dX = dY W run on gwd_conv_gemm with the transposed weight mirror.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y), c = gwd_unpack_bf16x2(u.z), d = gwd_unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = gwd_pack_bf16x2(f[0], f[1]); u.y = gwd_pack_bf16x2(f[2], f[3]);
  u.z = gwd_pack_bf16x2(f[4], f[5]); u.w = gwd_pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (z - mean) * rstd * gamma + beta over C channels (biased variance), so with
// xh = (z - mean) * rstd and g = dy * gamma:   dz = rstd * (g - mean(g) - xh * mean(g * xh)),
// dgamma = sum_rows dy * xh, dbeta = sum_rows dy.  A warp owns a row (lane l holds the 16-byte vectors l, l+32, ...),
// keeps the column partial sums of all its rows in registers, the 8 warps of a CTA combine them in shared memory and
// one atomicAdd per column and CTA reaches HBM.
// ------------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;

// LPR lanes own one row (32: one row per warp, NV chunks of 256 columns; 16 / 8: two / four rows per warp for C <= 128 / 64,
// so that narrow rows -- the 64-channel maps of the dense head -- still fill every lane)
template <int NV, int LPR>
__global__ void __launch_bounds__(kLnWarps * 32)
gwd_layernorm_bwd_kernel(const bf16* __restrict__ dy, int64_t dy_rs, const bf16* __restrict__ z, int64_t z_rs,
                         const float* __restrict__ gamma, const float* __restrict__ beta, int post_act, float eps,
                         const bf16* __restrict__ add, int64_t add_rs,
                         bf16* __restrict__ dz, int64_t dz_rs, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         int64_t rows, int C, int n) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  static_assert(LPR == 32 || NV == 1, "row groups narrower than a warp hold one chunk");
  constexpr int RPW = 32 / LPR;                   // rows per warp and iteration
  __shared__ float part[kLnWarps][NV * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cl = lane % LPR, sub = lane / LPR;
  const int64_t gwarp = static_cast<int64_t>(blockIdx.x) * kLnWarps + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kLnWarps;
  auto group_sum = [](float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  // the statistics run over the n LOGICAL channels; channels n..C are padding (zeros in z, zeros out in dz)
  float ag[NV][8], ab[NV][8], gm[NV][8], bt[NV][8];
  bool on[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    on[v] = (cl + 32 * v) * 8 < C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool in = (cl + 32 * v) * 8 + e < n;
      ag[v][e] = 0.f; ab[v][e] = 0.f;
      gm[v][e] = in ? gamma[(cl + 32 * v) * 8 + e] : 0.f;
      bt[v][e] = (in && post_act != GWD_ACT_NONE) ? beta[(cl + 32 * v) * 8 + e] : 0.f;
    }
  }
  const float invC = 1.f / static_cast<float>(n);
  // software pipeline: the raw 16-byte vectors of the NEXT row (and the residual gradient of the current one) are requested
  // before the current row's four dependent reductions run -- a warp with one row in flight kept ~20 KB per SM outstanding,
  // which caps the pass at a fraction of the HBM rate (0.9 TB/s measured on the 160-channel pyramid maps)
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  auto cvt8 = [](const uint4& u, float (&f)[8]) {
    const float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y), c = gwd_unpack_bf16x2(u.z), d = gwd_unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  };
  uint4 zn[NV], dn[NV];
  {
    const int64_t row = gwarp * RPW + sub;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const bool ld = on[v] && row < rows;
      zn[v] = ld ? *reinterpret_cast<const uint4*>(z + row * z_rs + (cl + 32 * v) * 8) : zero4;
      dn[v] = ld ? *reinterpret_cast<const uint4*>(dy + row * dy_rs + (cl + 32 * v) * 8) : zero4;
    }
  }
  for (int64_t base = gwarp * RPW; base < rows; base += nwarps * RPW) {
    const int64_t row = base + sub;
    const bool live = row < rows;
    uint4 zc[NV], dc[NV], ac[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { zc[v] = zn[v]; dc[v] = dn[v]; }
    {
      const int64_t nrow = row + nwarps * RPW;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const bool ld = on[v] && nrow < rows;
        zn[v] = ld ? *reinterpret_cast<const uint4*>(z + nrow * z_rs + (cl + 32 * v) * 8) : zero4;
        dn[v] = ld ? *reinterpret_cast<const uint4*>(dy + nrow * dy_rs + (cl + 32 * v) * 8) : zero4;
        ac[v] = (add != nullptr && on[v] && live) ? *reinterpret_cast<const uint4*>(add + row * add_rs + (cl + 32 * v) * 8) : zero4;
      }
    }
    float zv[NV][8], dv[NV][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      cvt8(zc[v], zv[v]);
      cvt8(dc[v], dv[v]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += zv[v][e];
    }
    const float mean = group_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        zv[v][e] = (live && (cl + 32 * v) * 8 + e < n) ? zv[v][e] - mean : 0.f;
        q += zv[v][e] * zv[v][e];
      }
    const float rstd = rsqrtf(group_sum(q) * invC + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        zv[v][e] *= rstd;                                  // xhat
        if (post_act != GWD_ACT_NONE) dv[v][e] *= gwd_act_grad(fmaf(zv[v][e], gm[v][e], bt[v][e]), post_act);   // through act(LN(z))
        ag[v][e] += dv[v][e] * zv[v][e];
        ab[v][e] += dv[v][e];
        dv[v][e] *= gm[v][e];                              // g (0 in the padding channels)
        m1 += dv[v][e];
        m2 += dv[v][e] * zv[v][e];
      }
    m1 = group_sum(m1) * invC;
    m2 = group_sum(m2) * invC;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!on[v] || !live) continue;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (cl + 32 * v) * 8 + e < n ? rstd * (dv[v][e] - m1 - zv[v][e] * m2) : 0.f;
      if (add != nullptr) {
        float a[8];
        cvt8(ac[v], a);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += a[e];
      }
      st8(dz + row * dz_rs + (cl + 32 * v) * 8, o);
    }
  }
  // column sums: warps (and the row groups of a warp) -> shared memory -> one atomic per column and CTA
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    if (dst == nullptr) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) part[warp][(lane + 32 * v) * 8 + e] = pass == 0 ? ag[v][e] : ab[v][e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kLnWarps * 32) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w)
#pragma unroll
        for (int r = 0; r < RPW; ++r) t += part[w][r * LPR * 8 + c];
      atomicAdd(dst + c, t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// activation backward from the OUTPUT value: ReLU: dy * (y > 0); sigmoid: dy * y * (1 - y); none: dy.
// Converts fp32 / bf16 inputs to a bf16 [rows, out_cols] matrix whose columns n..out_cols are zero.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_factor(float yy, int act, int from_input) { return gwd_act_factor(yy, act, from_input); }

template <typename TD, typename TY>
__global__ void gwd_act_bwd_kernel(const TD* __restrict__ dy, int64_t dy_rs, const TY* __restrict__ y, int64_t y_rs, int act,
                                   bf16* __restrict__ out, int64_t out_rs, int64_t rows, int n, int out_cols, float y_mul,
                                   float scale, int from_input) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  const int64_t total = rows * out_cols;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / out_cols;
    const int c = static_cast<int>(i - r * out_cols);
    float v = 0.f;
    if (c < n) {
      v = static_cast<float>(dy[r * dy_rs + c]);
      if (act != GWD_ACT_NONE) v *= act_factor(static_cast<float>(y[r * y_rs + c]) * y_mul, act, from_input);
      v *= scale;
    }
    out[r * out_rs + c] = __float2bfloat16(v);
  }
}

// ------------------------------------------------------------------------------------------------
// transpose (+ column sums).  64 x 64 tiles through shared memory, 4-byte accesses on both sides.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gwd_transpose_kernel(const bf16* __restrict__ x, int64_t x_rs, bf16* __restrict__ out, int64_t out_rs, int64_t rows,
                     int64_t rows_pad, int C, float* __restrict__ colsum) {
  __shared__ float tile[64][65];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int rr = ty; rr < 64; rr += 8) {
    const int64_t r = r0 + rr;
    const int c = c0 + 2 * tx;
    float2 v = make_float2(0.f, 0.f);
    if (r < rows && c < C)     // C is even
      v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * x_rs + c));
    tile[rr][2 * tx] = v.x;
    tile[rr][2 * tx + 1] = v.y;
  }
  __syncthreads();
  if (colsum != nullptr && threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < 64; ++rr) s += tile[rr][threadIdx.x];
    atomicAdd(colsum + c0 + threadIdx.x, s);
  }
  for (int cc = ty; cc < 64; cc += 8) {
    const int c = c0 + cc;
    const int64_t r = r0 + 2 * tx;
    if (c < C && r < rows_pad)     // rows_pad is even
      *reinterpret_cast<__nv_bfloat162*>(out + c * out_rs + r) = __floats2bfloat162_rn(tile[2 * tx][cc], tile[2 * tx + 1][cc]);
  }
}


// batched transpose: one launch for every weight mirror of the branch (89 matrices); CTA -> (matrix, 64x64 tile) through
// a prefix table of tile counts.  table[m] = {src, dst, rows, cols, src row stride, dst row stride}
__global__ void __launch_bounds__(256)
gwd_transpose_batch_kernel(const int64_t* __restrict__ table, const int32_t* __restrict__ tile_prefix, int n_mats) {
  __shared__ float tile[64][65];
  int lo = 0, hi = n_mats;                 // largest m with tile_prefix[m] <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tile_prefix[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const int64_t* e = table + 6 * lo;
  const bf16* x = reinterpret_cast<const bf16*>(e[0]);
  bf16* out = reinterpret_cast<bf16*>(e[1]);
  const int rows = static_cast<int>(e[2]), C = static_cast<int>(e[3]);
  const int64_t x_rs = e[4], out_rs = e[5];
  const int local = blockIdx.x - tile_prefix[lo];
  const int tiles_r = (rows + 63) >> 6;
  const int r0 = (local % tiles_r) * 64, c0 = (local / tiles_r) * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 64; rr += 8) {
    const int r = r0 + rr, c = c0 + 2 * tx;
    float2 v = make_float2(0.f, 0.f);
    if (r < rows && c < C) v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + r * x_rs + c));
    tile[rr][2 * tx] = v.x;
    tile[rr][2 * tx + 1] = v.y;
  }
  __syncthreads();
  for (int cc = ty; cc < 64; cc += 8) {
    const int c = c0 + cc, r = r0 + 2 * tx;
    if (c < C && r < rows)
      *reinterpret_cast<__nv_bfloat162*>(out + c * out_rs + r) = __floats2bfloat162_rn(tile[2 * tx][cc], tile[2 * tx + 1][cc]);
  }
}

// vector path of gwd_act_bwd: bf16 dy, bf16 y, no padding, 8 elements per thread, two vectors in flight.  ACT / FROM_INPUT are
// compile-time: the run-time switch sat inside the per-element loop of a kernel that is bound by instruction issue.
template <int ACT, int FROM_INPUT>
__global__ void __launch_bounds__(256)
gwd_act_bwd_vec_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y, bf16* __restrict__ out, int64_t n8, float y_mul, float scale) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + stride < n8; i += 2 * stride) {
    float d0[8], v0[8], d1[8], v1[8];
    ld8(dy + i * 8, d0);
    ld8(y + i * 8, v0);
    ld8(dy + (i + stride) * 8, d1);
    ld8(y + (i + stride) * 8, v1);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d0[e] = d0[e] * act_factor(v0[e] * y_mul, ACT, FROM_INPUT) * scale;
      d1[e] = d1[e] * act_factor(v1[e] * y_mul, ACT, FROM_INPUT) * scale;
    }
    st8(out + i * 8, d0);
    st8(out + (i + stride) * 8, d1);
  }
  for (; i < n8; i += stride) {
    float d[8], v[8];
    ld8(dy + i * 8, d);
    ld8(y + i * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = d[e] * act_factor(v[e] * y_mul, ACT, FROM_INPUT) * scale;
    st8(out + i * 8, d);
  }
}

// ------------------------------------------------------------------------------------------------
// attention backward, head_dim 32.  One CTA per (head, item): Q, K, V, dO of the head sit in shared memory as bf16
// rows of 17 words (conflict-free when lanes walk rows).  Pass A: a warp owns a query row, lanes own keys;
// recomputes the soft-max row, stores its log-sum-exp and D = sum_j P_ij dP_ij, and forms dQ_i.  Pass B: a warp owns
// a key row, lanes own queries; recomputes P and dS columns and forms dK_j, dV_j.  Cross-lane sums of the 32-wide
// rows use a 31-shuffle reduce-scatter (lane l ends up with channel l).  No atomics: results are deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kHD = 32, kRowW = 17, kBwdWarps = 8;

struct AttnBwdParams {
  const bf16 *q, *k, *v, *d_o, *o;
  bf16 *dq, *dk, *dv;
  int Lq, Lk, Lq_pad, Lk_pad;
  int64_t q_is, q_rs, k_is, k_rs, v_is, v_rs, do_is, do_rs, dq_is, dq_rs, dk_is, dk_rs, dv_is, dv_rs, o_is, o_rs;
  float scale, dq_mul, dk_mul;
  const uint32_t* drop_seed;      // dropout of the probabilities (nullptr: none); the forward's mask is regenerated
  uint32_t drop_site, drop_thr;
  float drop_scale;
  int heads;
  const uint8_t* key_padding;     // [items, Lk], 1 = padded key (src_key_padding_mask of nn.MultiheadAttention), or nullptr
  float* stats;                   // split kernels (long sequences): fp32 [items, heads, Lq, 2] = (lse, D) per query
};

__device__ __forceinline__ float reduce_scatter32(float (&acc)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = hi ? acc[i] : acc[i + o];
      const float keep = hi ? acc[i + o] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return acc[0];
}

__device__ __forceinline__ void load_row_bcast(const uint32_t* row, float (&f)[32]) {
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    float2 t = gwd_unpack_bf16x2(row[w]);
    f[2 * w] = t.x; f[2 * w + 1] = t.y;
  }
}

__device__ __forceinline__ void stage_rows(uint32_t* dst, const bf16* src, int64_t rs, int L) {
  // each row: 32 bf16 = 16 words; 16 threads per row, 4-byte loads (row starts are only 4-byte aligned in general)
  for (int i = threadIdx.x; i < L * 16; i += blockDim.x) {
    const int r = i >> 4, w = i & 15;
    dst[r * kRowW + w] = *reinterpret_cast<const uint32_t*>(src + r * rs + 2 * w);
  }
}

template <int NJ>
__global__ void __launch_bounds__(kBwdWarps * 32)
gwd_attention_bwd_kernel(AttnBwdParams p) {
  extern __shared__ uint32_t smem[];
  const int head = blockIdx.x, item = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* sQ = smem;
  uint32_t* sK = sQ + p.Lq * kRowW;
  uint32_t* sV = sK + p.Lk * kRowW;
  uint32_t* sO = sV + p.Lk * kRowW;
  float* sLse = reinterpret_cast<float*>(sO + p.Lq * kRowW);
  float* sD = sLse + p.Lq;
  const int64_t hoff = static_cast<int64_t>(head) * kHD;
  stage_rows(sQ, p.q + item * p.q_is + hoff, p.q_rs, p.Lq);
  stage_rows(sK, p.k + item * p.k_is + hoff, p.k_rs, p.Lk);
  stage_rows(sV, p.v + item * p.v_is + hoff, p.v_rs, p.Lk);
  stage_rows(sO, p.d_o + item * p.do_is + hoff, p.do_rs, p.Lq);
  __syncthreads();

  // ---- pass A: soft-max statistics and dQ
  for (int i = warp; i < p.Lq; i += kBwdWarps) {
    float qf[32], of[32];
    load_row_bcast(sQ + i * kRowW, qf);
    load_row_bcast(sO + i * kRowW, of);
    float s[NJ], dp[NJ];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = -INFINITY; dp[jj] = 0.f;
      if (j < p.Lk) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float2 kk = gwd_unpack_bf16x2(sK[j * kRowW + w]);
          const float2 vv = gwd_unpack_bf16x2(sV[j * kRowW + w]);
          a = fmaf(qf[2 * w], kk.x, a); a = fmaf(qf[2 * w + 1], kk.y, a);
          b = fmaf(of[2 * w], vv.x, b); b = fmaf(of[2 * w + 1], vv.y, b);
        }
        s[jj] = a * p.scale; dp[jj] = b;
        mx = fmaxf(mx, s[jj]);
      }
    }
    mx = gwd_warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) sum += (jj * 32 + lane < p.Lk) ? __expf(s[jj] - mx) : 0.f;
    sum = gwd_warp_sum(sum);
    const float lse = mx + __logf(sum);
    float D = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      s[jj] = (jj * 32 + lane < p.Lk) ? __expf(s[jj] - lse) : 0.f;     // P_ij
      D = fmaf(s[jj], dp[jj], D);
    }
    D = gwd_warp_sum(D);
    if (lane == 0) { sLse[i] = lse; sD[i] = D; }
    float acc[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const int j = jj * 32 + lane;
      if (j < p.Lk) {
        const float ds = s[jj] * (dp[jj] - D);
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float2 kk = gwd_unpack_bf16x2(sK[j * kRowW + w]);
          acc[2 * w] = fmaf(ds, kk.x, acc[2 * w]);
          acc[2 * w + 1] = fmaf(ds, kk.y, acc[2 * w + 1]);
        }
      }
    }
    const float r = reduce_scatter32(acc, lane) * p.dq_mul;
    p.dq[item * p.dq_is + i * p.dq_rs + hoff + lane] = __float2bfloat16(r);
  }
  __syncthreads();

  // ---- pass B: dK and dV
  for (int j = warp; j < p.Lk; j += kBwdWarps) {
    float kf[32], vf[32], ak[32], av[32];
    load_row_bcast(sK + j * kRowW, kf);
    load_row_bcast(sV + j * kRowW, vf);
#pragma unroll
    for (int d = 0; d < 32; ++d) { ak[d] = 0.f; av[d] = 0.f; }
    for (int i = lane; i < p.Lq; i += 32) {
      float qr[32];
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        const float2 qq = gwd_unpack_bf16x2(sQ[i * kRowW + w]);
        qr[2 * w] = qq.x; qr[2 * w + 1] = qq.y;
        a = fmaf(qq.x, kf[2 * w], a); a = fmaf(qq.y, kf[2 * w + 1], a);
      }
      const float pij = __expf(a * p.scale - sLse[i]);
      float b = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) {
        const float2 oo = gwd_unpack_bf16x2(sO[i * kRowW + w]);
        b = fmaf(oo.x, vf[2 * w], b); b = fmaf(oo.y, vf[2 * w + 1], b);
        av[2 * w] = fmaf(pij, oo.x, av[2 * w]);
        av[2 * w + 1] = fmaf(pij, oo.y, av[2 * w + 1]);
      }
      const float ds = pij * (b - sD[i]);
#pragma unroll
      for (int d = 0; d < 32; ++d) ak[d] = fmaf(ds, qr[d], ak[d]);
    }
    const float rk = reduce_scatter32(ak, lane) * p.dk_mul;
    const float rv = reduce_scatter32(av, lane);
    p.dk[item * p.dk_is + j * p.dk_rs + hoff + lane] = __float2bfloat16(rk);
    p.dv[item * p.dv_is + j * p.dv_rs + hoff + lane] = __float2bfloat16(rv);
  }
}


// ------------------------------------------------------------------------------------------------
// attention backward on the tensor cores (mma.sync m16n8k16, bf16 in, fp32 accumulate), head_dim 32, used when the
// forward output O is available (D_i = sum_j P_ij dP_ij = dO_i . O_i, so no separate reduction sweep is needed).
// One CTA per (head, item); Q, K, V, dO rows in shared memory with an 80-byte row stride (ldmatrix conflict-free), rows
// beyond L zero-filled up to a multiple of 64.  Two products do everything:
//   xyT : X_tile [16 x 32] . Y_chunk^T [32 x 64]   A by ldmatrix, B by ldmatrix            (S, dP, S^T, dP^T)
//   fy  : F [16 x 64] (accumulators re-packed as A fragments) . Y_chunk [64 x 32], B by ldmatrix.trans   (dQ, dK, dV)
// Pass A (a warp owns 16 queries): sweep 1 = online row max / sum of S -> log-sum-exp; sweep 2 = P, dP, dS -> dQ.
// Pass B (a warp owns 16 keys): S^T, dP^T per 64-query chunk with the per-query lse / D read from shared memory
// -> dV += P^T dO, dK += dS^T Q.  Deterministic (no atomics).
// ------------------------------------------------------------------------------------------------
constexpr int kMW = 20;      // shared-memory row stride in 32-bit words (32 bf16 + 8 padding)
constexpr float kLog2eT = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// A fragments (both k-steps: dims 0-15, 16-31) of the 16 rows starting at row0
__device__ __forceinline__ void load_a_tile(uint32_t (&a)[2][4], uint32_t base, int row0, int lane) {
  const uint32_t addr = base + static_cast<uint32_t>((row0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kMW * 4 + (lane >> 4) * 16);
  ldsm4(a[0], addr);
  ldsm4(a[1], addr + 32);
}
// acc[nt] = X_tile . Y[c0 + 8 nt .. +8]^T   (8 column tiles = 64 rows of Y)
__device__ __forceinline__ void xyT(float (&acc)[8][4], const uint32_t (&a)[2][4], uint32_t ybase, int c0, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm4(b, ybase + static_cast<uint32_t>((c0 + nt * 8 + (lane & 7)) * kMW * 4 + (lane >> 3) * 16));
    acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f;
    mma16816(acc[nt], a[0], b[0], b[1]);
    mma16816(acc[nt], a[1], b[2], b[3]);
  }
}
// out[4 dim tiles] += F [16 x 64] . Y[c0 .. c0+64][32]   with F given as fp32 accumulator tiles f[8][4]
__device__ __forceinline__ void fy(float (&out)[4][4], const float (&f)[8][4], uint32_t ybase, int c0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = gwd_pack_bf16x2(f[2 * ks][0], f[2 * ks][1]);
    a[1] = gwd_pack_bf16x2(f[2 * ks][2], f[2 * ks][3]);
    a[2] = gwd_pack_bf16x2(f[2 * ks + 1][0], f[2 * ks + 1][1]);
    a[3] = gwd_pack_bf16x2(f[2 * ks + 1][2], f[2 * ks + 1][3]);
    const uint32_t addr = ybase + static_cast<uint32_t>((c0 + ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kMW * 4 + (lane >> 4) * 16);
    uint32_t b[4];
    ldsm4t(b, addr);             // dims 0-15
    mma16816(out[0], a, b[0], b[1]);
    mma16816(out[1], a, b[2], b[3]);
    ldsm4t(b, addr + 32);        // dims 16-31
    mma16816(out[2], a, b[0], b[1]);
    mma16816(out[3], a, b[2], b[3]);
  }
}
__device__ __forceinline__ void store_tile(bf16* base, int64_t rs, int row0, int L, const float (&t)[4][4], float mul, int lane) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    if (row0 + g < L)
      *reinterpret_cast<uint32_t*>(base + (row0 + g) * rs + nt * 8 + 2 * tq) = gwd_pack_bf16x2(t[nt][0] * mul, t[nt][1] * mul);
    if (row0 + g + 8 < L)
      *reinterpret_cast<uint32_t*>(base + (row0 + g + 8) * rs + nt * 8 + 2 * tq) = gwd_pack_bf16x2(t[nt][2] * mul, t[nt][3] * mul);
  }
}

// bit j of word c: key 64 c + j exists (below Lk) and is not a padded key of this item
__device__ __forceinline__ void build_key_mask(unsigned long long* sKm, const AttnBwdParams& p, int item) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c = warp; c < p.Lk_pad / 64; c += nwarps) {
    const int k0 = c * 64 + lane, k1 = k0 + 32;
    const bool r0 = k0 < p.Lk && !(p.key_padding != nullptr && p.key_padding[static_cast<int64_t>(item) * p.Lk + k0]);
    const bool r1 = k1 < p.Lk && !(p.key_padding != nullptr && p.key_padding[static_cast<int64_t>(item) * p.Lk + k1]);
    const unsigned lo = __ballot_sync(0xffffffffu, r0), hi = __ballot_sync(0xffffffffu, r1);
    if (lane == 0) sKm[c] = (static_cast<unsigned long long>(hi) << 32) | lo;
  }
}
__device__ __forceinline__ bool key_is_real(const unsigned long long* sKm, int key) { return (sKm[key >> 6] >> (key & 63)) & 1ull; }

__global__ void __launch_bounds__(256) gwd_attention_bwd_mma_kernel(AttnBwdParams p) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  extern __shared__ __align__(16) uint32_t smem[];
  const int head = blockIdx.x, item = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t* sQ = smem;
  uint32_t* sK = sQ + p.Lq_pad * kMW;
  uint32_t* sV = sK + p.Lk_pad * kMW;
  uint32_t* sO = sV + p.Lk_pad * kMW;                 // dO
  float* sLse = reinterpret_cast<float*>(sO + p.Lq_pad * kMW);
  float* sD = sLse + p.Lq_pad;
  unsigned long long* sKm = reinterpret_cast<unsigned long long*>(sD + p.Lq_pad);     // [Lk_pad / 64] bit j of word c: key 64 c + j is real
  const int64_t hoff = static_cast<int64_t>(head) * kHD;
  build_key_mask(sKm, p, item);
  {
    const bf16* gq = p.q + item * p.q_is + hoff;
    const bf16* gk = p.k + item * p.k_is + hoff;
    const bf16* gv = p.v + item * p.v_is + hoff;
    const bf16* gdo = p.d_o + item * p.do_is + hoff;
    const bf16* go = p.o + item * p.o_is + hoff;
    for (int i = threadIdx.x; i < p.Lk_pad * 16; i += 256) {
      const int r = i >> 4, w = i & 15;
      const bool in = r < p.Lk;
      sK[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gk + r * p.k_rs + 2 * w) : 0u;
      sV[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gv + r * p.v_rs + 2 * w) : 0u;
    }
    for (int i = threadIdx.x; i < p.Lq_pad * 16; i += 256) {      // 16 consecutive lanes hold one row
      const int r = i >> 4, w = i & 15;
      const bool in = r < p.Lq;
      const uint32_t dov = in ? *reinterpret_cast<const uint32_t*>(gdo + r * p.do_rs + 2 * w) : 0u;
      const uint32_t ov = in ? *reinterpret_cast<const uint32_t*>(go + r * p.o_rs + 2 * w) : 0u;
      sQ[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gq + r * p.q_rs + 2 * w) : 0u;
      sO[r * kMW + w] = dov;
      const float2 a = gwd_unpack_bf16x2(dov), b = gwd_unpack_bf16x2(ov);
      float d = a.x * b.x + a.y * b.y;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (w == 0) { sD[r] = d; sLse[r] = INFINITY; }     // lse of real rows is written by pass A; +inf = "no such query"
    }
  }
  __syncthreads();
  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV), bO = smem_u32(sO);
  const float sc = p.scale * kLog2eT;
  // dropout: O = (P o M / (1 - p)) V, so dV sees the dropped P, dP = M / (1 - p) o (dO V^T), and D = rowsum(dO o O) still equals
  // rowsum(P o dP); the mask of (query, key) is regenerated from the forward's (seed, site, image, head)
  const bool drop = p.drop_seed != nullptr;
  const uint32_t drop_key = drop ? gwd_drop_key(*p.drop_seed, p.drop_site, static_cast<uint32_t>(item * p.heads + head)) : 0u;

  // ---------------- pass A
  for (int m0 = warp * 16; m0 < p.Lq; m0 += 8 * 16) {
    uint32_t qa[2][4], oa[2][4];
    load_a_tile(qa, bQ, m0, lane);
    load_a_tile(oa, bO, m0, lane);
    float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    for (int c0 = 0; c0 < p.Lk_pad; c0 += 64) {
      float s[8][4];
      xyT(s, qa, bK, c0, lane);
      float cm[2] = {-INFINITY, -INFINITY};
      const unsigned long long km = sKm[c0 >> 6] >> (2 * tq);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!((km >> (nt * 8 + (e & 1))) & 1ull)) s[nt][e] = -INFINITY;     // beyond Lk or a padded key
          cm[e >> 1] = fmaxf(cm[e >> 1], s[nt][e]);
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float mn = fmaxf(mx[h], cm[h]);
        if (mn > -INFINITY) {
          float add = 0.f;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) add += ex2f((s[nt][2 * h] - mn) * sc) + ex2f((s[nt][2 * h + 1] - mn) * sc);
          l[h] = l[h] * ex2f((mx[h] - mn) * sc) + add;
          mx[h] = mn;
        }
      }
    }
    float lse[2], D[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float ma = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
      float lh = l[h] * ex2f((mx[h] - ma) * sc);           // a lane that saw no valid column has mx = -inf -> factor 0
      lh += __shfl_xor_sync(0xffffffffu, lh, 1);
      lh += __shfl_xor_sync(0xffffffffu, lh, 2);
      lse[h] = ma * sc + __log2f(lh);
      const int row = m0 + g + 8 * h;
      D[h] = sD[row];
      if (tq == 0 && row < p.Lq) sLse[row] = lse[h];
    }
    float dq[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { dq[nt][0] = 0.f; dq[nt][1] = 0.f; dq[nt][2] = 0.f; dq[nt][3] = 0.f; }
    for (int c0 = 0; c0 < p.Lk_pad; c0 += 64) {
      float s[8][4], dp[8][4];
      xyT(s, qa, bK, c0, lane);
      xyT(dp, oa, bV, c0, lane);
      const unsigned long long km = sKm[c0 >> 6] >> (2 * tq);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = c0 + nt * 8 + 2 * tq + (e & 1);
          const float pr = ((km >> (nt * 8 + (e & 1))) & 1ull) ? ex2f(s[nt][e] * sc - lse[e >> 1]) : 0.f;
          float dpe = dp[nt][e];
          if (drop) {
            const uint32_t row = static_cast<uint32_t>(m0 + g + 8 * (e >> 1));
            dpe = gwd_drop_keep(drop_key, row * static_cast<uint32_t>(p.Lk) + col, p.drop_thr) ? dpe * p.drop_scale : 0.f;
          }
          s[nt][e] = pr * (dpe - D[e >> 1]);           // dS
        }
      fy(dq, s, bK, c0, lane);
    }
    store_tile(p.dq + item * p.dq_is + hoff, p.dq_rs, m0, p.Lq, dq, p.dq_mul, lane);
  }
  __syncthreads();

  // ---------------- pass B
  for (int n0 = warp * 16; n0 < p.Lk; n0 += 8 * 16) {
    uint32_t ka[2][4], va[2][4];
    load_a_tile(ka, bK, n0, lane);
    load_a_tile(va, bV, n0, lane);
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { dk[nt][e] = 0.f; dv[nt][e] = 0.f; }
    // a padded key took no part in the forward: its column of P is zero
    const float kvalid[2] = {key_is_real(sKm, n0 + g) ? 1.f : 0.f, key_is_real(sKm, n0 + g + 8) ? 1.f : 0.f};
    for (int c0 = 0; c0 < p.Lq_pad; c0 += 64) {
      float st[8][4], dpt[8][4];
      xyT(st, ka, bQ, c0, lane);
      xyT(dpt, va, bO, c0, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 ls = *reinterpret_cast<const float2*>(sLse + c0 + nt * 8 + 2 * tq);
        const float2 dd = *reinterpret_cast<const float2*>(sD + c0 + nt * 8 + 2 * tq);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pr = ex2f(st[nt][e] * sc - ((e & 1) ? ls.y : ls.x)) * kvalid[e >> 1];        // lse = +inf beyond Lq -> 0
          float mk = 1.f;
          if (drop) {
            const uint32_t qrow = static_cast<uint32_t>(c0 + nt * 8 + 2 * tq + (e & 1)), key = static_cast<uint32_t>(n0 + g + 8 * (e >> 1));
            mk = gwd_drop_keep(drop_key, qrow * static_cast<uint32_t>(p.Lk) + key, p.drop_thr) ? p.drop_scale : 0.f;
          }
          st[nt][e] = pr * mk;                                                    // (dropped) P^T: what met V in the forward
          dpt[nt][e] = pr * (dpt[nt][e] * mk - ((e & 1) ? dd.y : dd.x));          // dS^T
        }
      }
      fy(dv, st, bO, c0, lane);
      fy(dk, dpt, bQ, c0, lane);
    }
    store_tile(p.dk + item * p.dk_is + hoff, p.dk_rs, n0, p.Lk, dk, p.dk_mul, lane);
    store_tile(p.dv + item * p.dv_is + hoff, p.dv_rs, n0, p.Lk, dv, 1.f, lane);
  }
}


// ------------------------------------------------------------------------------------------------
// The same backward for LONG sequences (Lq or Lk in 513 .. 1 280: 800 x 1 024 and 960 x 1 280 images at 1/32), where Q, K, V and dO
// of a head no longer fit the shared memory of one CTA together.  Two launches:
//   _q  : CTA = (head, item, 128 queries); K, V of the head resident, its Q / dO / O rows -> log-sum-exp, D, dQ; (lse, D) -> stats
//   _kv : CTA = (head, item, 64 keys); Q, dO and the (lse, D) of all queries resident, its K / V rows -> dK, dV
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gwd_attention_bwd_mma_q_kernel(AttnBwdParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int head = blockIdx.x, item = blockIdx.y, q0 = blockIdx.z * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t* sK = smem;
  uint32_t* sV = sK + p.Lk_pad * kMW;
  uint32_t* sQ = sV + p.Lk_pad * kMW;                  // [128]
  uint32_t* sO = sQ + 128 * kMW;                       // dO [128]
  float* sD = reinterpret_cast<float*>(sO + 128 * kMW);
  unsigned long long* sKm = reinterpret_cast<unsigned long long*>(sD + 128);
  const int64_t hoff = static_cast<int64_t>(head) * kHD;
  build_key_mask(sKm, p, item);
  {
    const bf16* gq = p.q + item * p.q_is + hoff;
    const bf16* gk = p.k + item * p.k_is + hoff;
    const bf16* gv = p.v + item * p.v_is + hoff;
    const bf16* gdo = p.d_o + item * p.do_is + hoff;
    const bf16* go = p.o + item * p.o_is + hoff;
    for (int i = threadIdx.x; i < p.Lk_pad * 16; i += 256) {
      const int r = i >> 4, w = i & 15;
      const bool in = r < p.Lk;
      sK[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gk + r * p.k_rs + 2 * w) : 0u;
      sV[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gv + r * p.v_rs + 2 * w) : 0u;
    }
    for (int i = threadIdx.x; i < 128 * 16; i += 256) {
      const int r = i >> 4, w = i & 15;
      const int64_t gr = q0 + r;
      const bool in = gr < p.Lq;
      const uint32_t dov = in ? *reinterpret_cast<const uint32_t*>(gdo + gr * p.do_rs + 2 * w) : 0u;
      const uint32_t ov = in ? *reinterpret_cast<const uint32_t*>(go + gr * p.o_rs + 2 * w) : 0u;
      sQ[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gq + gr * p.q_rs + 2 * w) : 0u;
      sO[r * kMW + w] = dov;
      const float2 a = gwd_unpack_bf16x2(dov), b = gwd_unpack_bf16x2(ov);
      float d = a.x * b.x + a.y * b.y;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (w == 0) sD[r] = d;
    }
  }
  __syncthreads();
  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV), bO = smem_u32(sO);
  const float sc = p.scale * kLog2eT;
  const bool drop = p.drop_seed != nullptr;
  const uint32_t drop_key = drop ? gwd_drop_key(*p.drop_seed, p.drop_site, static_cast<uint32_t>(item * p.heads + head)) : 0u;
  const int ml = warp * 16, m0 = q0 + ml;              // local / global first query of this warp
  if (m0 >= p.Lq) return;
  uint32_t qa[2][4], oa[2][4];
  load_a_tile(qa, bQ, ml, lane);
  load_a_tile(oa, bO, ml, lane);
  float mx[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  for (int c0 = 0; c0 < p.Lk_pad; c0 += 64) {
    float s[8][4];
    xyT(s, qa, bK, c0, lane);
    float cm[2] = {-INFINITY, -INFINITY};
    const unsigned long long km = sKm[c0 >> 6] >> (2 * tq);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!((km >> (nt * 8 + (e & 1))) & 1ull)) s[nt][e] = -INFINITY;
        cm[e >> 1] = fmaxf(cm[e >> 1], s[nt][e]);
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float mn = fmaxf(mx[h], cm[h]);
      if (mn > -INFINITY) {
        float add = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) add += ex2f((s[nt][2 * h] - mn) * sc) + ex2f((s[nt][2 * h + 1] - mn) * sc);
        l[h] = l[h] * ex2f((mx[h] - mn) * sc) + add;
        mx[h] = mn;
      }
    }
  }
  float lse[2], D[2];
  float* st_out = p.stats + (static_cast<int64_t>(item) * p.heads + head) * p.Lq * 2;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float ma = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    float lh = l[h] * ex2f((mx[h] - ma) * sc);
    lh += __shfl_xor_sync(0xffffffffu, lh, 1);
    lh += __shfl_xor_sync(0xffffffffu, lh, 2);
    lse[h] = ma * sc + __log2f(lh);
    const int rl = ml + g + 8 * h;
    D[h] = sD[rl];
    if (tq == 0 && q0 + rl < p.Lq) {
      st_out[(q0 + rl) * 2] = lse[h];
      st_out[(q0 + rl) * 2 + 1] = D[h];
    }
  }
  float dq[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { dq[nt][0] = 0.f; dq[nt][1] = 0.f; dq[nt][2] = 0.f; dq[nt][3] = 0.f; }
  for (int c0 = 0; c0 < p.Lk_pad; c0 += 64) {
    float s[8][4], dp[8][4];
    xyT(s, qa, bK, c0, lane);
    xyT(dp, oa, bV, c0, lane);
    const unsigned long long km = sKm[c0 >> 6] >> (2 * tq);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = c0 + nt * 8 + 2 * tq + (e & 1);
        const float pr = ((km >> (nt * 8 + (e & 1))) & 1ull) ? ex2f(s[nt][e] * sc - lse[e >> 1]) : 0.f;
        float dpe = dp[nt][e];
        if (drop) {
          const uint32_t row = static_cast<uint32_t>(m0 + g + 8 * (e >> 1));
          dpe = gwd_drop_keep(drop_key, row * static_cast<uint32_t>(p.Lk) + col, p.drop_thr) ? dpe * p.drop_scale : 0.f;
        }
        s[nt][e] = pr * (dpe - D[e >> 1]);
      }
    fy(dq, s, bK, c0, lane);
  }
  store_tile(p.dq + item * p.dq_is + hoff, p.dq_rs, m0, p.Lq, dq, p.dq_mul, lane);
}

__global__ void __launch_bounds__(128) gwd_attention_bwd_mma_kv_kernel(AttnBwdParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int head = blockIdx.x, item = blockIdx.y, k0 = blockIdx.z * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t* sQ = smem;
  uint32_t* sO = sQ + p.Lq_pad * kMW;                  // dO
  uint32_t* sK = sO + p.Lq_pad * kMW;                  // [64]
  uint32_t* sV = sK + 64 * kMW;
  float* sLse = reinterpret_cast<float*>(sV + 64 * kMW);
  float* sD = sLse + p.Lq_pad;
  const int64_t hoff = static_cast<int64_t>(head) * kHD;
  {
    const bf16* gq = p.q + item * p.q_is + hoff;
    const bf16* gk = p.k + item * p.k_is + hoff;
    const bf16* gv = p.v + item * p.v_is + hoff;
    const bf16* gdo = p.d_o + item * p.do_is + hoff;
    const float* st_in = p.stats + (static_cast<int64_t>(item) * p.heads + head) * p.Lq * 2;
    for (int i = threadIdx.x; i < 64 * 16; i += 128) {
      const int r = i >> 4, w = i & 15;
      const int64_t gr = k0 + r;
      const bool in = gr < p.Lk;
      sK[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gk + gr * p.k_rs + 2 * w) : 0u;
      sV[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gv + gr * p.v_rs + 2 * w) : 0u;
    }
    for (int i = threadIdx.x; i < p.Lq_pad * 16; i += 128) {
      const int r = i >> 4, w = i & 15;
      const bool in = r < p.Lq;
      sQ[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gq + r * p.q_rs + 2 * w) : 0u;
      sO[r * kMW + w] = in ? *reinterpret_cast<const uint32_t*>(gdo + r * p.do_rs + 2 * w) : 0u;
    }
    for (int r = threadIdx.x; r < p.Lq_pad; r += 128) {
      sLse[r] = r < p.Lq ? st_in[2 * r] : INFINITY;          // +inf = "no such query": its P is 0
      sD[r] = r < p.Lq ? st_in[2 * r + 1] : 0.f;
    }
  }
  __syncthreads();
  const uint32_t bQ = smem_u32(sQ), bK = smem_u32(sK), bV = smem_u32(sV), bO = smem_u32(sO);
  const float sc = p.scale * kLog2eT;
  const bool drop = p.drop_seed != nullptr;
  const uint32_t drop_key = drop ? gwd_drop_key(*p.drop_seed, p.drop_site, static_cast<uint32_t>(item * p.heads + head)) : 0u;
  const int nl = warp * 16, n0 = k0 + nl;
  if (n0 >= p.Lk) return;
  uint32_t ka[2][4], va[2][4];
  load_a_tile(ka, bK, nl, lane);
  load_a_tile(va, bV, nl, lane);
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[nt][e] = 0.f; dv[nt][e] = 0.f; }
  float kvalid[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int key = n0 + g + 8 * h;
    kvalid[h] = (key < p.Lk && !(p.key_padding != nullptr && p.key_padding[static_cast<int64_t>(item) * p.Lk + key])) ? 1.f : 0.f;
  }
  for (int c0 = 0; c0 < p.Lq_pad; c0 += 64) {
    float st[8][4], dpt[8][4];
    xyT(st, ka, bQ, c0, lane);
    xyT(dpt, va, bO, c0, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 ls = *reinterpret_cast<const float2*>(sLse + c0 + nt * 8 + 2 * tq);
      const float2 dd = *reinterpret_cast<const float2*>(sD + c0 + nt * 8 + 2 * tq);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pr = ex2f(st[nt][e] * sc - ((e & 1) ? ls.y : ls.x)) * kvalid[e >> 1];
        float mk = 1.f;
        if (drop) {
          const uint32_t qrow = static_cast<uint32_t>(c0 + nt * 8 + 2 * tq + (e & 1)), key = static_cast<uint32_t>(n0 + g + 8 * (e >> 1));
          mk = gwd_drop_keep(drop_key, qrow * static_cast<uint32_t>(p.Lk) + key, p.drop_thr) ? p.drop_scale : 0.f;
        }
        st[nt][e] = pr * mk;
        dpt[nt][e] = pr * (dpt[nt][e] * mk - ((e & 1) ? dd.y : dd.x));
      }
    }
    fy(dv, st, bO, c0, lane);
    fy(dk, dpt, bQ, c0, lane);
  }
  store_tile(p.dk + item * p.dk_is + hoff, p.dk_rs, n0, p.Lk, dk, p.dk_mul, lane);
  store_tile(p.dv + item * p.dv_is + hoff, p.dv_rs, n0, p.Lk, dv, 1.f, lane);
}


// ------------------------------------------------------------------------------------------------
// Linear weight gradient  dW[n, k] += sum_r dY[r, n] X[r, k]  (+ db[n] += sum_r dY[r, n]).
// Both operands are contracted over their SLOW axis (the token rows), i.e. they are "MN-major" for a GEMM: on
// mma.sync that is simply ldmatrix.trans on both fragments, so no transposed copies of dY and X are made.  The
// problem is short and fat (N x K up to 2048 x 256 outputs, R = 1 600 .. 9 600 rows), so the rows are split across
// CTAs (split-K) until ~2 CTAs per SM exist, and partial tiles are reduced with vector fp32 atomics into the
// (pre-zeroed) flat gradient buffer.  CTA tile 128 (n) x 128 (k) x 32 rows per stage, 3-stage cp.async ring,
// 8 warps of 32 x 64.
// ------------------------------------------------------------------------------------------------
constexpr int kWgT = 128, kWgR = 32, kWgLd = 136, kWgStages = 3;

struct WgradParams {
  const bf16* dy; const bf16* x;
  float* dw; float* db;
  int64_t dy_rs, x_rs, dw_rs;
  int R, N, K, rows_per_split;
  int H, W, sy, sx;      // H > 0: rows are pixels of [B,H,W] maps and X is read at (y + sy, x + sx), zero outside the map
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz));
}

__global__ void __launch_bounds__(256) gwd_wgrad_kernel(WgradParams p) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  extern __shared__ __align__(16) uint8_t wsm[];
  bf16* sA = reinterpret_cast<bf16*>(wsm);                          // [stages][32][136]  dY tile
  bf16* sB = sA + kWgStages * kWgR * kWgLd;                         // [stages][32][136]  X tile
  const int n0 = blockIdx.x * kWgT, k0 = blockIdx.y * kWgT;
  const int r_begin = blockIdx.z * p.rows_per_split;
  const int r_end = min(p.R, r_begin + p.rows_per_split);
  const int iters = (r_end - r_begin + kWgR - 1) / kWgR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB);

  auto load_stage = [&](int it, int st) {
    const int r0 = r_begin + it * kWgR;
#pragma unroll
    for (int c = threadIdx.x; c < kWgR * 16; c += 256) {
      const int rr = c >> 4, c8 = (c & 15) * 8;
      const int r = r0 + rr;
      const bool rin = r < r_end;
      const bool va = rin && (n0 + c8 < p.N);
      bool vb = rin && (k0 + c8 < p.K);
      int64_t xr = r;
      if (p.H > 0 && vb) {      // one tap of a 3x3 convolution: the input pixel this output pixel saw through that tap
        const int x = r % p.W, y = (r / p.W) % p.H;
        vb = (y + p.sy >= 0) && (y + p.sy < p.H) && (x + p.sx >= 0) && (x + p.sx < p.W);
        xr = r + p.sy * p.W + p.sx;
      }
      cp_async16(aBase + static_cast<uint32_t>(((st * kWgR + rr) * kWgLd + c8) * 2), va ? p.dy + static_cast<int64_t>(r) * p.dy_rs + n0 + c8 : p.dy, va);
      cp_async16(bBase + static_cast<uint32_t>(((st * kWgR + rr) * kWgLd + c8) * 2), vb ? p.x + xr * p.x_rs + k0 + c8 : p.x, vb);
    }
  };

  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[mt][nt][0] = 0.f; acc[mt][nt][1] = 0.f; acc[mt][nt][2] = 0.f; acc[mt][nt][3] = 0.f; }
  float bsum = 0.f;
  const bool do_bias = p.db != nullptr && blockIdx.y == 0 && threadIdx.x < kWgT;

#pragma unroll
  for (int s = 0; s < kWgStages - 1; ++s) {
    if (s < iters) load_stage(s, s);
    asm volatile("cp.async.commit_group;");
  }
  for (int it = 0; it < iters; ++it) {
    asm volatile("cp.async.wait_group %0;" :: "n"(kWgStages - 2));
    __syncthreads();
    if (it + kWgStages - 1 < iters) load_stage(it + kWgStages - 1, (it + kWgStages - 1) % kWgStages);
    asm volatile("cp.async.commit_group;");
    const int st = it % kWgStages;
    const uint32_t aS = aBase + static_cast<uint32_t>(st * kWgR * kWgLd * 2), bS = bBase + static_cast<uint32_t>(st * kWgR * kWgLd * 2);
#pragma unroll
    for (int kk = 0; kk < kWgR / 16; ++kk) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)       // A = dY^T: matrices (k+0,m+0) (k+0,m+8) (k+8,m+0) (k+8,m+8), transposed on load
        ldsm4t(a[mt], aS + static_cast<uint32_t>(((kk * 16 + (lane & 7) + 8 * (lane >> 4)) * kWgLd + wm * 32 + mt * 16 + 8 * ((lane >> 3) & 1)) * 2));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        ldsm4t(b, bS + static_cast<uint32_t>(((kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kWgLd + wn * 64 + np * 16 + 8 * (lane >> 4)) * 2));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * np], a[mt], b[0], b[1]);
          mma16816(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
        }
      }
    }
    if (do_bias) {
      const bf16* col = sA + st * kWgR * kWgLd + threadIdx.x;
#pragma unroll 8
      for (int rr = 0; rr < kWgR; ++rr) bsum += __bfloat162float(col[rr * kWgLd]);
    }
  }
  asm volatile("cp.async.wait_group 0;");
  const bool single = gridDim.z == 1;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int kcol = k0 + wn * 64 + nt * 8 + 2 * tq;
      if (kcol >= p.K) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = n0 + wm * 32 + mt * 16 + g + 8 * h;
        if (n >= p.N) continue;
        float2* dst = reinterpret_cast<float2*>(p.dw + static_cast<int64_t>(n) * p.dw_rs + kcol);
        const float2 v = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
        if (single) { float2 o = *dst; o.x += v.x; o.y += v.y; *dst = o; }
        else atomicAdd(dst, v);
      }
    }
  if (do_bias && n0 + threadIdx.x < p.N) atomicAdd(p.db + n0 + threadIdx.x, bsum);
}


// ------------------------------------------------------------------------------------------------
// 3x3 weight gradient for NARROW convolutions (N, C <= 64: the dense prediction head at 1/2 and full resolution, where
// the contraction runs over millions of pixels): dW[tap][n][c] += sum_p dY[p][n] X[p + shift(tap)][c] for all 9 taps in
// ONE pass over dY and X.  The tap-by-tap kernel above reads both maps 9 times and fills 1/16 of its 128 x 128 tile here.
// Persistent CTAs of 9 warps walk 4 x 64 pixel tiles: the dY tile and the X tile with its 1-pixel halo (zero-filled
// outside the image) arrive by cp.async into a double buffer (pixel rows padded by 8 channels: conflict-free ldmatrix);
// warp t owns tap t and keeps its whole N x C accumulator in registers across all tiles of the CTA (mma.sync m16n8k16,
// 16 consecutive pixels of a row = one K step, both operands by ldmatrix.trans, the X fragment simply addressed at the
// shifted pixel), so the only global writes are 9 N C vector atomics per CTA at the end.
// ------------------------------------------------------------------------------------------------
constexpr int kWsTH = 4, kWsTW = 64, kWsThreads = 288;

template <int NP, int CP>
__global__ void __launch_bounds__(kWsThreads)
gwd_conv3x3_wgrad_small_kernel(const bf16* __restrict__ dy, int64_t dy_cs, const bf16* __restrict__ x, int64_t x_cs, int B, int H,
                               int W, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int LDA = NP + 8, LDB = CP + 8;
  constexpr int kTileA = kWsTH * kWsTW * LDA, kTileB = (kWsTH + 2) * (kWsTW + 2) * LDB;
  extern __shared__ __align__(16) uint8_t wsm[];
  bf16* sA = reinterpret_cast<bf16*>(wsm);           // [2][TH*TW][LDA]           dY
  bf16* sB = sA + 2 * kTileA;                         // [2][(TH+2)*(TW+2)][LDB]   X with halo
  const int lane = threadIdx.x & 31, tap = threadIdx.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const int sx = tap / 3 - 1, sy = tap % 3 - 1;      // tap = dx * 3 + dy
  const int tiles_x = (W + kWsTW - 1) / kWsTW, tiles_y = (H + kWsTH - 1) / kWsTH;
  const int64_t tiles = static_cast<int64_t>(B) * tiles_y * tiles_x;

  auto load_tile = [&](int64_t t, int buf) {
    const int tx = static_cast<int>(t % tiles_x), ty = static_cast<int>((t / tiles_x) % tiles_y), b = static_cast<int>(t / (tiles_x * tiles_y));
    const int x0 = tx * kWsTW, y0 = ty * kWsTH;
    const uint32_t aB = smem_u32(sA + buf * kTileA), bB = smem_u32(sB + buf * kTileB);
    for (int i = threadIdx.x; i < kWsTH * kWsTW * (NP / 8); i += kWsThreads) {
      const int c8 = (i % (NP / 8)) * 8, pix = i / (NP / 8);
      const int px = x0 + pix % kWsTW, py = y0 + pix / kWsTW;
      const bool v = px < W && py < H;
      cp_async16(aB + static_cast<uint32_t>((pix * LDA + c8) * 2), v ? dy + ((static_cast<int64_t>(b) * H + py) * W + px) * dy_cs + c8 : dy, v);
    }
    for (int i = threadIdx.x; i < (kWsTH + 2) * (kWsTW + 2) * (CP / 8); i += kWsThreads) {
      const int c8 = (i % (CP / 8)) * 8, pix = i / (CP / 8);
      const int px = x0 - 1 + pix % (kWsTW + 2), py = y0 - 1 + pix / (kWsTW + 2);
      const bool v = px >= 0 && px < W && py >= 0 && py < H;
      cp_async16(bB + static_cast<uint32_t>((pix * LDB + c8) * 2), v ? x + ((static_cast<int64_t>(b) * H + py) * W + px) * x_cs + c8 : x, v);
    }
  };

  float acc[NP / 16][CP / 8][4];
#pragma unroll
  for (int mt = 0; mt < NP / 16; ++mt)
#pragma unroll
    for (int nt = 0; nt < CP / 8; ++nt) { acc[mt][nt][0] = 0.f; acc[mt][nt][1] = 0.f; acc[mt][nt][2] = 0.f; acc[mt][nt][3] = 0.f; }
  float bsum = 0.f;
  const int bch = threadIdx.x % NP, bgrp = threadIdx.x / NP;       // bias gradient: channel bch over every (288 / NP)-th pixel

  int64_t t = blockIdx.x;
  if (t < tiles) load_tile(t, 0);
  asm volatile("cp.async.commit_group;");
  int buf = 0;
  for (; t < tiles; t += gridDim.x, buf ^= 1) {
    if (t + gridDim.x < tiles) load_tile(t + gridDim.x, buf ^ 1);
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 1;");
    __syncthreads();
    const uint32_t aS = smem_u32(sA + buf * kTileA), bS = smem_u32(sB + buf * kTileB);
#pragma unroll 1
    for (int ry = 0; ry < kWsTH; ++ry) {
#pragma unroll
      for (int kx = 0; kx < kWsTW / 16; ++kx) {
        // A = dY^T (n x pixels): matrices (k+0,m+0) (k+0,m+8) (k+8,m+0) (k+8,m+8), transposed on load
        const int pa = ry * kWsTW + kx * 16 + (lane & 7) + 8 * (lane >> 4);
        // B = X (pixels x c) at the pixel this output pixel saw through the tap (halo origin = (-1, -1))
        const int pb = (ry + 1 + sy) * (kWsTW + 2) + kx * 16 + 1 + sx + (lane & 7) + 8 * ((lane >> 3) & 1);
        uint32_t a[NP / 16][4];
#pragma unroll
        for (int mt = 0; mt < NP / 16; ++mt) ldsm4t(a[mt], aS + static_cast<uint32_t>((pa * LDA + mt * 16 + 8 * ((lane >> 3) & 1)) * 2));
#pragma unroll
        for (int np = 0; np < CP / 16; ++np) {
          uint32_t bfr[4];
          ldsm4t(bfr, bS + static_cast<uint32_t>((pb * LDB + np * 16 + 8 * (lane >> 4)) * 2));
#pragma unroll
          for (int mt = 0; mt < NP / 16; ++mt) {
            mma16816(acc[mt][2 * np], a[mt], bfr[0], bfr[1]);
            mma16816(acc[mt][2 * np + 1], a[mt], bfr[2], bfr[3]);
          }
        }
      }
    }
    if (db != nullptr && bgrp < kWsThreads / NP) {
      const bf16* col = sA + buf * kTileA + bch;
      for (int pix = bgrp; pix < kWsTH * kWsTW; pix += kWsThreads / NP) bsum += __bfloat162float(col[pix * LDA]);
    }
    __syncthreads();       // the next iteration's prefetch overwrites this buffer's twin only; this one is refilled after it
  }
  asm volatile("cp.async.wait_group 0;");
  float* out = dw + static_cast<int64_t>(tap) * NP * CP;
#pragma unroll
  for (int mt = 0; mt < NP / 16; ++mt)
#pragma unroll
    for (int nt = 0; nt < CP / 8; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        atomicAdd(reinterpret_cast<float2*>(out + (mt * 16 + g + 8 * h) * CP + nt * 8 + 2 * tq), make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]));
  if (db != nullptr && bgrp < kWsThreads / NP && blockIdx.x < tiles) atomicAdd(db + bch, bsum);
}

template <int NP, int CP>
static int launch_wgrad_small(const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, int B, int H, int W, float* dw, float* db,
                              cudaStream_t stream) {
  constexpr size_t smem = 2 * (static_cast<size_t>(kWsTH) * kWsTW * (NP + 8) + (kWsTH + 2) * (kWsTW + 2) * (CP + 8)) * sizeof(bf16);
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_conv3x3_wgrad_small_kernel<NP, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int n = 0;
    GWD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gwd_conv3x3_wgrad_small_kernel<NP, CP>, kWsThreads, smem));
    ctas_per_sm = std::max(1, n);
  }
  const int64_t tiles = static_cast<int64_t>(B) * gwd_ceil_div(H, kWsTH) * gwd_ceil_div(W, kWsTW);
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(tiles, static_cast<int64_t>(ctas_per_sm) * gwd_num_sms()));
  gwd_conv3x3_wgrad_small_kernel<NP, CP><<<grid, kWsThreads, smem, stream>>>(static_cast<const bf16*>(dy), dy_cs, static_cast<const bf16*>(x),
                                                                             x_cs, B, H, W, dw, db);
  GWD_LAUNCHED();
  return GWD_OK;
}

// ------------------------------------------------------------------------------------------------
// Set criterion, forward AND backward in one launch (src/models/glassrgbd.py:154-175 weighted cross entropy over
// {line, no-object}, :231-244 L1 over matched pairs / num_items) for all S decoder stages: one CTA per stage.
//   loss_ce[s]   = sum_i w[c_i] nll_i / sum_i w[c_i]         c_i = label of the target matched to query i, else C-1
//   loss_line[s] = sum_{matched} |line - target|_1 / num_items
//   dlogits = w_ce[s] w[c_i] / W_s (softmax - onehot),   dlines = w_line[s] sign(line - target) / num_items (0 unmatched)
// The assignment (stage, image, query, target) comes from the Hungarian solve on the host as int32 columns.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;

struct SetLossParams {
  const float* logits; const float* lines; const float* tgt_lines; const int64_t* tgt_labels;
  const int32_t* match;        // [4][M]: stage, image, query, target (rows of the concatenated targets)
  const int32_t* stage_off;    // [S+1] ranges of `match` per stage
  const float* class_w;        // [C]
  const float* w_ce; const float* w_line;   // [S]
  const float* num_items;      // [1] (device: all-reduced over ranks by the caller)
  float* losses;               // [S][2] = loss_ce, loss_line
  float* dlogits; float* dlines;
  int S, BQ, Q, C, D, M;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = gwd_warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256) gwd_set_loss_kernel(SetLossParams p) {
  extern __shared__ uint8_t cls[];       // [B*Q]
  __shared__ float red[8];
  const int s = blockIdx.x;
  const float inv_items = 1.f / p.num_items[0];
  for (int i = threadIdx.x; i < p.BQ; i += blockDim.x) cls[i] = static_cast<uint8_t>(p.C - 1);
  float* dl = p.dlines + static_cast<int64_t>(s) * p.BQ * p.D;
  for (int i = threadIdx.x; i < p.BQ * p.D; i += blockDim.x) dl[i] = 0.f;
  __syncthreads();
  float l1 = 0.f;
  const float wl = p.w_line[s] * inv_items;
  for (int m = p.stage_off[s] + threadIdx.x; m < p.stage_off[s + 1]; m += blockDim.x) {
    const int b = p.match[p.M + m], q = p.match[2 * p.M + m], t = p.match[3 * p.M + m];
    cls[b * p.Q + q] = static_cast<uint8_t>(p.tgt_labels[t]);
    const float* src = p.lines + (static_cast<int64_t>(s) * p.BQ + b * p.Q + q) * p.D;
    const float* tg = p.tgt_lines + static_cast<int64_t>(t) * p.D;
    for (int d = 0; d < p.D; ++d) {
      const float diff = src[d] - tg[d];
      l1 += fabsf(diff);
      dl[(b * p.Q + q) * p.D + d] = diff > 0.f ? wl : (diff < 0.f ? -wl : 0.f);
    }
  }
  __syncthreads();
  float wn = 0.f, wsum = 0.f;
  const float* lg = p.logits + static_cast<int64_t>(s) * p.BQ * p.C;
  for (int i = threadIdx.x; i < p.BQ; i += blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < p.C; ++c) mx = fmaxf(mx, lg[i * p.C + c]);
    float se = 0.f;
    for (int c = 0; c < p.C; ++c) se += expf(lg[i * p.C + c] - mx);
    const int c = cls[i];
    const float w = p.class_w[c];
    wn += w * (mx + logf(se) - lg[i * p.C + c]);
    wsum += w;
  }
  const float l1_all = block_sum(l1, red);
  const float wn_all = block_sum(wn, red);
  const float W = block_sum(wsum, red);
  if (threadIdx.x == 0) {
    p.losses[2 * s] = wn_all / W;
    p.losses[2 * s + 1] = l1_all * inv_items;
  }
  float* dg = p.dlogits + static_cast<int64_t>(s) * p.BQ * p.C;
  const float k = p.w_ce[s] / W;
  for (int i = threadIdx.x; i < p.BQ; i += blockDim.x) {
    float mx = -INFINITY;
    for (int c = 0; c < p.C; ++c) mx = fmaxf(mx, lg[i * p.C + c]);
    float se = 0.f;
    for (int c = 0; c < p.C; ++c) se += expf(lg[i * p.C + c] - mx);
    const int ci = cls[i];
    const float kw = k * p.class_w[ci];
    for (int c = 0; c < p.C; ++c) dg[i * p.C + c] = kw * (expf(lg[i * p.C + c] - mx) / se - (c == ci ? 1.f : 0.f));
  }
}

// ------------------------------------------------------------------------------------------------
// optimizer
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gwd_sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 v = g4[i];
    acc += static_cast<double>(v.x * v.x + v.y * v.y) + static_cast<double>(v.z * v.z + v.w * v.w);
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
    acc += static_cast<double>(g[i]) * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w];
    atomicAdd(out, t);
  }
}

struct AdamParams {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm, grad_scale;
};

// torch.optim.AdamW (decoupled weight decay, no amsgrad) after torch.nn.utils.clip_grad_norm_:
//   g <- g * grad_scale * min(1, max_norm / (||g * grad_scale|| + 1e-6));  p <- p (1 - lr wd);
//   m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;  p <- p - (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
gwd_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 bf16* __restrict__ mirror, int64_t n, AdamParams a, const double* __restrict__ sumsq) {
  float gs = a.grad_scale;
  if (a.max_norm > 0.f && sumsq != nullptr) {
    const float total = static_cast<float>(sqrt(*sumsq)) * a.grad_scale;
    gs *= fminf(1.f, a.max_norm / (total + 1e-6f));
  }
  const float step = a.lr / a.bc1;
  const float decay = 1.f - a.lr * a.wd;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    gi *= gs;
    mi = a.beta1 * mi + (1.f - a.beta1) * gi;
    vi = a.beta2 * vi + (1.f - a.beta2) * gi * gi;
    pi = pi * decay - step * (mi / (sqrtf(vi) / a.bc2_sqrt + a.eps));
  };
  // 16-byte vectors (the flat buffers are 16-byte aligned; the caller checks), scalar tail
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; reinterpret_cast<float4*>(p)[i] = p4;
    if (mirror != nullptr) {
      uint2 o;
      o.x = gwd_pack_bf16x2(p4.x, p4.y); o.y = gwd_pack_bf16x2(p4.z, p4.w);
      reinterpret_cast<uint2*>(mirror)[i] = o;
    }
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    float mi = m[i], vi = v[i], pi = p[i];
    upd(g[i], mi, vi, pi);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (mirror != nullptr) mirror[i] = __float2bfloat16(pi);
  }
}

}  // namespace

#define GWD_STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int gwd_layernorm_bwd(const void* dy, int64_t dy_rs, const void* z, int64_t z_rs, const float* gamma, const float* beta,
                                 int32_t post_act, float eps, const void* add, int64_t add_rs, void* dz, int64_t dz_rs,
                                 float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(dy && z && gamma && dz && rows > 0, "gwd_layernorm_bwd: null pointer / empty");
  if (n <= 0) n = C;
  GWD_CHECK_ARG(n <= C, "gwd_layernorm_bwd: n > C");
  GWD_CHECK_ARG(post_act == GWD_ACT_NONE || beta != nullptr, "gwd_layernorm_bwd: beta needed to differentiate through act(LN(z))");
  GWD_CHECK_ARG(C % 8 == 0 && C > 0 && C <= 512 && dy_rs % 8 == 0 && z_rs % 8 == 0 && dz_rs % 8 == 0 && add_rs % 8 == 0,
                "gwd_layernorm_bwd: C and strides must be multiples of 8, C <= 512");
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(rows, kLnWarps * 2), 4 * gwd_num_sms()));
  auto launch = [&](auto kern) {
    return gwd_launch(kern, dim3(grid), dim3(kLnWarps * 32), 0, stream, 1, static_cast<const bf16*>(dy), dy_rs,
                      static_cast<const bf16*>(z), z_rs, gamma, beta, post_act, eps, static_cast<const bf16*>(add), add_rs,
                      static_cast<bf16*>(dz), dz_rs, dgamma, dbeta, rows, C, n);
  };
  if (C <= 64) GWD_CUDA(launch(gwd_layernorm_bwd_kernel<1, 8>));
  else if (C <= 128) GWD_CUDA(launch(gwd_layernorm_bwd_kernel<1, 16>));
  else if (C <= 256) GWD_CUDA(launch(gwd_layernorm_bwd_kernel<1, 32>));
  else GWD_CUDA(launch(gwd_layernorm_bwd_kernel<2, 32>));
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_act_bwd(const void* dy, int32_t dy_f32, int64_t dy_rs, const void* y, int32_t y_f32, int64_t y_rs,
                           int32_t act, void* out, int64_t out_rs, int64_t rows, int32_t n, int32_t out_cols, float y_mul,
                           float scale, int32_t from_input, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(dy && out && rows > 0 && n > 0 && out_cols >= n && out_rs >= out_cols, "gwd_act_bwd: bad argument");
  GWD_CHECK_ARG(act == GWD_ACT_NONE || act == GWD_ACT_RELU || act == GWD_ACT_SIGMOID || act == GWD_ACT_ELU ||
                    (act == GWD_ACT_GELU && from_input),
                "gwd_act_bwd: activation %d unsupported (GELU is not invertible from its output: pass its input, from_input = 1)", act);
  GWD_CHECK_ARG(act == GWD_ACT_NONE || y != nullptr, "gwd_act_bwd: y needed");
  bf16* o = static_cast<bf16*>(out);
  if (!dy_f32 && !y_f32 && act != GWD_ACT_NONE && n == out_cols && dy_rs == n && y_rs == n && out_rs == n && n % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int64_t n8 = rows * n / 8;
    const unsigned g = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(n8, 256), 16 * gwd_num_sms()));
    const bf16* dyp = static_cast<const bf16*>(dy);
    const bf16* yp = static_cast<const bf16*>(y);
#define GWD_ACT_VEC(A, F) GWD_CUDA(gwd_launch(gwd_act_bwd_vec_kernel<A, F>, dim3(g), dim3(256), 0, stream, 1, dyp, yp, o, n8, y_mul, scale))
    const int fi = from_input ? 1 : 0;
    if (act == GWD_ACT_RELU && !fi) GWD_ACT_VEC(GWD_ACT_RELU, 0);
    else if (act == GWD_ACT_RELU) GWD_ACT_VEC(GWD_ACT_RELU, 1);
    else if (act == GWD_ACT_ELU && !fi) GWD_ACT_VEC(GWD_ACT_ELU, 0);
    else if (act == GWD_ACT_ELU) GWD_ACT_VEC(GWD_ACT_ELU, 1);
    else if (act == GWD_ACT_GELU) GWD_ACT_VEC(GWD_ACT_GELU, 1);          // (GELU from its output is rejected above)
    else if (act == GWD_ACT_SIGMOID && !fi) GWD_ACT_VEC(GWD_ACT_SIGMOID, 0);
    else GWD_ACT_VEC(GWD_ACT_SIGMOID, 1);
#undef GWD_ACT_VEC
    GWD_LAUNCHED();
    return GWD_OK;
  }
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(rows * out_cols, 256), 8 * gwd_num_sms()));
  if (dy_f32 && y_f32)
    GWD_CUDA(gwd_launch(gwd_act_bwd_kernel<float, float>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const float*>(dy), dy_rs, static_cast<const float*>(y), y_rs, act, o, out_rs, rows, n, out_cols, y_mul, scale, from_input));
  else if (dy_f32)
    GWD_CUDA(gwd_launch(gwd_act_bwd_kernel<float, bf16>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const float*>(dy), dy_rs, static_cast<const bf16*>(y), y_rs, act, o, out_rs, rows, n, out_cols, y_mul, scale, from_input));
  else if (y_f32)
    GWD_CUDA(gwd_launch(gwd_act_bwd_kernel<bf16, float>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const bf16*>(dy), dy_rs, static_cast<const float*>(y), y_rs, act, o, out_rs, rows, n, out_cols, y_mul, scale, from_input));
  else
    GWD_CUDA(gwd_launch(gwd_act_bwd_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const bf16*>(dy), dy_rs, static_cast<const bf16*>(y), y_rs, act, o, out_rs, rows, n, out_cols, y_mul, scale, from_input));
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_transpose(const void* x, int64_t x_rs, void* out, int64_t out_rs, int64_t rows, int64_t rows_pad, int32_t C,
                             float* colsum, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && rows > 0 && C > 0 && rows_pad >= rows && out_rs >= rows_pad, "gwd_transpose: bad argument");
  GWD_CHECK_ARG(C % 2 == 0 && x_rs % 2 == 0 && rows_pad % 2 == 0 && out_rs % 2 == 0 &&
                    (reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                "gwd_transpose: even sizes / 4-byte alignment needed");
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(rows_pad, 64)), static_cast<unsigned>(gwd_ceil_div(C, 64)));
  gwd_transpose_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(x), x_rs, static_cast<bf16*>(out), out_rs, rows,
                                                 rows_pad, C, colsum);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_attention_bwd(const gwd_attn_bwd_desc* d, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d && d->q && d->k && d->v && d->d_o && d->dq && d->dk && d->dv, "gwd_attention_bwd: null pointer");
  GWD_CHECK_ARG(d->hd == kHD, "gwd_attention_bwd: head dim must be 32 (got %d)", d->hd);
  const bool long_seq = d->Lq > 512 || d->Lk > 512;
  GWD_CHECK_ARG(d->items > 0 && d->heads > 0 && d->Lq > 0 && d->Lk > 0 && d->Lk <= 1280 && d->Lq <= 1280,
                "gwd_attention_bwd: 1 <= Lq, Lk <= 1280");
  GWD_CHECK_ARG(!long_seq || (d->o != nullptr && d->stats_ws != nullptr),
                "gwd_attention_bwd: sequences beyond 512 need the forward output `o` and the stats_ws scratch (items x heads x Lq x 2 floats)");
  GWD_CHECK_ARG(d->key_padding == nullptr || d->o != nullptr, "gwd_attention_bwd: key_padding needs the forward output `o` (tensor-core kernels)");
  const int64_t strides[14] = {d->q_item_stride, d->q_row_stride, d->k_item_stride, d->k_row_stride, d->v_item_stride,
                               d->v_row_stride, d->do_item_stride, d->do_row_stride, d->dq_item_stride, d->dq_row_stride,
                               d->dk_item_stride, d->dk_row_stride, d->dv_item_stride, d->dv_row_stride};
  for (int i = 0; i < 14; ++i) GWD_CHECK_ARG(strides[i] % 2 == 0, "gwd_attention_bwd: strides must be even");
  GWD_CHECK_ARG(((reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) |
                  reinterpret_cast<uintptr_t>(d->d_o)) & 3) == 0, "gwd_attention_bwd: inputs must be 4-byte aligned");
  AttnBwdParams p;
  p.q = static_cast<const bf16*>(d->q); p.k = static_cast<const bf16*>(d->k); p.v = static_cast<const bf16*>(d->v);
  p.d_o = static_cast<const bf16*>(d->d_o);
  p.dq = static_cast<bf16*>(d->dq); p.dk = static_cast<bf16*>(d->dk); p.dv = static_cast<bf16*>(d->dv);
  p.Lq = d->Lq; p.Lk = d->Lk;
  p.q_is = strides[0]; p.q_rs = strides[1]; p.k_is = strides[2]; p.k_rs = strides[3]; p.v_is = strides[4]; p.v_rs = strides[5];
  p.do_is = strides[6]; p.do_rs = strides[7]; p.dq_is = strides[8]; p.dq_rs = strides[9]; p.dk_is = strides[10];
  p.dk_rs = strides[11]; p.dv_is = strides[12]; p.dv_rs = strides[13];
  p.scale = d->scale;
  p.dq_mul = d->dq_mul != 0.f ? d->dq_mul : d->scale;
  p.dk_mul = d->dk_mul != 0.f ? d->dk_mul : d->scale;
  p.drop_seed = nullptr; p.drop_site = 0; p.drop_thr = 0; p.drop_scale = 1.f; p.heads = d->heads;
  p.key_padding = d->key_padding; p.stats = d->stats_ws;
  if (d->dropout_seed != nullptr && d->dropout_p > 0.f) {
    GWD_CHECK_ARG(d->o != nullptr, "gwd_attention_bwd: dropout needs the forward output `o` (tensor-core kernel)");
    p.drop_seed = d->dropout_seed; p.drop_site = d->dropout_site;
    p.drop_thr = static_cast<uint32_t>(static_cast<double>(d->dropout_p) * 4294967296.0);
    p.drop_scale = 1.f / (1.f - d->dropout_p);
  }
  dim3 grid(d->heads, d->items);
  static const bool mma_enabled = []() { const char* e = getenv("GWD_ATTN_BWD_MMA"); return !(e && e[0] == '0'); }();
  if (d->o != nullptr && mma_enabled) {
    GWD_CHECK_ARG(d->o_item_stride % 2 == 0 && d->o_row_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(d->o) & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(d->dq) | reinterpret_cast<uintptr_t>(d->dk) | reinterpret_cast<uintptr_t>(d->dv)) & 3) == 0,
                  "gwd_attention_bwd: o / dq / dk / dv must be 4-byte aligned with even strides");
    p.o = static_cast<const bf16*>(d->o); p.o_is = d->o_item_stride; p.o_rs = d->o_row_stride;
    p.Lq_pad = static_cast<int>(gwd_ceil_div(d->Lq, 64)) * 64;
    p.Lk_pad = static_cast<int>(gwd_ceil_div(d->Lk, 64)) * 64;
    if (long_seq) {
      const size_t sm_q = static_cast<size_t>(2 * p.Lk_pad + 256) * kMW * 4 + 128 * 4 + static_cast<size_t>(p.Lk_pad / 64) * 8;
      const size_t sm_kv = static_cast<size_t>(2 * p.Lq_pad + 128) * kMW * 4 + static_cast<size_t>(2 * p.Lq_pad) * 4;
      GWD_CUDA(cudaFuncSetAttribute(gwd_attention_bwd_mma_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_q)));
      GWD_CUDA(cudaFuncSetAttribute(gwd_attention_bwd_mma_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm_kv)));
      gwd_attention_bwd_mma_q_kernel<<<dim3(d->heads, d->items, static_cast<unsigned>(gwd_ceil_div(d->Lq, 128))), 256, sm_q, stream>>>(p);
      GWD_LAUNCHED();
      gwd_attention_bwd_mma_kv_kernel<<<dim3(d->heads, d->items, static_cast<unsigned>(gwd_ceil_div(d->Lk, 64))), 128, sm_kv, stream>>>(p);
      GWD_LAUNCHED();
      return GWD_OK;
    }
    const size_t sm = static_cast<size_t>(2 * p.Lq_pad + 2 * p.Lk_pad) * kMW * 4 + static_cast<size_t>(2 * p.Lq_pad) * 4 +
                      static_cast<size_t>(p.Lk_pad / 64) * 8;
    GWD_CUDA(cudaFuncSetAttribute(gwd_attention_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm)));
    GWD_CUDA(gwd_launch(gwd_attention_bwd_mma_kernel, grid, dim3(256), sm, stream, 1, p));
    GWD_LAUNCHED();
    return GWD_OK;
  }
  GWD_CHECK_ARG(!long_seq, "gwd_attention_bwd: sequences beyond 512 need the tensor-core path");
  const size_t smem = static_cast<size_t>(2 * d->Lq + 2 * d->Lk) * kRowW * 4 + static_cast<size_t>(2 * d->Lq) * 4;
  auto launch = [&](auto kern) -> int {
    GWD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kBwdWarps * 32, smem, stream>>>(p);
    return GWD_OK;
  };
  int rc;
  if (d->Lk <= 128) rc = launch(gwd_attention_bwd_kernel<4>);
  else if (d->Lk <= 320) rc = launch(gwd_attention_bwd_kernel<10>);
  else rc = launch(gwd_attention_bwd_kernel<16>);
  if (rc != GWD_OK) return rc;
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_sumsq(const float* g, int64_t n, double* out_accum, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(g && out_accum && n > 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0, "gwd_sumsq: bad argument");
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(n, 256 * 16), 4 * gwd_num_sms()));
  gwd_sumsq_kernel<<<grid, 256, 0, stream>>>(g, n, out_accum);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_adamw_step(float* p, const float* g, float* m, float* v, void* mirror_bf16, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int32_t step, float max_norm, float grad_scale,
                              const double* sumsq, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "gwd_adamw_step: bad argument");
  GWD_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(mirror_bf16) & 7) == 0,
                "gwd_adamw_step: buffers must be 16-byte (mirror: 8-byte) aligned");
  AdamParams a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  a.max_norm = max_norm; a.grad_scale = grad_scale;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(n, 256 * 4), 16 * gwd_num_sms()));
  gwd_adamw_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, static_cast<bf16*>(mirror_bf16), n, a, sumsq);
  GWD_LAUNCHED();
  return GWD_OK;
}

static int launch_wgrad(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int32_t N, int32_t K, float* dw,
                        int64_t dw_rs, float* db, int H, int W, int sy, int sx, cudaStream_t stream);

int gwd_linear_wgrad_tc_try(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int N, int K, float* dw,
                            int64_t dw_rs, float* db, int* db_done, cudaStream_t stream);     // gwd_wgrad_tc.cu (tcgen05 path)

namespace {
// db[n] += sum_r dY[r][n]: 8 columns per thread.  A row segment of min(N, 256) columns is held by LPR = 1..32 lanes (a power of
// two >= columns / 8), so narrow matrices (N = 64: the 1/4-scale Swin stage) put 32 / LPR rows in every warp instead of leaving
// three quarters of the lanes idle; four independent 16-byte loads per thread and iteration; the lanes that share a column meet
// by shuffles, the 8 warps in shared memory, one atomic per column and CTA.
__global__ void __launch_bounds__(256) gwd_colsum_kernel(const bf16* __restrict__ dy, int64_t dy_rs, int64_t rows, int N, int lpr,
                                                         float* __restrict__ db) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  __shared__ float red[8][32][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cv = lane & (lpr - 1), sub = lane / lpr, rpw = 32 / lpr;
  const int c = (blockIdx.x * 32 + cv) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < N) {
    const int64_t step = static_cast<int64_t>(gridDim.y) * 8 * rpw;
    int64_t r = (static_cast<int64_t>(blockIdx.y) * 8 + warp) * rpw + sub;
    for (; r + 3 * step < rows; r += 4 * step) {
      float f0[8], f1[8], f2[8], f3[8];
      ld8(dy + r * dy_rs + c, f0);
      ld8(dy + (r + step) * dy_rs + c, f1);
      ld8(dy + (r + 2 * step) * dy_rs + c, f2);
      ld8(dy + (r + 3 * step) * dy_rs + c, f3);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += (f0[i] + f1[i]) + (f2[i] + f3[i]);
    }
    for (; r < rows; r += step) {
      float f[8];
      ld8(dy + r * dy_rs + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    for (int o = lpr; o < 32; o <<= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    red[warp][lane][i] = acc[i];
  }
  __syncthreads();
  if (warp == 0 && lane < lpr && c < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][lane][i];
      atomicAdd(db + c + i, t);
    }
  }
}
}  // namespace

extern "C" int gwd_linear_wgrad(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int32_t N, int32_t K,
                                float* dw, int64_t dw_rs, float* db, void* stream_) {
  GWD_STREAM;
  // many-row Linears (window tokens of the 1/4- and 1/8-scale Swin stages): tcgen05 kernel of gwd_wgrad_tc.cu + a column-sum pass
  static const bool force_mma = [] { const char* e = getenv("GWD_WGRAD"); return e && strcmp(e, "mma") == 0; }();
  if (!force_mma && dy && x && dw && N % 8 == 0) {
    int db_done = 0;
    const int rc = gwd_linear_wgrad_tc_try(dy, dy_rs, x, x_rs, rows, N, K, dw, dw_rs, db, &db_done, stream);
    if (rc < 0) return rc;
    if (rc == 0) {
      if (db != nullptr && !db_done) {
        int lpr = 1;
        while (lpr < 32 && lpr * 8 < N) lpr <<= 1;
        const dim3 grid(static_cast<unsigned>(gwd_ceil_div(N, 256)),
                        static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(rows, 64 * (32 / lpr)), 2 * gwd_num_sms())));
        GWD_CUDA(gwd_launch(gwd_colsum_kernel, grid, dim3(256), 0, stream, 1, static_cast<const bf16*>(dy), dy_rs, rows, N, lpr, db));
        GWD_LAUNCHED();
      }
      return GWD_OK;
    }
  }
  return launch_wgrad(dy, dy_rs, x, x_rs, rows, N, K, dw, dw_rs, db, 0, 0, 0, 0, stream);
}

int gwd_conv3x3_wgrad_tc_try(const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, int B, int H, int W, int N, int C,
                             float* dw, cudaStream_t stream);     // gwd_wgrad_tc.cu (tcgen05 path)

// 3x3 convolution (stride 1, zero padding 1) weight gradient in the packed tap-major layout of gwd_conv_gemm:
// dw[dx*3+dy][n][c] += sum_{b,y,x} dy[b,y,x,n] * x[b, y+dy-1, x+dx-1, c]; one split-K pass per tap
extern "C" int gwd_conv3x3_wgrad(const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, int32_t B, int32_t H, int32_t W,
                                 int32_t N, int32_t C, float* dw, float* db, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(B > 0 && H > 0 && W > 0, "gwd_conv3x3_wgrad: empty map");
  const int64_t rows = static_cast<int64_t>(B) * H * W;
  const bool small_ok = dy && x && dw && dy_cs % 8 == 0 && x_cs % 8 == 0 && dy_cs >= N && x_cs >= C &&
                        ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(dw) & 7) == 0;
  if (small_ok) {   // narrow convolutions over many pixels: all 9 taps in one pass
#define GWD_WS(NP, CP) if (N == NP && C == CP) return launch_wgrad_small<NP, CP>(dy, dy_cs, x, x_cs, B, H, W, dw, db, stream)
    GWD_WS(16, 16); GWD_WS(16, 32); GWD_WS(16, 64); GWD_WS(32, 16); GWD_WS(32, 32); GWD_WS(32, 64); GWD_WS(64, 16); GWD_WS(64, 32); GWD_WS(64, 64);
#undef GWD_WS
  }
  // wide convolutions (the 1/4-scale pyramid): all taps on the tcgen05 kernel of gwd_wgrad_tc.cu (MN-major TMA operands)
  static const bool force_mma = [] { const char* e = getenv("GWD_WGRAD"); return e && strcmp(e, "mma") == 0; }();
  if (db == nullptr && dy && x && dw && !force_mma) {
    const int rc = gwd_conv3x3_wgrad_tc_try(dy, dy_cs, x, x_cs, B, H, W, N, C, dw, stream);
    if (rc <= 0) return rc;
  }
  for (int dx = 0; dx < 3; ++dx)
    for (int dyy = 0; dyy < 3; ++dyy) {
      const int tap = dx * 3 + dyy;
      int rc = launch_wgrad(dy, dy_cs, x, x_cs, rows, N, C, dw + static_cast<int64_t>(tap) * N * C, C, tap == 4 ? db : nullptr, H, W,
                            dyy - 1, dx - 1, stream);
      if (rc != GWD_OK) return rc;
    }
  return GWD_OK;
}

static int launch_wgrad(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int32_t N, int32_t K, float* dw,
                        int64_t dw_rs, float* db, int H, int W, int sy, int sx, cudaStream_t stream) {
  GWD_CHECK_ARG(dy && x && dw && rows > 0 && N > 0 && K > 0, "gwd_linear_wgrad: null pointer / empty");
  GWD_CHECK_ARG(N % 8 == 0 && K % 8 == 0 && dy_rs % 8 == 0 && x_rs % 8 == 0 && dw_rs % 2 == 0 && dy_rs >= N && x_rs >= K &&
                    dw_rs >= K && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dw) & 7) == 0,
                "gwd_linear_wgrad: N, K, strides must be multiples of 8 (dw: 2) and pointers 16-byte (dw: 8-byte) aligned");
  GWD_CHECK_ARG(rows < (1ll << 31), "gwd_linear_wgrad: too many rows");
  WgradParams p;
  p.dy = static_cast<const bf16*>(dy); p.x = static_cast<const bf16*>(x); p.dw = dw; p.db = db;
  p.dy_rs = dy_rs; p.x_rs = x_rs; p.dw_rs = dw_rs; p.R = static_cast<int>(rows); p.N = N; p.K = K;
  p.H = H; p.W = W; p.sy = sy; p.sx = sx;
  const int tiles = static_cast<int>(gwd_ceil_div(N, kWgT) * gwd_ceil_div(K, kWgT));
  static const int ctas_per_sm_x2 = []() { const char* e = getenv("GWD_WGRAD_CTAS_X2"); return e ? atoi(e) : 4; }();
  int64_t split = std::max<int64_t>(1, std::min<int64_t>(gwd_ceil_div(ctas_per_sm_x2 * gwd_num_sms() / 2, tiles), gwd_ceil_div(rows, 2 * kWgR)));
  p.rows_per_split = static_cast<int>(gwd_ceil_div(gwd_ceil_div(rows, split), kWgR) * kWgR);
  split = gwd_ceil_div(rows, p.rows_per_split);
  const size_t smem = static_cast<size_t>(2) * kWgStages * kWgR * kWgLd * sizeof(bf16);
  GWD_CUDA(cudaFuncSetAttribute(gwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(N, kWgT)), static_cast<unsigned>(gwd_ceil_div(K, kWgT)), static_cast<unsigned>(split));
  GWD_CUDA(gwd_launch(gwd_wgrad_kernel, grid, dim3(256), smem, stream, 1, p));
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_set_loss(const float* logits, const float* lines, const float* tgt_lines, const int64_t* tgt_labels,
                            const int32_t* match, const int32_t* stage_off, const float* class_w, const float* w_ce,
                            const float* w_line, const float* num_items, int32_t S, int32_t B, int32_t Q, int32_t C, int32_t D,
                            int32_t M, float* losses, float* dlogits, float* dlines, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(logits && lines && tgt_lines && tgt_labels && match && stage_off && class_w && w_ce && w_line && num_items &&
                    losses && dlogits && dlines, "gwd_set_loss: null pointer");
  GWD_CHECK_ARG(S > 0 && B > 0 && Q > 0 && C >= 2 && C <= kMaxClasses && D > 0 && M >= 0 && B * Q <= 160 * 1024,
                "gwd_set_loss: bad shape (2 <= classes <= 8, B*Q <= 163840)");
  SetLossParams p;
  p.logits = logits; p.lines = lines; p.tgt_lines = tgt_lines; p.tgt_labels = tgt_labels; p.match = match; p.stage_off = stage_off;
  p.class_w = class_w; p.w_ce = w_ce; p.w_line = w_line; p.num_items = num_items; p.losses = losses; p.dlogits = dlogits;
  p.dlines = dlines; p.S = S; p.BQ = B * Q; p.Q = Q; p.C = C; p.D = D; p.M = M;
  const size_t smem = static_cast<size_t>(B) * Q;
  if (smem > 48 * 1024) GWD_CUDA(cudaFuncSetAttribute(gwd_set_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  gwd_set_loss_kernel<<<S, 256, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_transpose_batch(const int64_t* table, const int32_t* tile_prefix, int32_t n_mats, int32_t total_tiles,
                                   void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(table && tile_prefix && n_mats > 0 && total_tiles > 0, "gwd_transpose_batch: bad argument");
  gwd_transpose_batch_kernel<<<total_tiles, 256, 0, stream>>>(table, tile_prefix, n_mats);
  GWD_LAUNCHED();
  return GWD_OK;
}

// ------------------------------------------------------------------------------------------------
// dense losses, backward: SilogLoss (src/models/glassrgbd.py:360-374 under the per-scale loop of
// src/engine_glassrgbd.py:65-82) and SegLoss = mean cross entropy (src/models/glassrgbd.py:376-383)
// ------------------------------------------------------------------------------------------------
namespace {

// loss = weight * 10 * sqrt(V), V = S2/n - vf (S1/n)^2 over the valid pixels (sums3 = n, S1, S2 from gwd_silog_sums):
//   dloss/dd_i = weight * 10 / (2 sqrt(V)) * (2 d_i / n - 2 vf S1 / n^2);   dd/dp = 1/p (log only) or 1 + 1/p
// sig_scale > 0: p = sig_scale * sigmoid(z) and the gradient is taken to z (times p (1 - p / sig_scale)).
// out_cols == 0: fp32 [B*h*w]; else bf16 rows of out_cols columns, gradient in column 0, zeros elsewhere.
__global__ void __launch_bounds__(256)
gwd_silog_bwd_kernel(const float* __restrict__ pred, int B, int h, int w, const float* __restrict__ gt, int H, int W, float lo,
                     float hi, int log_only, const double* __restrict__ sums, float vf, float weight, float sig_scale,
                     void* __restrict__ out, int out_cols, float* __restrict__ loss_out) {
  const double n = sums[0];
  const double mean = n > 0 ? sums[1] / n : 0.0;
  const double V = n > 0 ? sums[2] / n - static_cast<double>(vf) * mean * mean : 0.0;
  const double root = V > 0 ? sqrt(V) : 0.0;
  if (loss_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) loss_out[0] = static_cast<float>(weight * 10.0 * root);
  const float a = root > 0 ? static_cast<float>(weight * 10.0 / root / n) : 0.f;   // d_i coefficient
  const float c = static_cast<float>(static_cast<double>(vf) * mean);
  const int64_t total = static_cast<int64_t>(B) * h * w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = i % w;
    const int y = (i / w) % h;
    const int b = i / (static_cast<int64_t>(w) * h);
    const int sy = min(static_cast<int>((static_cast<int64_t>(y) * H) / h), H - 1);
    const int sx = min(static_cast<int>((static_cast<int64_t>(x) * W) / w), W - 1);
    const float g = gt[(static_cast<int64_t>(b) * H + sy) * W + sx];
    float v = 0.f;
    if (g >= lo && g < hi) {
      const float q = pred[i];
      const float d = log_only ? (logf(q) - logf(g)) : ((q + logf(q)) - (g + logf(g)));
      v = a * (d - c) * (log_only ? 1.f / q : 1.f + 1.f / q);
      if (sig_scale > 0.f) v *= q * (1.f - q / sig_scale);
    }
    if (out_cols == 0) {
      static_cast<float*>(out)[i] = v;
    } else {
      bf16* row = static_cast<bf16*>(out) + i * out_cols;
      float r[8] = {v, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st8(row, r);
      r[0] = 0.f;
      for (int k = 8; k < out_cols; k += 8) st8(row + k, r);
    }
  }
}

constexpr int kCeMaxClasses = 8;

// sums[0] += valid pixels, sums[1] += sum of -log softmax(logits)[gt]
__global__ void __launch_bounds__(256)
gwd_seg_ce_sums_kernel(const float* __restrict__ logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride,
                       const int64_t* __restrict__ gt, int64_t HW, int64_t total, int C, int ignore, double* __restrict__ sums) {
  double a0 = 0, a1 = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t g = gt[i];
    if (g == ignore || g < 0 || g >= C) continue;
    const float* px = logits + (i / HW) * image_stride + (i % HW) * pixel_stride;
    float m = px[0];
    for (int k = 1; k < C; ++k) m = fmaxf(m, px[k * class_stride]);
    float s = 0.f;
    for (int k = 0; k < C; ++k) s += expf(px[k * class_stride] - m);
    a0 += 1.0;
    a1 += static_cast<double>(m + logf(s) - px[g * class_stride]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0 && a0 > 0) {
    atomicAdd(&sums[0], a0);
    atomicAdd(&sums[1], a1);
  }
}

// dlogits[i, k] = weight / n * (softmax_k - [k == gt]) as bf16 rows of out_cols columns (zeros beyond C); loss = weight * S / n
__global__ void __launch_bounds__(256)
gwd_seg_ce_bwd_kernel(const float* __restrict__ logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride,
                      const int64_t* __restrict__ gt, int64_t HW, int64_t total, int C, int ignore,
                      const double* __restrict__ sums, float weight, bf16* __restrict__ out, int out_cols,
                      float* __restrict__ loss_out) {
  const double n = sums[0];
  if (loss_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) loss_out[0] = n > 0 ? static_cast<float>(weight * sums[1] / n) : 0.f;
  const float a = n > 0 ? static_cast<float>(weight / n) : 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float r[kCeMaxClasses];
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k) r[k] = 0.f;
    const int64_t g = gt[i];
    if (!(g == ignore || g < 0 || g >= C)) {
      const float* px = logits + (i / HW) * image_stride + (i % HW) * pixel_stride;
      float m = px[0];
      for (int k = 1; k < C; ++k) m = fmaxf(m, px[k * class_stride]);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kCeMaxClasses; ++k)
        if (k < C) { r[k] = expf(px[k * class_stride] - m); s += r[k]; }
      const float inv = a / s;
#pragma unroll
      for (int k = 0; k < kCeMaxClasses; ++k)
        if (k < C) r[k] = r[k] * inv - (k == g ? a : 0.f);
    }
    bf16* row = out + i * out_cols;
    st8(row, r);
#pragma unroll
    for (int k = 0; k < kCeMaxClasses; ++k) r[k] = 0.f;
    for (int k = 8; k < out_cols; k += 8) st8(row + k, r);
  }
}

}  // namespace

extern "C" int gwd_silog_bwd(const float* pred, int32_t B, int32_t h, int32_t w, const float* gt, int32_t H, int32_t W, float lo,
                             float hi, int32_t log_only, const double* sums3, float variance_focus, float weight,
                             float sig_scale, void* out, int32_t out_cols, float* loss_out, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(pred && gt && sums3 && out && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "gwd_silog_bwd: bad argument");
  GWD_CHECK_ARG(out_cols >= 0 && out_cols % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "gwd_silog_bwd: out_cols must be 0 (fp32) or a multiple of 8 (bf16 rows), out 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * h * w;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(total, 256), 16 * gwd_num_sms()));
  gwd_silog_bwd_kernel<<<grid, 256, 0, stream>>>(pred, B, h, w, gt, H, W, lo, hi, log_only, sums3, variance_focus, weight,
                                                 sig_scale, out, out_cols, loss_out);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_seg_ce(const float* logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride,
                          const int64_t* gt, int32_t B, int64_t HW, int32_t C, int32_t ignore_index, float weight, double* sums2,
                          void* dlogits, int32_t out_cols, float* loss_out, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(logits && gt && sums2 && B > 0 && HW > 0 && C >= 2 && C <= kCeMaxClasses, "gwd_seg_ce: bad argument (2 <= classes <= 8)");
  GWD_CHECK_ARG(dlogits == nullptr || (out_cols >= 8 && out_cols % 8 == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0),
                "gwd_seg_ce: dlogits rows must be a multiple of 8 bf16 columns, 16-byte aligned");
  const int64_t total = static_cast<int64_t>(B) * HW;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(total, 256), 16 * gwd_num_sms()));
  GWD_CUDA(cudaMemsetAsync(sums2, 0, sizeof(double) * 2, stream));
  gwd_seg_ce_sums_kernel<<<grid, 256, 0, stream>>>(logits, pixel_stride, class_stride, image_stride, gt, HW, total, C, ignore_index, sums2);
  GWD_LAUNCHED();
  if (dlogits != nullptr) {
    gwd_seg_ce_bwd_kernel<<<grid, 256, 0, stream>>>(logits, pixel_stride, class_stride, image_stride, gt, HW, total, C, ignore_index,
                                                    sums2, weight, static_cast<bf16*>(dlogits), out_cols, loss_out);
    GWD_LAUNCHED();
  }
  return GWD_OK;
}


// ------------------------------------------------------------------------------------------------
// PyramidLayer branches, backward (src/models/points/points_sample.py:106-125): bilinear (align_corners=True) up-sampling
// and AvgPool2d(k, k)
// ------------------------------------------------------------------------------------------------
namespace {

// dx[b, y, x, :] = sum over the high-resolution pixels (Y, X) whose bilinear footprint (computed exactly as the forward,
// gwd_bilinear_up) touches (y, x).  A gather: block = 64 channel vectors x 4 row slices, the slices reduced in shared memory.
__global__ void __launch_bounds__(256)
gwd_bilinear_up_bwd_kernel(const bf16* __restrict__ dy, int64_t dy_rs, int H, int W, bf16* __restrict__ dx, int64_t dx_rs, int h,
                           int w, int C) {
  __shared__ float red[4][64][8];
  const int cv = C / 8;
  const int t = blockIdx.x * 64 + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  const bool live = t < w * cv;
  const int x = live ? t / cv : 0, c = live ? (t - x * cv) * 8 : 0;
  const float ry = (H > 1) ? static_cast<float>(h - 1) / (H - 1) : 0.f;
  const float rx = (W > 1) ? static_cast<float>(w - 1) / (W - 1) : 0.f;
  int Ylo = 0, Yhi = H - 1, Xlo = 0, Xhi = W - 1;
  if (ry > 0.f) { Ylo = max(0, static_cast<int>(floorf((y - 1) / ry)) - 1); Yhi = min(H - 1, static_cast<int>(ceilf((y + 1) / ry)) + 1); }
  if (rx > 0.f) { Xlo = max(0, static_cast<int>(floorf((x - 1) / rx)) - 1); Xhi = min(W - 1, static_cast<int>(ceilf((x + 1) / rx)) + 1); }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
    for (int Y = Ylo + threadIdx.y; Y <= Yhi; Y += 4) {
      const float fy = ry * Y;
      const int y0 = static_cast<int>(fy), y1 = min(y0 + 1, h - 1);
      const float ly = fy - y0;
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      const bf16* row = dy + (static_cast<int64_t>(b) * H + Y) * W * dy_rs + c;
      for (int X = Xlo; X <= Xhi; ++X) {
        const float fx = rx * X;
        const int x0 = static_cast<int>(fx), x1 = min(x0 + 1, w - 1);
        const float lx = fx - x0;
        const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
        if (wx == 0.f) continue;
        float f[8];
        ld8(row + static_cast<int64_t>(X) * dy_rs, f);
        const float wgt = wy * wx;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && live) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = red[0][threadIdx.x][i] + red[1][threadIdx.x][i] + red[2][threadIdx.x][i] + red[3][threadIdx.x][i];
    st8(dx + ((static_cast<int64_t>(b) * h + y) * w + x) * dx_rs + c, acc);
  }
}

// the same gather for LARGE footprints (the pool-8 / pool-16 branches: a low-resolution pixel collects ~18 x 18 / ~34 x 34
// high-resolution ones, and there are only h x w x B = 120 ... 2 400 of them): one CTA per low-resolution pixel, its 256 threads
// are (footprint slots) x (channel vectors), the slots meet in shared memory.  The row-slice kernel above ran these on 24-48 CTAs
// with ~300 dependent loads per thread (0.19 ms for a 3 x 5 map).
__global__ void __launch_bounds__(256)
gwd_bilinear_up_bwd_wide_kernel(const bf16* __restrict__ dy, int64_t dy_rs, int H, int W, bf16* __restrict__ dx, int64_t dx_rs, int h,
                                int w, int C) {
  __shared__ float red[256][8];
  const int cv = C / 8;
  const int slots = 256 / cv;
  const int x = blockIdx.x, y = blockIdx.y, b = blockIdx.z;
  const int slot = threadIdx.x / cv, vec = threadIdx.x - slot * cv;
  const float ry = (H > 1) ? static_cast<float>(h - 1) / (H - 1) : 0.f;
  const float rx = (W > 1) ? static_cast<float>(w - 1) / (W - 1) : 0.f;
  int Ylo = 0, Yhi = H - 1, Xlo = 0, Xhi = W - 1;
  if (ry > 0.f) { Ylo = max(0, static_cast<int>(floorf((y - 1) / ry)) - 1); Yhi = min(H - 1, static_cast<int>(ceilf((y + 1) / ry)) + 1); }
  if (rx > 0.f) { Xlo = max(0, static_cast<int>(floorf((x - 1) / rx)) - 1); Xhi = min(W - 1, static_cast<int>(ceilf((x + 1) / rx)) + 1); }
  const int wf = Xhi - Xlo + 1, nf = (Yhi - Ylo + 1) * wf;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (slot < slots) {
    for (int p = slot; p < nf; p += slots) {
      const int Y = Ylo + p / wf, X = Xlo + p % wf;
      const float fy = ry * Y, fx = rx * X;                      // the forward's footprint arithmetic (gwd_bilinear_ac_kernel)
      const int y0 = static_cast<int>(fy), y1 = min(y0 + 1, h - 1);
      const int x0 = static_cast<int>(fx), x1 = min(x0 + 1, w - 1);
      const float ly = fy - y0, lx = fx - x0;
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
      const float wgt = wy * wx;
      if (wgt == 0.f) continue;
      float f[8];
      ld8(dy + ((static_cast<int64_t>(b) * H + Y) * W + X) * dy_rs + vec * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.x < cv) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int sl = 0; sl < slots; ++sl)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += red[sl * cv + threadIdx.x][i];
    st8(dx + ((static_cast<int64_t>(b) * h + y) * w + x) * dx_rs + threadIdx.x * 8, acc);
  }
}

// out[b, Y, X, :] = add[b, Y, X, :] + (Y / k < H / k and X / k < W / k ? d[b, Y / k, X / k, :] * scale / k^2 : 0)   (floor-mode pool)
__global__ void __launch_bounds__(256)
gwd_avgpool_bwd_kernel(const bf16* __restrict__ d, int64_t d_rs, int k, float scale, const bf16* __restrict__ add, int64_t add_rs,
                       bf16* __restrict__ out, int64_t out_rs, int H, int W, int C) {
  const int cv = C / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * cv) return;
  const int X = t / cv, c = (t - X * cv) * 8;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int oh = H / k, ow = W / k;
  const int64_t pix = (static_cast<int64_t>(b) * H + Y) * W + X;
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (Y / k < oh && X / k < ow) {
    ld8(d + ((static_cast<int64_t>(b) * oh + Y / k) * ow + X / k) * d_rs + c, f);
    const float m = scale / static_cast<float>(k * k);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] *= m;
  }
  if (add != nullptr) {
    float u[8];
    ld8(add + pix * add_rs + c, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += u[i];
  }
  st8(out + pix * out_rs + c, f);
}

}  // namespace

extern "C" int gwd_bilinear_up_bwd(const void* dy, int64_t dy_rs, int32_t B, int32_t H, int32_t W, void* dx, int64_t dx_rs,
                                   int32_t h, int32_t w, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(dy && dx && C > 0 && C % 8 == 0 && dy_rs % 8 == 0 && dx_rs % 8 == 0 && dy_rs >= C && dx_rs >= C,
                "gwd_bilinear_up_bwd: bad argument");
  GWD_CHECK_ARG(B > 0 && H > 0 && W > 0 && h > 0 && w > 0 && h <= 65535 && B <= 65535, "gwd_bilinear_up_bwd: bad extents");
  const int64_t footprint = (2 * static_cast<int64_t>(H) / h + 2) * (2 * static_cast<int64_t>(W) / w + 2);
  if (footprint >= 200 && C / 8 <= 256 && w <= 65535) {
    gwd_bilinear_up_bwd_wide_kernel<<<dim3(w, h, B), 256, 0, stream>>>(static_cast<const bf16*>(dy), dy_rs, H, W, static_cast<bf16*>(dx),
                                                                         dx_rs, h, w, C);
    GWD_LAUNCHED();
    return GWD_OK;
  }
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(w) * (C / 8), 64)), h, B);
  gwd_bilinear_up_bwd_kernel<<<grid, dim3(64, 4), 0, stream>>>(static_cast<const bf16*>(dy), dy_rs, H, W, static_cast<bf16*>(dx), dx_rs, h,
                                                              w, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_avgpool_bwd(const void* d, int64_t d_rs, int32_t k, float scale, const void* add, int64_t add_rs, void* out,
                               int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d && out && k > 0 && H >= k && W >= k && C > 0 && C % 8 == 0 && d_rs % 8 == 0 && add_rs % 8 == 0 && out_rs % 8 == 0,
                "gwd_avgpool_bwd: bad argument");
  GWD_CHECK_ARG(B > 0 && H <= 65535 && B <= 65535, "gwd_avgpool_bwd: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W) * (C / 8), 256)), H, B);
  gwd_avgpool_bwd_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(d), d_rs, k, scale, static_cast<const bf16*>(add), add_rs,
                                                   static_cast<bf16*>(out), out_rs, H, W, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

// ------------------------------------------------------------------------------------------------
// PointBasedPred, backward (src/models/points/points_sample.py:257-280): soft-max mixture of anchor depths and the
// bilinear point samples (F.grid_sample, align_corners=False, zero padding) of the reference features / previous depth
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxPoints = 128;

__device__ __forceinline__ float pt_unnorm(float g, int size) { return ((g + 1.f) * size - 1.f) * 0.5f; }

// pred[p] = sum_k a_k A[b,k] with a = softmax_k(logits[p, :K]):
//   dlogits[p,k] = a_k (A[b,k] - pred[p]) dpred[p]   (bf16 rows of Kp columns, exact zeros beyond K)
//   danchor[b,k] += sum_p a_k dpred[p]               (warp shuffle -> shared memory -> one atomic per (CTA, k))
// grid = (pixel blocks, B): a CTA never straddles two images.
__global__ void __launch_bounds__(256)
gwd_anchor_mix_bwd_kernel(const bf16* __restrict__ logits, int64_t l_rs, const float* __restrict__ anchor,
                          const float* __restrict__ dpred, int64_t HW, int K, int Kp, bf16* __restrict__ dlogits, int64_t dl_rs,
                          float* __restrict__ danchor) {
  __shared__ float s_an[kMaxPoints], s_da[kMaxPoints];
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < kMaxPoints; k += blockDim.x) {
    s_an[k] = k < K ? anchor[static_cast<int64_t>(b) * K + k] : 0.f;
    s_da[k] = 0.f;
  }
  __syncthreads();
  const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool live = p < HW;
  const int64_t pix = static_cast<int64_t>(b) * HW + (live ? p : 0);
  const bf16* row = logits + pix * l_rs;
  const float dp = live ? dpred[pix] : 0.f;
  float mx = -INFINITY;
  for (int k = 0; k < K; k += 8) {
    float f[8];
    ld8(row + k, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (k + i < K) mx = fmaxf(mx, f[i]);
  }
  float sum = 0.f, acc = 0.f;
  for (int k = 0; k < K; k += 8) {
    float f[8];
    ld8(row + k, f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (k + i < K) {
        const float e = __expf(f[i] - mx);
        sum += e;
        acc = fmaf(e, s_an[k + i], acc);
      }
  }
  const float inv = 1.f / sum, pred = acc * inv;
  const int lane = threadIdx.x & 31;
  for (int k = 0; k < Kp; k += 8) {
    float f[8], o[8];
    ld8(row + k, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool in = k + i < K;
      const float a = in ? __expf(f[i] - mx) * inv : 0.f;
      o[i] = a * (s_an[k + i] - pred) * dp;
      const float v = gwd_warp_sum(a * dp);
      if (in && lane == 0) atomicAdd(&s_da[k + i], v);
    }
    if (live) st8(dlogits + pix * dl_rs + k, o);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(danchor + static_cast<int64_t>(b) * K + k, s_da[k]);
}

// Backward of the K-point bilinear sample of a C-channel map as a GATHER over the map's pixels (deterministic, no atomics,
// writes the exact zeros of untouched pixels): dx[b,Y,X,c] = sum_k w_k(Y,X) d[b,k,c], w_k = the forward's corner weight of
// point k at pixel (Y,X).  Point footprints (x0, y0, lx, ly) of the image sit in shared memory.
__global__ void __launch_bounds__(256)
gwd_sample_bilinear_bwd_kernel(const float* __restrict__ d, const float* __restrict__ coords, int K, bf16* __restrict__ dx,
                               int64_t dx_rs, int H, int W, int C) {
  __shared__ int s_x0[kMaxPoints], s_y0[kMaxPoints];
  __shared__ float s_lx[kMaxPoints], s_ly[kMaxPoints];
  const int Y = blockIdx.y, b = blockIdx.z;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float fx = pt_unnorm(coords[(static_cast<int64_t>(b) * K + k) * 2], W);
    const float fy = pt_unnorm(coords[(static_cast<int64_t>(b) * K + k) * 2 + 1], H);
    const int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
    s_x0[k] = x0; s_y0[k] = y0; s_lx[k] = fx - x0; s_ly[k] = fy - y0;
  }
  __syncthreads();
  const int cv = C / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * cv) return;
  const int X = t / cv, c = (t - X * cv) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < K; ++k) {
    const int ddx = X - s_x0[k], ddy = Y - s_y0[k];
    if ((ddx | ddy) & ~1) continue;                       // both offsets must be 0 or 1
    const float wgt = (ddx ? s_lx[k] : 1.f - s_lx[k]) * (ddy ? s_ly[k] : 1.f - s_ly[k]);
    const float4* src = reinterpret_cast<const float4*>(d + (static_cast<int64_t>(b) * K + k) * C + c);
    const float4 u = src[0], v = src[1];
    acc[0] = fmaf(wgt, u.x, acc[0]); acc[1] = fmaf(wgt, u.y, acc[1]); acc[2] = fmaf(wgt, u.z, acc[2]); acc[3] = fmaf(wgt, u.w, acc[3]);
    acc[4] = fmaf(wgt, v.x, acc[4]); acc[5] = fmaf(wgt, v.y, acc[5]); acc[6] = fmaf(wgt, v.z, acc[6]); acc[7] = fmaf(wgt, v.w, acc[7]);
  }
  st8(dx + ((static_cast<int64_t>(b) * H + Y) * W + X) * dx_rs + c, acc);
}

// the same for the single-channel fp32 map of anchor depths: out[b,Y,X] = add[b,Y,X] + sum_k w_k(Y,X) d[b,k]
__global__ void __launch_bounds__(256)
gwd_sample_scalar_bwd_kernel(const float* __restrict__ d, const float* __restrict__ coords, int K, const float* __restrict__ add,
                             float* __restrict__ out, int H, int W) {
  __shared__ int s_x0[kMaxPoints], s_y0[kMaxPoints];
  __shared__ float s_lx[kMaxPoints], s_ly[kMaxPoints], s_d[kMaxPoints];
  const int Y = blockIdx.y, b = blockIdx.z;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float fx = pt_unnorm(coords[(static_cast<int64_t>(b) * K + k) * 2], W);
    const float fy = pt_unnorm(coords[(static_cast<int64_t>(b) * K + k) * 2 + 1], H);
    const int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
    s_x0[k] = x0; s_y0[k] = y0; s_lx[k] = fx - x0; s_ly[k] = fy - y0;
    s_d[k] = d[static_cast<int64_t>(b) * K + k];
  }
  __syncthreads();
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  if (X >= W) return;
  const int64_t pix = (static_cast<int64_t>(b) * H + Y) * W + X;
  float acc = add != nullptr ? add[pix] : 0.f;
  for (int k = 0; k < K; ++k) {
    const int ddx = X - s_x0[k], ddy = Y - s_y0[k];
    if ((ddx | ddy) & ~1) continue;
    acc = fmaf((ddx ? s_lx[k] : 1.f - s_lx[k]) * (ddy ? s_ly[k] : 1.f - s_ly[k]), s_d[k], acc);
  }
  out[pix] = acc;
}

}  // namespace

extern "C" int gwd_anchor_mix_bwd(const void* logits, int64_t l_rs, const float* anchor, const float* dpred, int32_t B, int64_t HW,
                                  int32_t K, int32_t Kp, void* dlogits, int64_t dl_rs, float* danchor, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(logits && anchor && dpred && dlogits && danchor, "gwd_anchor_mix_bwd: null pointer");
  GWD_CHECK_ARG(K > 0 && K <= Kp && Kp <= kMaxPoints && Kp % 8 == 0 && l_rs % 8 == 0 && dl_rs % 8 == 0 && l_rs >= Kp && dl_rs >= Kp,
                "gwd_anchor_mix_bwd: bad K / Kp / row strides (K=%d Kp=%d)", K, Kp);
  GWD_CHECK_ARG(B > 0 && B <= 65535 && HW > 0, "gwd_anchor_mix_bwd: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(HW, 256)), B);
  gwd_anchor_mix_bwd_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(logits), l_rs, anchor, dpred, HW, K, Kp,
                                                      static_cast<bf16*>(dlogits), dl_rs, danchor);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_sample_bilinear_bwd(const float* d, const float* coords, int32_t K, void* dx, int64_t dx_rs, int32_t B,
                                       int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d && coords && dx && K > 0 && K <= kMaxPoints && C > 0 && C % 8 == 0 && dx_rs % 8 == 0 && dx_rs >= C,
                "gwd_sample_bilinear_bwd: bad argument");
  GWD_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0, "gwd_sample_bilinear_bwd: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W) * (C / 8), 256)), H, B);
  gwd_sample_bilinear_bwd_kernel<<<grid, 256, 0, stream>>>(d, coords, K, static_cast<bf16*>(dx), dx_rs, H, W, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_sample_scalar_bwd(const float* d, const float* coords, int32_t K, const float* add, float* out, int32_t B,
                                     int32_t H, int32_t W, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d && coords && out && K > 0 && K <= kMaxPoints, "gwd_sample_scalar_bwd: bad argument");
  GWD_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && H <= 65535 && W > 0, "gwd_sample_scalar_bwd: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(W, 256)), H, B);
  gwd_sample_scalar_bwd_kernel<<<grid, 256, 0, stream>>>(d, coords, K, add, out, H, W);
  GWD_LAUNCHED();
  return GWD_OK;
}
