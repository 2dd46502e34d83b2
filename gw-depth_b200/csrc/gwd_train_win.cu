// gwd_train_win.cu -- backward of the two attention cores of the class-window Swin blocks
// (WindowClassAttention, src/models/multiscale_transformerr.py:455-580, under torch.autograd):
//   * gwd_window_attention_bwd : biased (shifted-)window multi-head self-attention over N = ws*ws <= 64 tokens, head dim
//                                <= 32: dq | dk | dv into the fused projection-gradient buffer and the gradient of the
//                                relative-position bias [heads, N, N]
//   * gwd_token_attention_bwd  : class-token CHANNEL attention (soft-max over the key channels of a head, contraction over
//                                the window's tokens): gradients of both token queries and of global_k | global_v
// First version on CUDA cores (fp32 in shared memory): a (window, head) problem is 49 x 49 x <= 32 -- 1/40 of the smallest
// tcgen05 tile -- and the whole pass is ~5 GFMA at the finest scale; CTAs are PERSISTENT over the windows of one head so
// that the bias gradient is accumulated in registers and reaches HBM as one atomic per entry and CTA.
#include <stdlib.h>
#include <algorithm>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kMaxN = 64, kMaxHd = 32;

struct WinBwdParams {
  const bf16* qkv; const bf16* d_o; bf16* dqkv;
  const float* bias; const float* mask; float* dbias;
  int items, heads, N, hd, C, mask_windows;
  int64_t qkv_rs, do_rs, dqkv_rs;
  float scale;
};

// grid = (heads, ctas_per_head), 128 threads
__global__ void __launch_bounds__(128) gwd_window_attention_bwd_kernel(const WinBwdParams p) {
  extern __shared__ float sm[];
  const int N = p.N, hd = p.hd, ld = hd + 1, lp = N + 1;
  float* sq = sm;                    // [N][ld]
  float* sk = sq + N * ld;
  float* sv = sk + N * ld;
  float* sdo = sv + N * ld;
  float* sP = sdo + N * ld;          // [N][lp]  probabilities
  float* sdS = sP + N * lp;          // [N][lp]  d scores
  float* sD = sdS + N * lp;          // [N]      sum_j P dP
  const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NN = N * N;
  float dbacc[(kMaxN * kMaxN + 127) / 128];
#pragma unroll
  for (int e = 0; e < (kMaxN * kMaxN + 127) / 128; ++e) dbacc[e] = 0.f;
  const float* bias_h = p.bias ? p.bias + static_cast<int64_t>(h) * NN : nullptr;

  for (int item = blockIdx.y; item < p.items; item += gridDim.y) {
    const int64_t row0 = static_cast<int64_t>(item) * N;
    for (int idx = tid; idx < N * hd; idx += 128) {
      const int n = idx / hd, d = idx - n * hd;
      const bf16* r = p.qkv + (row0 + n) * p.qkv_rs + h * hd + d;
      sq[n * ld + d] = __bfloat162float(r[0]);
      sk[n * ld + d] = __bfloat162float(r[p.C]);
      sv[n * ld + d] = __bfloat162float(r[2 * p.C]);
      sdo[n * ld + d] = __bfloat162float(p.d_o[(row0 + n) * p.do_rs + h * hd + d]);
    }
    __syncthreads();
    const float* mask_w = p.mask ? p.mask + static_cast<int64_t>(item % p.mask_windows) * NN : nullptr;
    // scores and dP = dO V^T
    for (int e = tid; e < NN; e += 128) {
      const int i = e / N, j = e - i * N;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < hd; ++d) {
        s = fmaf(sq[i * ld + d], sk[j * ld + d], s);
        dp = fmaf(sdo[i * ld + d], sv[j * ld + d], dp);
      }
      s *= p.scale;
      if (bias_h) s += bias_h[e];
      if (mask_w) s += mask_w[e];
      sP[i * lp + j] = s;
      sdS[i * lp + j] = dp;
    }
    __syncthreads();
    // soft-max rows, D_i = sum_j P_ij dP_ij, dS = P (dP - D)
    for (int i = warp; i < N; i += 4) {
      float mx = -INFINITY;
      for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sP[i * lp + j]);
      mx = gwd_warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float e = __expf(sP[i * lp + j] - mx);
        sP[i * lp + j] = e;
        sum += e;
      }
      sum = gwd_warp_sum(sum);
      const float inv = 1.f / sum;
      float dsum = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float pr = sP[i * lp + j] * inv;
        sP[i * lp + j] = pr;
        dsum = fmaf(pr, sdS[i * lp + j], dsum);
      }
      dsum = gwd_warp_sum(dsum);
      for (int j = lane; j < N; j += 32) sdS[i * lp + j] = sP[i * lp + j] * (sdS[i * lp + j] - dsum);
      if (lane == 0) sD[i] = dsum;
    }
    __syncthreads();
    if (p.dbias) {
      int e = tid;
#pragma unroll
      for (int r = 0; r < (kMaxN * kMaxN + 127) / 128; ++r, e += 128)
        if (e < NN) dbacc[r] += sdS[(e / N) * lp + (e % N)];
    }
    // dq = scale dS K, dk = scale dS^T Q, dv = P^T dO
    for (int idx = tid; idx < N * hd; idx += 128) {
      const int n = idx / hd, d = idx - n * hd;
      float aq = 0.f, ak = 0.f, av = 0.f;
      for (int j = 0; j < N; ++j) {
        aq = fmaf(sdS[n * lp + j], sk[j * ld + d], aq);
        ak = fmaf(sdS[j * lp + n], sq[j * ld + d], ak);
        av = fmaf(sP[j * lp + n], sdo[j * ld + d], av);
      }
      bf16* o = p.dqkv + (row0 + n) * p.dqkv_rs + h * hd + d;
      o[0] = __float2bfloat16(aq * p.scale);
      o[p.C] = __float2bfloat16(ak * p.scale);
      o[2 * p.C] = __float2bfloat16(av);
    }
    __syncthreads();
  }
  if (p.dbias) {
    float* db = p.dbias + static_cast<int64_t>(h) * NN;
    int e = tid;
#pragma unroll
    for (int r = 0; r < (kMaxN * kMaxN + 127) / 128; ++r, e += 128)
      if (e < NN) atomicAdd(db + e, dbacc[r]);
  }
}

struct TokBwdParams {
  const bf16* dq; const bf16* sq; const bf16* tk; const bf16* tv;
  const bf16* d_dout; const bf16* d_sout;
  bf16* g_dq; bf16* g_sq; bf16* g_tk; bf16* g_tv;
  int items, N, heads, td, tc;
  int64_t q_rs, k_rs, v_rs, o_rs, gq_rs, gk_rs, gv_rs;
  float scale;
};

constexpr int kTokR = 16, kTokC = 32;     // (depth | seg) query channels per head <= 16, key channels per head <= 32

// one CTA (128 threads) per (window, head): rows r = [depth td | seg td], a[r][c] = softmax_c(scale sum_n q[n][r] k[n][c]),
// out[n][r] = sum_c a[r][c] v[n][c]
__global__ void __launch_bounds__(128) gwd_token_attention_bwd_kernel(const TokBwdParams p) {
  __shared__ float q[kMaxN][kTokR + 1], go[kMaxN][kTokR + 1], k[kMaxN][kTokC + 1], v[kMaxN][kTokC + 1];
  __shared__ float a[kTokR][kTokC + 1], dz[kTokR][kTokC + 1];
  const int N = p.N, td = p.td, tc = p.tc, R = 2 * td;
  const int tid = threadIdx.x;
  for (int wh = blockIdx.x; wh < p.items * p.heads; wh += gridDim.x) {
    const int item = wh / p.heads, h = wh - item * p.heads;
    const int64_t row0 = static_cast<int64_t>(item) * N;
    for (int idx = tid; idx < N * R; idx += 128) {
      const int n = idx / R, r = idx - n * R;
      const bool seg = r >= td;
      const int c = h * td + (seg ? r - td : r);
      q[n][r] = __bfloat162float((seg ? p.sq : p.dq)[(row0 + n) * p.q_rs + c]);
      go[n][r] = __bfloat162float((seg ? p.d_sout : p.d_dout)[(row0 + n) * p.o_rs + c]);
    }
    for (int idx = tid; idx < N * tc; idx += 128) {
      const int n = idx / tc, c = idx - n * tc;
      k[n][c] = __bfloat162float(p.tk[(row0 + n) * p.k_rs + h * tc + c]);
      v[n][c] = __bfloat162float(p.tv[(row0 + n) * p.v_rs + h * tc + c]);
    }
    __syncthreads();
    // scores and d a = sum_n go[n][r] v[n][c]
    for (int e = tid; e < R * tc; e += 128) {
      const int r = e / tc, c = e - r * tc;
      float s = 0.f, da = 0.f;
      for (int n = 0; n < N; ++n) {
        s = fmaf(q[n][r], k[n][c], s);
        da = fmaf(go[n][r], v[n][c], da);
      }
      a[r][c] = s * p.scale;
      dz[r][c] = da;
    }
    __syncthreads();
    if (tid < R) {      // soft-max over the key channels of row tid and its backward (rows are <= 64 wide)
      const int r = tid;
      float mx = -INFINITY;
      for (int c = 0; c < tc; ++c) mx = fmaxf(mx, a[r][c]);
      float sum = 0.f;
      for (int c = 0; c < tc; ++c) { const float e = __expf(a[r][c] - mx); a[r][c] = e; sum += e; }
      const float inv = 1.f / sum;
      float dsum = 0.f;
      for (int c = 0; c < tc; ++c) { a[r][c] *= inv; dsum = fmaf(a[r][c], dz[r][c], dsum); }
      for (int c = 0; c < tc; ++c) dz[r][c] = a[r][c] * (dz[r][c] - dsum) * p.scale;
    }
    __syncthreads();
    // d q[n][r] = sum_c dz[r][c] k[n][c]
    for (int idx = tid; idx < N * R; idx += 128) {
      const int n = idx / R, r = idx - n * R;
      float acc = 0.f;
      for (int c = 0; c < tc; ++c) acc = fmaf(dz[r][c], k[n][c], acc);
      const bool seg = r >= td;
      (seg ? p.g_sq : p.g_dq)[(row0 + n) * p.gq_rs + h * td + (seg ? r - td : r)] = __float2bfloat16(acc);
    }
    // d k[n][c] = sum_r dz[r][c] q[n][r];  d v[n][c] = sum_r a[r][c] go[n][r]
    for (int idx = tid; idx < N * tc; idx += 128) {
      const int n = idx / tc, c = idx - n * tc;
      float gk = 0.f, gv = 0.f;
      for (int r = 0; r < R; ++r) {
        gk = fmaf(dz[r][c], q[n][r], gk);
        gv = fmaf(a[r][c], go[n][r], gv);
      }
      p.g_tk[(row0 + n) * p.gk_rs + h * tc + c] = __float2bfloat16(gk);
      p.g_tv[(row0 + n) * p.gv_rs + h * tc + c] = __float2bfloat16(gv);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Register-row version for 7 x 7 windows (N = 49) and head dims 4 / 8 / 16 -- the shapes of the three class-window stages.
// The first kernel above spends its time in shared-memory loads (2 LDS per FMA in every product).  Here a thread OWNS a score
// row: q_i and dO_i live in registers, k_j / v_j arrive as 16-byte shared-memory BROADCASTS (one LDS serves the warp), the
// soft-max statistics and D_i = sum_j P_ij dP_ij come from one online pass, a second pass re-forms the scores, emits P and dS
// into a conflict-free [49][49] transpose buffer (odd row stride) for the column products (dk, dv), accumulates dq in
// registers and the bias gradient in a shared-memory tile that only its owner thread touches.  A CTA = two 64-thread slots =
// two heads of the same window (they share the shift mask); CTAs are persistent over the windows of their head pair, with
// the two bias tables resident in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kWN = 49;

template <int HD>
__device__ __forceinline__ void load_row_bf16(const bf16* src, float* dst) {
  if constexpr (HD == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(src);
    const float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y);
    *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
  } else {
#pragma unroll
    for (int c = 0; c < HD; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + c);
      const float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y), e = gwd_unpack_bf16x2(u.z), f = gwd_unpack_bf16x2(u.w);
      *reinterpret_cast<float4*>(dst + c) = make_float4(a.x, a.y, b.x, b.y);
      *reinterpret_cast<float4*>(dst + c + 4) = make_float4(e.x, e.y, f.x, f.y);
    }
  }
}
template <int HD>
__device__ __forceinline__ void store_row_bf16(bf16* dst, const float (&v)[HD], float mul) {
  if constexpr (HD == 4) {
    *reinterpret_cast<uint2*>(dst) = make_uint2(gwd_pack_bf16x2(v[0] * mul, v[1] * mul), gwd_pack_bf16x2(v[2] * mul, v[3] * mul));
  } else {
#pragma unroll
    for (int c = 0; c < HD; c += 8)
      *reinterpret_cast<uint4*>(dst + c) = make_uint4(gwd_pack_bf16x2(v[c] * mul, v[c + 1] * mul), gwd_pack_bf16x2(v[c + 2] * mul, v[c + 3] * mul),
                                                      gwd_pack_bf16x2(v[c + 4] * mul, v[c + 5] * mul), gwd_pack_bf16x2(v[c + 6] * mul, v[c + 7] * mul));
  }
}

// PACK: P and dS of a (query, key) pair live in ONE 32-bit word (bf16 x 2) instead of two fp32 arrays: 73 instead of 93 KB of
// shared memory per CTA (three CTAs per SM instead of two -- the kernel is latency bound at 8 warps per SM) and one shared-memory
// access instead of two on both sides of the hand-over to the column pass
template <int HD, bool PACK>
__global__ void __launch_bounds__(128) gwd_window_attention_bwd_rows_kernel(const WinBwdParams p) {
  extern __shared__ __align__(16) float smr[];
  constexpr int N = kWN, NN = kWN * kWN;
  float* sBias = smr;                         // [2][N*N]
  float* sMask = sBias + 2 * NN;              // [N*N]
  float* slot0 = sMask + NN + 1;              // keep 16-byte alignment: 3 * NN + 1 = 7204 floats
  const int slot = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int h = 2 * blockIdx.x + slot;
  constexpr int kSlot = PACK ? 4 * N * HD + 2 * NN + 2 : 4 * N * HD + 3 * NN + 1;      // multiples of 4 (16-byte rows)
  float* sq = slot0 + slot * kSlot;
  float* sk = sq + N * HD;
  float* sv = sk + N * HD;
  float* sdo = sv + N * HD;
  float* sP = sdo + N * HD;                   // [N][N]
  float* sdS = sP + NN;                       // (PACK: unused, sP holds the pairs)
  uint32_t* sPD = reinterpret_cast<uint32_t*>(sP);
  float* sDb = PACK ? sP + NN : sdS + NN;                      // [N][N] bias-gradient accumulator of this slot's head (row t owned by thread t)
  for (int e = threadIdx.x; e < 2 * NN; e += 128)
    sBias[e] = p.bias ? p.bias[static_cast<int64_t>(2 * blockIdx.x) * NN + e] : 0.f;
  for (int e = t; e < NN; e += 64) sDb[e] = 0.f;
  const bool row = t < N;
  const float* bias_row = sBias + slot * NN + (row ? t : 0) * N;

  for (int item = blockIdx.y; item < p.items; item += gridDim.y) {
    const int64_t row0 = static_cast<int64_t>(item) * N;
    if (row) {
      const bf16* r = p.qkv + (row0 + t) * p.qkv_rs + h * HD;
      load_row_bf16<HD>(r, sq + t * HD);
      load_row_bf16<HD>(r + p.C, sk + t * HD);
      load_row_bf16<HD>(r + 2 * p.C, sv + t * HD);
      load_row_bf16<HD>(p.d_o + (row0 + t) * p.do_rs + h * HD, sdo + t * HD);
    }
    if (p.mask) {
      const float* mw = p.mask + static_cast<int64_t>(item % p.mask_windows) * NN;
      for (int e = threadIdx.x; e < NN; e += 128) sMask[e] = mw[e];
    }
    __syncthreads();
    if (row) {
      float q[HD], g[HD];
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(sq + t * HD + c), b = *reinterpret_cast<const float4*>(sdo + t * HD + c);
        q[c] = a.x; q[c + 1] = a.y; q[c + 2] = a.z; q[c + 3] = a.w;
        g[c] = b.x; g[c + 1] = b.y; g[c + 2] = b.z; g[c + 3] = b.w;
      }
      // pass 1 (online): row max m, l = sum_j e^(s_j - m), acc = sum_j e^(s_j - m) dP_j  ->  D = acc / l
      float m = -INFINITY, l = 0.f, acc = 0.f;
#pragma unroll 7
      for (int j = 0; j < N; ++j) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          const float4 kk = *reinterpret_cast<const float4*>(sk + j * HD + c), vv = *reinterpret_cast<const float4*>(sv + j * HD + c);
          a = fmaf(q[c], kk.x, a); a = fmaf(q[c + 1], kk.y, a); a = fmaf(q[c + 2], kk.z, a); a = fmaf(q[c + 3], kk.w, a);
          b = fmaf(g[c], vv.x, b); b = fmaf(g[c + 1], vv.y, b); b = fmaf(g[c + 2], vv.z, b); b = fmaf(g[c + 3], vv.w, b);
        }
        a = fmaf(a, p.scale, bias_row[j]);
        if (p.mask) a += sMask[t * N + j];
        const float mn = fmaxf(m, a);
        const float corr = __expf(m - mn), e = __expf(a - mn);
        l = fmaf(l, corr, e);
        acc = fmaf(acc, corr, e * b);
        m = mn;
      }
      const float inv = 1.f / l, dsum = acc * inv;
      // pass 2: P_j, dS_j = P_j (dP_j - D) -> shared memory (for the column products), bias gradient, dq
      float dq[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = 0.f;
#pragma unroll 7
      for (int j = 0; j < N; ++j) {
        float a = 0.f, b = 0.f;
        float4 kk[HD / 4];
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          kk[c / 4] = *reinterpret_cast<const float4*>(sk + j * HD + c);
          const float4 vv = *reinterpret_cast<const float4*>(sv + j * HD + c);
          a = fmaf(q[c], kk[c / 4].x, a); a = fmaf(q[c + 1], kk[c / 4].y, a); a = fmaf(q[c + 2], kk[c / 4].z, a); a = fmaf(q[c + 3], kk[c / 4].w, a);
          b = fmaf(g[c], vv.x, b); b = fmaf(g[c + 1], vv.y, b); b = fmaf(g[c + 2], vv.z, b); b = fmaf(g[c + 3], vv.w, b);
        }
        a = fmaf(a, p.scale, bias_row[j]);
        if (p.mask) a += sMask[t * N + j];
        const float pr = __expf(a - m) * inv;
        const float ds = pr * (b - dsum);
        if (PACK) sPD[t * N + j] = gwd_pack_bf16x2(pr, ds);
        else { sP[t * N + j] = pr; sdS[t * N + j] = ds; }
        sDb[t * N + j] += ds;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          dq[c] = fmaf(ds, kk[c / 4].x, dq[c]); dq[c + 1] = fmaf(ds, kk[c / 4].y, dq[c + 1]);
          dq[c + 2] = fmaf(ds, kk[c / 4].z, dq[c + 2]); dq[c + 3] = fmaf(ds, kk[c / 4].w, dq[c + 3]);
        }
      }
      store_row_bf16<HD>(p.dqkv + (row0 + t) * p.dqkv_rs + h * HD, dq, p.scale);
    }
    __syncthreads();
    if (row) {      // thread = key / value row j: column sums over the query rows
      float dk[HD], dv[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll 7
      for (int i = 0; i < N; ++i) {
        float ds, pr;
        if (PACK) { const float2 pd = gwd_unpack_bf16x2(sPD[i * N + t]); pr = pd.x; ds = pd.y; }
        else { ds = sdS[i * N + t]; pr = sP[i * N + t]; }
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          const float4 qq = *reinterpret_cast<const float4*>(sq + i * HD + c), gg = *reinterpret_cast<const float4*>(sdo + i * HD + c);
          dk[c] = fmaf(ds, qq.x, dk[c]); dk[c + 1] = fmaf(ds, qq.y, dk[c + 1]); dk[c + 2] = fmaf(ds, qq.z, dk[c + 2]); dk[c + 3] = fmaf(ds, qq.w, dk[c + 3]);
          dv[c] = fmaf(pr, gg.x, dv[c]); dv[c + 1] = fmaf(pr, gg.y, dv[c + 1]); dv[c + 2] = fmaf(pr, gg.z, dv[c + 2]); dv[c + 3] = fmaf(pr, gg.w, dv[c + 3]);
        }
      }
      bf16* o = p.dqkv + (row0 + t) * p.dqkv_rs + h * HD;
      store_row_bf16<HD>(o + p.C, dk, p.scale);
      store_row_bf16<HD>(o + 2 * p.C, dv, 1.f);
    }
    __syncthreads();
  }
  if (p.dbias) {
    __syncthreads();
    float* dst = p.dbias + static_cast<int64_t>(h) * NN;
    for (int e = t; e < NN; e += 64) atomicAdd(dst + e, sDb[e]);
  }
}

template <int HD, bool PACK>
int launch_rows_impl(const WinBwdParams& p, cudaStream_t stream) {
  const size_t slot = PACK ? 4 * kWN * HD + 2 * kWN * kWN + 2 : 4 * kWN * HD + 3 * kWN * kWN + 1;
  const size_t smem = sizeof(float) * (3 * kWN * kWN + 1 + 2 * slot);
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_window_attention_bwd_rows_kernel<HD, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    int nb = 0;
    GWD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, gwd_window_attention_bwd_rows_kernel<HD, PACK>, 128, smem));
    ctas_per_sm = std::max(1, nb);
  }
  const int pairs = p.heads / 2;
  int per = std::max(1, (gwd_num_sms() * ctas_per_sm) / pairs);
  per = std::min(per, p.items);
  gwd_window_attention_bwd_rows_kernel<HD, PACK><<<dim3(pairs, per), 128, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}
template <int HD>
int launch_rows(const WinBwdParams& p, cudaStream_t stream) {
  static const bool pack = [] { const char* e = getenv("GWD_WINBWD_PACK"); return !(e && e[0] == '0'); }();
  return pack ? launch_rows_impl<HD, true>(p, stream) : launch_rows_impl<HD, false>(p, stream);
}

// Warp-per-problem version of the class-token channel attention backward: the CTA-per-(window, head) kernel above spends its
// time in four block-wide barriers around a few hundred FMAs per thread.  Here every warp owns a (window, head) problem in its
// own shared-memory slice (__syncwarp only), loads and stores move 8-byte vectors (td and tc are multiples of 4), and the
// outputs are computed four channels at a time.
template <int TD>      // token-query channels per head (4 in the reference configuration); rows R = 2 TD
__global__ void __launch_bounds__(128) gwd_token_attention_bwd_warp_kernel(const TokBwdParams p) {
  extern __shared__ __align__(16) float smw[];
  constexpr int R = 2 * TD;
  const int N = p.N, tc = p.tc;
  const int tcp = (tc + 4) % 8 == 0 ? tc + 8 : tc + 4;   // row pitch of the key / value tiles: 16-byte aligned, = 4 mod 8 (banks)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = smw + warp * (2 * N * R + 2 * N * tcp + 2 * R * tcp);   // [N][R]
  float* go = q + N * R;                                 // [N][R]
  float* k = go + N * R;                                 // [N][tcp]
  float* v = k + N * tcp;                                // [N][tcp]
  float* a = v + N * tcp;                                // [R][tcp]
  float* dz = a + R * tcp;                               // [R][tcp]
  const int c4n = tc >> 2;
  auto ld4 = [](const bf16* src, float* dst) {
    const uint2 u = *reinterpret_cast<const uint2*>(src);
    const float2 x = gwd_unpack_bf16x2(u.x), y = gwd_unpack_bf16x2(u.y);
    *reinterpret_cast<float4*>(dst) = make_float4(x.x, x.y, y.x, y.y);
  };
  const int64_t total = static_cast<int64_t>(p.items) * p.heads;
  for (int64_t wh = static_cast<int64_t>(blockIdx.x) * 4 + warp; wh < total; wh += static_cast<int64_t>(gridDim.x) * 4) {
    const int item = static_cast<int>(wh / p.heads), h = static_cast<int>(wh - static_cast<int64_t>(item) * p.heads);
    const int64_t row0 = static_cast<int64_t>(item) * N;
    for (int idx = lane; idx < N * (R / 4); idx += 32) {           // query / output-gradient pieces: [depth TD | seg TD]
      const int n = idx / (R / 4), r4 = (idx - n * (R / 4)) * 4;
      const bool seg = r4 >= TD;
      const int c = h * TD + (seg ? r4 - TD : r4);
      ld4((seg ? p.sq : p.dq) + (row0 + n) * p.q_rs + c, q + n * R + r4);
      ld4((seg ? p.d_sout : p.d_dout) + (row0 + n) * p.o_rs + c, go + n * R + r4);
    }
    for (int idx = lane; idx < N * c4n; idx += 32) {
      const int n = idx / c4n, c = (idx - n * c4n) * 4;
      ld4(p.tk + (row0 + n) * p.k_rs + h * tc + c, k + n * tcp + c);
      ld4(p.tv + (row0 + n) * p.v_rs + h * tc + c, v + n * tcp + c);
    }
    __syncwarp();
    // scores[r][c..c+3] and d a = sum_n go[n][r] v[n][c..c+3]
    for (int e = lane; e < R * c4n; e += 32) {
      const int r = e / c4n, c = (e - r * c4n) * 4;
      float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), da = sc;
      for (int n = 0; n < N; ++n) {
        const float qq = q[n * R + r], gg = go[n * R + r];
        const float4 kk = *reinterpret_cast<const float4*>(k + n * tcp + c), vv = *reinterpret_cast<const float4*>(v + n * tcp + c);
        sc.x = fmaf(qq, kk.x, sc.x); sc.y = fmaf(qq, kk.y, sc.y); sc.z = fmaf(qq, kk.z, sc.z); sc.w = fmaf(qq, kk.w, sc.w);
        da.x = fmaf(gg, vv.x, da.x); da.y = fmaf(gg, vv.y, da.y); da.z = fmaf(gg, vv.z, da.z); da.w = fmaf(gg, vv.w, da.w);
      }
      *reinterpret_cast<float4*>(a + r * tcp + c) = make_float4(sc.x * p.scale, sc.y * p.scale, sc.z * p.scale, sc.w * p.scale);
      *reinterpret_cast<float4*>(dz + r * tcp + c) = da;
    }
    __syncwarp();
    if (lane < R) {      // soft-max over the key channels of row `lane` and its backward
      float* ar = a + lane * tcp;
      float* dr = dz + lane * tcp;
      float mx = -INFINITY;
      for (int c = 0; c < tc; ++c) mx = fmaxf(mx, ar[c]);
      float sum = 0.f;
      for (int c = 0; c < tc; ++c) { const float e = __expf(ar[c] - mx); ar[c] = e; sum += e; }
      const float inv = 1.f / sum;
      float dsum = 0.f;
      for (int c = 0; c < tc; ++c) { ar[c] *= inv; dsum = fmaf(ar[c], dr[c], dsum); }
      for (int c = 0; c < tc; ++c) dr[c] = ar[c] * (dr[c] - dsum) * p.scale;
    }
    __syncwarp();
    // d q[n][r..r+3] = sum_c dz[r..r+3][c] k[n][c]
    for (int idx = lane; idx < N * (R / 4); idx += 32) {
      const int n = idx / (R / 4), r4 = (idx - n * (R / 4)) * 4;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      for (int c = 0; c < tc; ++c) {
        const float kk = k[n * tcp + c];
        o0 = fmaf(dz[r4 * tcp + c], kk, o0); o1 = fmaf(dz[(r4 + 1) * tcp + c], kk, o1);
        o2 = fmaf(dz[(r4 + 2) * tcp + c], kk, o2); o3 = fmaf(dz[(r4 + 3) * tcp + c], kk, o3);
      }
      const bool seg = r4 >= TD;
      bf16* dst = (seg ? p.g_sq : p.g_dq) + (row0 + n) * p.gq_rs + h * TD + (seg ? r4 - TD : r4);
      *reinterpret_cast<uint2*>(dst) = make_uint2(gwd_pack_bf16x2(o0, o1), gwd_pack_bf16x2(o2, o3));
    }
    // d k[n][c..c+3] = sum_r dz[r][c..c+3] q[n][r];  d v[n][c..c+3] = sum_r a[r][c..c+3] go[n][r]
    for (int idx = lane; idx < N * c4n; idx += 32) {
      const int n = idx / c4n, c = (idx - n * c4n) * 4;
      float4 gk = make_float4(0.f, 0.f, 0.f, 0.f), gv = gk;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float qq = q[n * R + r], gg = go[n * R + r];
        const float4 zz = *reinterpret_cast<const float4*>(dz + r * tcp + c), aa = *reinterpret_cast<const float4*>(a + r * tcp + c);
        gk.x = fmaf(zz.x, qq, gk.x); gk.y = fmaf(zz.y, qq, gk.y); gk.z = fmaf(zz.z, qq, gk.z); gk.w = fmaf(zz.w, qq, gk.w);
        gv.x = fmaf(aa.x, gg, gv.x); gv.y = fmaf(aa.y, gg, gv.y); gv.z = fmaf(aa.z, gg, gv.z); gv.w = fmaf(aa.w, gg, gv.w);
      }
      *reinterpret_cast<uint2*>(p.g_tk + (row0 + n) * p.gk_rs + h * tc + c) = make_uint2(gwd_pack_bf16x2(gk.x, gk.y), gwd_pack_bf16x2(gk.z, gk.w));
      *reinterpret_cast<uint2*>(p.g_tv + (row0 + n) * p.gv_rs + h * tc + c) = make_uint2(gwd_pack_bf16x2(gv.x, gv.y), gwd_pack_bf16x2(gv.z, gv.w));
    }
    __syncwarp();
  }
}

}  // namespace

#define GWD_STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int gwd_window_attention_bwd(const void* qkv, int64_t qkv_rs, const void* d_o, int64_t do_rs, void* dqkv, int64_t dqkv_rs,
                                        const float* bias, const float* mask, int32_t mask_windows, float* dbias, int32_t items,
                                        int32_t heads, int32_t N, int32_t hd, float scale, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(qkv && d_o && dqkv, "gwd_window_attention_bwd: null pointer");
  GWD_CHECK_ARG(items > 0 && heads > 0 && heads <= 65535 && N > 0 && N <= kMaxN && hd > 0 && hd <= kMaxHd,
                "gwd_window_attention_bwd: needs N <= 64 tokens and head dim <= 32 (N=%d hd=%d)", N, hd);
  GWD_CHECK_ARG(mask == nullptr || mask_windows > 0, "gwd_window_attention_bwd: mask without mask_windows");
  WinBwdParams p;
  p.qkv = static_cast<const bf16*>(qkv); p.d_o = static_cast<const bf16*>(d_o); p.dqkv = static_cast<bf16*>(dqkv);
  p.bias = bias; p.mask = mask; p.dbias = dbias;
  p.items = items; p.heads = heads; p.N = N; p.hd = hd; p.C = heads * hd; p.mask_windows = mask_windows > 0 ? mask_windows : 1;
  p.qkv_rs = qkv_rs; p.do_rs = do_rs; p.dqkv_rs = dqkv_rs; p.scale = scale;
  GWD_CHECK_ARG(qkv_rs >= 3 * p.C && dqkv_rs >= 3 * p.C && do_rs >= p.C, "gwd_window_attention_bwd: row strides too small");
  static const bool generic_only = [] { const char* e = getenv("GWD_WINBWD"); return e && e[0] == 'g'; }();
  const bool aligned = ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0 &&
                       qkv_rs % 8 == 0 && do_rs % 8 == 0 && dqkv_rs % 8 == 0;
  if (!generic_only && N == kWN && heads % 2 == 0 && aligned) {
    if (hd == 4) return launch_rows<4>(p, stream);
    if (hd == 8) return launch_rows<8>(p, stream);
    if (hd == 16) return launch_rows<16>(p, stream);
  }
  const size_t smem = sizeof(float) * (4 * N * (hd + 1) + 2 * N * (N + 1) + N);
  static bool attr_set = false;
  if (!attr_set) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_window_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  int per_head = std::max(1, (gwd_num_sms() * 4) / heads);
  per_head = std::min(per_head, items);
  gwd_window_attention_bwd_kernel<<<dim3(heads, per_head), 128, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_token_attention_bwd(const void* dq, const void* sq, const void* tk, const void* tv, const void* d_dout,
                                       const void* d_sout, void* g_dq, void* g_sq, void* g_tk, void* g_tv, int32_t items, int32_t N,
                                       int32_t heads, int32_t td, int32_t tc, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs,
                                       int64_t gq_rs, int64_t gk_rs, int64_t gv_rs, float scale, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(dq && sq && tk && tv && d_dout && d_sout && g_dq && g_sq && g_tk && g_tv, "gwd_token_attention_bwd: null pointer");
  GWD_CHECK_ARG(items > 0 && heads > 0 && N > 0 && N <= kMaxN && td > 0 && 2 * td <= kTokR && tc > 0 && tc <= kTokC,
                "gwd_token_attention_bwd: needs N <= 64, 2 td <= 16, tc <= 32 (N=%d td=%d tc=%d)", N, td, tc);
  TokBwdParams p;
  p.dq = static_cast<const bf16*>(dq); p.sq = static_cast<const bf16*>(sq); p.tk = static_cast<const bf16*>(tk);
  p.tv = static_cast<const bf16*>(tv); p.d_dout = static_cast<const bf16*>(d_dout); p.d_sout = static_cast<const bf16*>(d_sout);
  p.g_dq = static_cast<bf16*>(g_dq); p.g_sq = static_cast<bf16*>(g_sq); p.g_tk = static_cast<bf16*>(g_tk); p.g_tv = static_cast<bf16*>(g_tv);
  p.items = items; p.N = N; p.heads = heads; p.td = td; p.tc = tc;
  p.q_rs = q_rs; p.k_rs = k_rs; p.v_rs = v_rs; p.o_rs = o_rs; p.gq_rs = gq_rs; p.gk_rs = gk_rs; p.gv_rs = gv_rs; p.scale = scale;
  const int64_t units = static_cast<int64_t>(items) * heads;
  static const bool generic_only = [] { const char* e = getenv("GWD_TOKBWD"); return e && e[0] == 'g'; }();
  const bool vec_ok = td == 4 && tc % 4 == 0 && q_rs % 4 == 0 && k_rs % 4 == 0 && v_rs % 4 == 0 && o_rs % 4 == 0 && gq_rs % 4 == 0 &&
                      gk_rs % 4 == 0 && gv_rs % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(sq) | reinterpret_cast<uintptr_t>(tk) |
                        reinterpret_cast<uintptr_t>(tv) | reinterpret_cast<uintptr_t>(d_dout) | reinterpret_cast<uintptr_t>(d_sout) |
                        reinterpret_cast<uintptr_t>(g_dq) | reinterpret_cast<uintptr_t>(g_sq) | reinterpret_cast<uintptr_t>(g_tk) |
                        reinterpret_cast<uintptr_t>(g_tv)) & 7) == 0;
  if (vec_ok && !generic_only) {      // warp per (window, head): the reference configuration (token_dim 64 over 16 heads)
    const int tcp = (tc + 4) % 8 == 0 ? tc + 8 : tc + 4;
    const size_t smem = sizeof(float) * 4 * (2 * N * 8 + 2 * N * tcp + 2 * 8 * tcp);
    static bool attr_set = false;
    if (!attr_set) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_token_attention_bwd_warp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
    const unsigned g = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(units, 4), static_cast<int64_t>(gwd_num_sms()) * 4));
    gwd_token_attention_bwd_warp_kernel<4><<<g, 128, smem, stream>>>(p);
    GWD_LAUNCHED();
    return GWD_OK;
  }
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(units, static_cast<int64_t>(gwd_num_sms()) * 8));
  gwd_token_attention_bwd_kernel<<<grid, 128, 0, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}
