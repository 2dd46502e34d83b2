// gwd_train_line.cu -- backward of the "glass-structure context" of the 1/32 line-window attention
// (WindowAttention.forward, src/models/multiscale_transformerr.py:267-332, under torch.autograd) and the small
// resampling kernels the backbone backward needs for its stride-2 convolutions.
//
//   ref_k = mu + exp(logsigma) * ref_qk(x_ref)[:, :D]                          gwd_ref_affine / gwd_ref_affine_bwd
//   a0[b,h,t,r] = scale * q[t,h,:] . ref_k[b,r,h,:]                            gwd_ref_scores_bwd
//   a_{i+1} = a_i + gelu(instance_norm(conv3x3_{16->16}(a_i) + bias)), i<3     gwd_ref_diffuse_bwd (+ gwd_diffuse_filter_pack)
//   q_new[t,h,:] = scale * softmax_r(a3[b,h,t,:]) @ ref_v[b,:,h,:]             gwd_ref_requery_bwd
//
// Everything here lives at the 1/32 scale (T = 441 window tokens, R = 40 / 60 reference points, 16 heads of 32 channels
// per image): ~0.1 GFLOP per launch, so the kernels are fp32 CUDA-core kernels organised for coalesced global traffic and
// conflict-free shared memory, one CTA per (image, head) where a reduction over the tokens is needed (deterministic, no
// atomics); only the 16x16x3x3 filter gradient is reduced with atomics (2 320 values per persistent CTA).
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kHeads = 16;          // channels of the diffusion convolution = attention heads of the 1/32 stage
constexpr int kFilt = kHeads * kHeads * 9;
constexpr int kTokTile = 128;

__device__ __forceinline__ float gelu_grad(float v) { return gwd_gelu_grad(v); }   // derivative of the forward's tanh-form GELU

// ------------------------------------------------------------------------------------------------
// filter re-layout: FlatModule keeps the 16x16x3x3 filter as [tap = kx*3+ky][oc][ic]; the convolution kernels of
// gwd_attn.cu read [oc][ic][ky][kx] followed by the 16 biases.  fwd = the filter itself, bwd = its adjoint
// (channels transposed, taps flipped, zero biases): conv(d_out, bwd) is the data gradient.
// ------------------------------------------------------------------------------------------------
__global__ void gwd_diffuse_filter_pack_kernel(const float* __restrict__ w_phys, const float* __restrict__ bias,
                                               float* __restrict__ fwd, float* __restrict__ bwd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kFilt) {
    const int oc = i / (kHeads * 9), rem = i - oc * kHeads * 9;
    const int ic = rem / 9, k = rem - ic * 9;
    const int ky = k / 3, kx = k - ky * 3;
    const float w = w_phys[((kx * 3 + ky) * kHeads + oc) * kHeads + ic];
    fwd[i] = w;
    bwd[(ic * kHeads + oc) * 9 + (2 - ky) * 3 + (2 - kx)] = w;
  } else if (i < kFilt + kHeads) {
    fwd[i] = bias[i - kFilt];
    bwd[i] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// ref_k = mu + exp(logsigma) * ref[:, :D]      (multiscale_transformerr.py:281-288)
// ------------------------------------------------------------------------------------------------
__global__ void gwd_ref_affine_kernel(const float* __restrict__ ref, int64_t ref_rs, const float* __restrict__ mu,
                                      const float* __restrict__ ls, float* __restrict__ out, int64_t rows, int D) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= rows * D) return;
  const int64_t r = i / D;
  const int c = static_cast<int>(i - r * D);
  out[i] = fmaf(__expf(ls[c]), ref[r * ref_rs + c], mu[c]);
}

// d_kv fp32 [rows, 2D] = (d ref_k | d ref_v)  ->  d_ref bf16 [rows, 2D] = (d ref_k * sigma | d ref_v);
// dmu[c] += sum_rows d ref_k, dls[c] += sum_rows d ref_k * sigma * ref[:, c].  A CTA owns 32 columns (one warp-wide coalesced
// segment per row) and splits the rows over its 8 warps; the 8 partial sums meet in shared memory (fixed order: deterministic).
__global__ void __launch_bounds__(256) gwd_ref_affine_bwd_kernel(const float* __restrict__ d_kv, const float* __restrict__ ref,
                                                                int64_t ref_rs, const float* __restrict__ ls, bf16* __restrict__ d_ref,
                                                                float* __restrict__ dmu, float* __restrict__ dls, int rows, int D) {
  __shared__ float red[2][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  if (c >= 2 * D) return;                       // (2D is a multiple of 32: whole CTAs leave together)
  const bool is_k = c < D;
  const float sig = is_k ? __expf(ls[c]) : 1.f;
  float sm = 0.f, sl = 0.f;
  for (int r = warp; r < rows; r += 8) {
    const float g = d_kv[static_cast<int64_t>(r) * 2 * D + c];
    if (is_k) {
      sm += g;
      sl = fmaf(g * sig, ref[r * ref_rs + c], sl);
    }
    d_ref[static_cast<int64_t>(r) * 2 * D + c] = __float2bfloat16(g * sig);
  }
  red[0][warp][lane] = sm;
  red[1][warp][lane] = sl;
  __syncthreads();
  if (warp == 0 && is_k) {
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[0][w][lane]; b2 += red[1][w][lane]; }
    dmu[c] += a;
    dls[c] += b2;
  }
}

// ------------------------------------------------------------------------------------------------
// shared structure of the two score-side backward kernels: one CTA per (head, image, tile of 128 tokens).  The [tile][R] score
// rows are staged in shared memory (coalesced, padded rows) next to the [tile][hd] bf16 token rows; a thread owns a token for
// the per-token part, then the threads re-map to the R x hd outputs of the token reduction (a warp = one reference point x 32
// channels: broadcast + stride-1 shared-memory reads) and add the tile's partial sums to the (pre-zeroed) output with one
// atomic per value: 4 tiles per (image, head) at 480x640.
// ------------------------------------------------------------------------------------------------
template <bool SOFTMAX>
__global__ void __launch_bounds__(kTokTile) gwd_ref_bwd_kernel(
    const float* __restrict__ S,        // SOFTMAX: a3 (scores ahead of the soft-max); else d a0
    const float* __restrict__ RV, int64_t rv_rs,   // SOFTMAX: ref_v; else ref_k          fp32 [B*R, *]
    const bf16* __restrict__ X, int64_t x_rs,      // SOFTMAX: d q_new; else q            bf16 [B*T, *]
    float* __restrict__ dS,             // SOFTMAX: d a3 out; else unused
    bf16* __restrict__ dX, int64_t dx_rs,          // SOFTMAX: unused; else d q out
    float* __restrict__ dRV, int64_t drv_rs,       // d ref_v / d ref_k out, fp32 [B*R, *]
    int T, int heads, int hd, int R, float scale) {
  extern __shared__ __align__(16) float sm[];
  float* rv = sm;                        // [R][hd]
  float* st = rv + R * hd;               // [128][R + 1] score rows (become A / stay dS)
  float* xt = st + kTokTile * (R + 1);   // [128][hd + 1] token rows
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < R * hd; i += kTokTile) {
    const int r = i / hd, d = i - r * hd;
    rv[i] = RV[(static_cast<int64_t>(b) * R + r) * rv_rs + h * hd + d];
  }
  const int nout = R * hd;
  float acc[16];                         // R * hd / 128 <= 16 outputs of the token reduction per thread (R <= 64, hd <= 32)
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  {
    const int tok0 = blockIdx.z * kTokTile;
    const int ntok = min(kTokTile, T - tok0);
    const int64_t sbase = ((static_cast<int64_t>(b) * heads + h) * T + tok0) * R;
    for (int i = tid; i < ntok * R; i += kTokTile) st[(i / R) * (R + 1) + i % R] = __ldg(S + sbase + i);
    for (int i = tid; i < ntok * hd; i += kTokTile) {
      const int tl = i / hd, d = i - tl * hd;
      xt[tl * (hd + 1) + d] = __bfloat162float(X[(static_cast<int64_t>(b) * T + tok0 + tl) * x_rs + h * hd + d]);
    }
    __syncthreads();
    if (tid < ntok) {
      float* row = st + tid * (R + 1);
      const float* xr = xt + tid * (hd + 1);
      if (SOFTMAX) {
        // A = softmax(row); dA[r] = scale * x . rv[r]; dS = A (dA - sum A dA); row <- A (for the d ref_v reduction)
        float mx = -INFINITY;
        for (int r = 0; r < R; ++r) mx = fmaxf(mx, row[r]);
        float sum = 0.f;
        for (int r = 0; r < R; ++r) { const float e = __expf(row[r] - mx); row[r] = e; sum += e; }
        const float inv = 1.f / sum;
        float dot = 0.f;
        float* out = dS + sbase + static_cast<int64_t>(tid) * R;     // per-thread rows: R floats apart (tiny tensor)
        for (int r = 0; r < R; ++r) {
          float da = 0.f;
          for (int d = 0; d < hd; ++d) da = fmaf(xr[d], rv[r * hd + d], da);
          da *= scale;
          const float a = row[r] * inv;
          row[r] = a;
          out[r] = da;                   // parked; finished below once the row dot product is known
          dot = fmaf(a, da, dot);
        }
        for (int r = 0; r < R; ++r) out[r] = row[r] * (out[r] - dot);
      } else {
        // d q[d] = scale * sum_r dS[r] rk[r][d]
        bf16* o = dX + (static_cast<int64_t>(b) * T + tok0 + tid) * dx_rs + h * hd;
        for (int d = 0; d < hd; d += 2) {
          float a0 = 0.f, a1 = 0.f;
          for (int r = 0; r < R; ++r) {
            a0 = fmaf(row[r], rv[r * hd + d], a0);
            a1 = fmaf(row[r], rv[r * hd + d + 1], a1);
          }
          *reinterpret_cast<uint32_t*>(o + d) = gwd_pack_bf16x2(a0 * scale, a1 * scale);
        }
      }
    }
    __syncthreads();
    // token reduction: out[r][d] += sum_tok st[tok][r] * xt[tok][d]
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int o = tid + e * kTokTile;
      if (o < nout) {
        const int r = o / hd, d = o - r * hd;
        float a = acc[e];
        for (int tl = 0; tl < ntok; ++tl) a = fmaf(st[tl * (R + 1) + r], xt[tl * (hd + 1) + d], a);
        acc[e] = a;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int o = tid + e * kTokTile;
    if (o < nout) {
      const int r = o / hd, d = o - r * hd;
      atomicAdd(dRV + (static_cast<int64_t>(b) * R + r) * drv_rs + h * hd + d, acc[e] * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// diffusion round backward, normalisation part.  Forward: y = (raw - mean) * rstd over the [P,R] plane of one (image,
// channel), a_out = a_in + gelu(y).  With g = d a_out:  dy = g gelu'(y),
// d raw = rstd (dy - mean(dy) - y mean(dy y)).   Phase 1 reduces (sum dy, sum dy y) per plane, phase 2 writes d raw.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void plane_stats(const double* stats, int img, int per_img, float& mean, float& rstd) {
  const double m = stats[img * 2] / per_img;
  const double var = stats[img * 2 + 1] / per_img - m * m;
  mean = static_cast<float>(m);
  rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
}

__global__ void __launch_bounds__(256) gwd_diffuse_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ raw,
                                                                    const double* __restrict__ stats, double* __restrict__ stats2,
                                                                    int per_img) {
  const int img = blockIdx.y;
  float mean, rstd;
  plane_stats(stats, img, per_img, mean, rstd);
  const int64_t off = static_cast<int64_t>(img) * per_img;
  float s1 = 0.f, s2 = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const float y = (raw[off + i] - mean) * rstd;
    const float dy = g[off + i] * gelu_grad(y);
    s1 += dy;
    s2 = fmaf(dy, y, s2);
  }
  __shared__ float red[2][8];
  s1 = gwd_warp_sum(s1);
  s2 = gwd_warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += static_cast<double>(red[threadIdx.x][w]);
    atomicAdd(&stats2[img * 2 + threadIdx.x], v);
  }
}

__global__ void __launch_bounds__(256) gwd_diffuse_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ raw,
                                                                   const double* __restrict__ stats, const double* __restrict__ stats2,
                                                                   float* __restrict__ d_raw, int per_img) {
  const int img = blockIdx.y;
  float mean, rstd;
  plane_stats(stats, img, per_img, mean, rstd);
  const float m1 = static_cast<float>(stats2[img * 2] / per_img), m2 = static_cast<float>(stats2[img * 2 + 1] / per_img);
  const int64_t off = static_cast<int64_t>(img) * per_img;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += gridDim.x * blockDim.x) {
    const float y = (raw[off + i] - mean) * rstd;
    const float dy = g[off + i] * gelu_grad(y);
    d_raw[off + i] = rstd * (dy - m1 - y * m2);
  }
}

// ------------------------------------------------------------------------------------------------
// filter gradient of the diffusion convolution: dw[kx*3+ky][oc][ic] += sum_{b,y,x} d_raw[b,oc,y,x] a_in[b,ic,y+ky-1,x+kx-1],
// db[oc] += sum d_raw[b,oc].  Persistent CTAs over (image, band of 7 rows) items; thread t = (oc, ic) keeps its nine taps in
// registers over all its items (the 16 input planes are (9 * (R+2)) words apart = distinct banks for R = 40 / 60), one
// atomic per value and CTA at the end.
// ------------------------------------------------------------------------------------------------
constexpr int kBand = 7;
__global__ void __launch_bounds__(256) gwd_diffuse_wgrad_kernel(const float* __restrict__ d_raw, const float* __restrict__ a_in,
                                                               float* __restrict__ dw, float* __restrict__ db, int B, int P, int R) {
  extern __shared__ __align__(16) float sm[];
  const int TR = kBand + 2, TC = R + 2;
  float* tin = sm;                           // [16][TR][TC] zero padded
  float* tdr = sm + kHeads * TR * TC;        // [16][kBand][R]
  const int tid = threadIdx.x, oc = tid >> 4, ic = tid & 15;
  float acc[9], accb = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  const int bands = (P + kBand - 1) / kBand;
  for (int item = blockIdx.x; item < bands * B; item += gridDim.x) {
    const int b = item / bands, y0 = (item - b * bands) * kBand;
    const int rows = min(kBand, P - y0);
    __syncthreads();
    for (int i = tid; i < kHeads * TR * TC; i += 256) {
      const int c = i / (TR * TC), rem = i - c * TR * TC;
      const int ty = rem / TC, tx = rem - ty * TC;
      const int y = y0 + ty - 1, x = tx - 1;
      const bool ok = y >= 0 && y < P && ty < rows + 2 && x >= 0 && x < R;
      tin[i] = ok ? a_in[((static_cast<int64_t>(b) * kHeads + c) * P + y) * R + x] : 0.f;
    }
    for (int i = tid; i < kHeads * kBand * R; i += 256) {
      const int c = i / (kBand * R), rem = i - c * kBand * R;
      const int ty = rem / R;
      tdr[i] = ty < rows ? d_raw[((static_cast<int64_t>(b) * kHeads + c) * P + y0) * R + rem] : 0.f;
    }
    __syncthreads();
    const float* pi = tin + ic * TR * TC;
    const float* pd = tdr + oc * kBand * R;
    for (int ty = 0; ty < rows; ++ty)
      for (int tx = 0; tx < R; ++tx) {
        const float d = pd[ty * R + tx];
        const float* w = pi + ty * TC + tx;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc[kx * 3 + ky] = fmaf(d, w[ky * TC + kx], acc[kx * 3 + ky]);
        accb += d;
      }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) atomicAdd(dw + (k * kHeads + oc) * kHeads + ic, acc[k]);
  if (ic == 0) atomicAdd(db + oc, accb);
}

// ------------------------------------------------------------------------------------------------
// stride-2 helpers of the backbone backward (bf16 channels-last, 16-byte vectors)
//   subsample2 : y[b,i,j,:] = x[b,2i,2j,:]                     (the input of a stride-2 1x1 projection, for its weight gradient)
//   zero_stuff2: y[b,i,j,:] = add[b,i,j,:] + (i,j both even ? s[b,i/2,j/2,:] : 0)
//                (adjoint of subsample2: the data gradient of a stride-2 1x1 conv, and the zero-stuffed output gradient that
//                 turns a stride-2 3x3 conv's backward into stride-1 convolutions)
// ------------------------------------------------------------------------------------------------
__global__ void gwd_subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int h, int w, int C8) {
  const int64_t total = static_cast<int64_t>(B) * h * w * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    const int64_t p = i / C8;
    const int j = p % w, ii = (p / w) % h, b = p / (static_cast<int64_t>(w) * h);
    y[i] = x[((static_cast<int64_t>(b) * H + 2 * ii) * W + 2 * j) * C8 + c];
  }
}

__global__ void gwd_zero_stuff2_kernel(const uint4* __restrict__ s, const uint4* __restrict__ add, uint4* __restrict__ y, int B, int H,
                                       int W, int h, int w, int C8) {
  const int64_t total = static_cast<int64_t>(B) * H * W * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    const int64_t p = i / C8;
    const int j = p % W, ii = (p / W) % H, b = p / (static_cast<int64_t>(W) * H);
    uint4 v = add ? add[i] : make_uint4(0u, 0u, 0u, 0u);
    if (!(ii & 1) && !(j & 1)) {
      const uint4 u = s[((static_cast<int64_t>(b) * h + (ii >> 1)) * w + (j >> 1)) * C8 + c];
      if (add) {
        const float2 a0 = gwd_unpack_bf16x2(v.x), a1 = gwd_unpack_bf16x2(v.y), a2 = gwd_unpack_bf16x2(v.z), a3 = gwd_unpack_bf16x2(v.w);
        const float2 b0 = gwd_unpack_bf16x2(u.x), b1 = gwd_unpack_bf16x2(u.y), b2 = gwd_unpack_bf16x2(u.z), b3 = gwd_unpack_bf16x2(u.w);
        v.x = gwd_pack_bf16x2(a0.x + b0.x, a0.y + b0.y); v.y = gwd_pack_bf16x2(a1.x + b1.x, a1.y + b1.y);
        v.z = gwd_pack_bf16x2(a2.x + b2.x, a2.y + b2.y); v.w = gwd_pack_bf16x2(a3.x + b3.x, a3.y + b3.y);
      } else {
        v = u;
      }
    }
    y[i] = v;
  }
}

// per-row scale of a flat fp32 gradient segment (FrozenBatchNorm folded into the convolution: the kernels differentiate
// w.r.t. the FOLDED filter w * s[n]; the parameter is w, so dw = dw_folded * s[n]) -- and the folded bf16 mirror
__global__ void gwd_scale_rows_kernel(float* __restrict__ g, const float* __restrict__ scale, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    g[i] *= scale[i];
}
__global__ void gwd_fold_mirror_kernel(const float* __restrict__ p, const float* __restrict__ scale, bf16* __restrict__ mirror, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    mirror[i] = __float2bfloat16(p[i] * scale[i]);
}

// ------------------------------------------------------------------------------------------------
// element-wise dropout of a [rows, C] bf16 matrix: out = res (optional) + keep(row, col) ? x / (1 - p) : 0.  The same launch
// with the same (seed, site) is the BACKWARD of itself (the mask is regenerated).  8 channels per thread.
// ------------------------------------------------------------------------------------------------
__global__ void gwd_dropout_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, uint4* __restrict__ out, int64_t n8,
                                   const uint32_t* __restrict__ seed, uint32_t site, uint32_t threshold, float inv_keep) {
  const uint32_t key = gwd_drop_key(*seed, site, 0u);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 u = x[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t r[4] = {0u, 0u, 0u, 0u};
    if (res) { const uint4 q = res[i]; r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = gwd_unpack_bf16x2(w[j]), a = gwd_unpack_bf16x2(r[j]);
      const uint32_t e = static_cast<uint32_t>(i) * 8u + 2u * j;
      const float lo = gwd_drop_keep(key, e, threshold) ? v.x * inv_keep : 0.f;
      const float hi = gwd_drop_keep(key, e + 1u, threshold) ? v.y * inv_keep : 0.f;
      o[j] = gwd_pack_bf16x2(a.x + lo, a.y + hi);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

unsigned grid_1d(int64_t n, int block) {
  int64_t g = gwd_ceil_div(n, block);
  const int64_t cap = static_cast<int64_t>(gwd_num_sms()) * 16;
  return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define GWD_STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

extern "C" int gwd_diffuse_filter_pack(const float* w_phys, const float* bias, float* fwd, float* bwd, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(w_phys && bias && fwd && bwd, "gwd_diffuse_filter_pack: null pointer");
  gwd_diffuse_filter_pack_kernel<<<(kFilt + kHeads + 255) / 256, 256, 0, stream>>>(w_phys, bias, fwd, bwd);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_affine(const float* ref, int64_t ref_rs, const float* mu, const float* logsigma, float* out, int64_t rows,
                              int32_t D, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(ref && mu && logsigma && out && rows > 0 && D > 0 && ref_rs >= D, "gwd_ref_affine: bad argument");
  gwd_ref_affine_kernel<<<static_cast<unsigned>(gwd_ceil_div(rows * D, 256)), 256, 0, stream>>>(ref, ref_rs, mu, logsigma, out, rows, D);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_affine_bwd(const float* d_kv, const float* ref, int64_t ref_rs, const float* logsigma, void* d_ref,
                                  float* dmu, float* dlogsigma, int32_t rows, int32_t D, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d_kv && ref && logsigma && d_ref && dmu && dlogsigma && rows > 0 && D > 0 && ref_rs >= D,
                "gwd_ref_affine_bwd: bad argument");
  GWD_CHECK_ARG(D % 32 == 0, "gwd_ref_affine_bwd: D must be a multiple of 32");
  gwd_ref_affine_bwd_kernel<<<static_cast<unsigned>(2 * D / 32), 256, 0, stream>>>(
      d_kv, ref, ref_rs, logsigma, static_cast<bf16*>(d_ref), dmu, dlogsigma, rows, D);
  GWD_LAUNCHED();
  return GWD_OK;
}

static int ref_bwd_smem(int R, int hd, size_t* smem) {
  *smem = (static_cast<size_t>(R) * hd + kTokTile * (R + 1) + kTokTile * (hd + 1)) * sizeof(float);
  return *smem <= 200 * 1024;
}

extern "C" int gwd_ref_requery_bwd(const float* a, const float* refv, int64_t ref_rs, const void* d_qnew, int64_t dq_rs, float* d_a,
                                   float* d_refv, int64_t drv_rs, int32_t B, int32_t T, int32_t heads, int32_t hd, int32_t R,
                                   float scale, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(a && refv && d_qnew && d_a && d_refv && a != d_a, "gwd_ref_requery_bwd: null / aliased pointer");
  GWD_CHECK_ARG(hd <= 32 && hd % 2 == 0 && R * hd <= 16 * kTokTile, "gwd_ref_requery_bwd: needs hd <= 32 and R * hd <= 2048");
  size_t smem;
  GWD_CHECK_ARG(ref_bwd_smem(R, hd, &smem), "gwd_ref_requery_bwd: %d reference points do not fit shared memory", R);
  GWD_CUDA(cudaFuncSetAttribute(gwd_ref_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  gwd_ref_bwd_kernel<true><<<dim3(heads, B, static_cast<unsigned>(gwd_ceil_div(T, kTokTile))), kTokTile, smem, stream>>>(a, refv, ref_rs, static_cast<const bf16*>(d_qnew), dq_rs, d_a,
                                                                     nullptr, 0, d_refv, drv_rs, T, heads, hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_scores_bwd(const float* d_a, const float* refk, int64_t ref_rs, const void* q, int64_t q_rs, void* d_q,
                                  int64_t dq_rs, float* d_refk, int64_t drk_rs, int32_t B, int32_t T, int32_t heads, int32_t hd,
                                  int32_t R, float scale, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d_a && refk && q && d_q && d_refk, "gwd_ref_scores_bwd: null pointer");
  GWD_CHECK_ARG(hd <= 32 && hd % 2 == 0 && dq_rs % 2 == 0 && R * hd <= 16 * kTokTile,
                "gwd_ref_scores_bwd: needs hd <= 32 and R * hd <= 2048");
  size_t smem;
  GWD_CHECK_ARG(ref_bwd_smem(R, hd, &smem), "gwd_ref_scores_bwd: %d reference points do not fit shared memory", R);
  GWD_CUDA(cudaFuncSetAttribute(gwd_ref_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  gwd_ref_bwd_kernel<false><<<dim3(heads, B, static_cast<unsigned>(gwd_ceil_div(T, kTokTile))), kTokTile, smem, stream>>>(d_a, refk, ref_rs, static_cast<const bf16*>(q), q_rs, nullptr,
                                                                      static_cast<bf16*>(d_q), dq_rs, d_refk, drk_rs, T, heads, hd, R,
                                                                      scale);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_diffuse_bwd(const float* g, const float* raw, const double* stats, const float* a_in, const float* filt_bwd,
                                   float* d_raw_ws, double* stats2_ws, float* d_a_in, float* dw_phys, float* db, int32_t B,
                                   int32_t heads, int32_t P, int32_t R, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(g && raw && stats && a_in && filt_bwd && d_raw_ws && stats2_ws && d_a_in && dw_phys && db && d_a_in != g,
                "gwd_ref_diffuse_bwd: null / aliased pointer");
  GWD_CHECK_ARG(heads == kHeads, "gwd_ref_diffuse_bwd: built for %d heads (got %d)", kHeads, heads);
  const int per_img = P * R, planes = B * heads;
  // stats2_ws: fp64 [4 * B * heads]: (sum dy, sum dy y) per plane, then scratch for the convolution launcher
  GWD_CUDA(cudaMemsetAsync(stats2_ws, 0, sizeof(double) * 2 * planes, stream));
  int chunks = static_cast<int>(gwd_ceil_div(per_img, 256 * 8));
  if (chunks < 1) chunks = 1;
  gwd_diffuse_bwd_reduce_kernel<<<dim3(chunks, planes), 256, 0, stream>>>(g, raw, stats, stats2_ws, per_img);
  GWD_LAUNCHED();
  gwd_diffuse_bwd_apply_kernel<<<dim3(chunks, planes), 256, 0, stream>>>(g, raw, stats, stats2_ws, d_raw_ws, per_img);
  GWD_LAUNCHED();
  const size_t smem = (static_cast<size_t>(kHeads) * (kBand + 2) * (R + 2) + static_cast<size_t>(kHeads) * kBand * R) * sizeof(float);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_ref_diffuse_bwd: %d reference points do not fit shared memory", R);
  GWD_CUDA(cudaFuncSetAttribute(gwd_diffuse_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int items = static_cast<int>(gwd_ceil_div(P, kBand)) * B;
  const int ctas = std::min(items, 2 * gwd_num_sms());
  gwd_diffuse_wgrad_kernel<<<ctas, 256, smem, stream>>>(d_raw_ws, a_in, dw_phys, db, B, P, R);
  GWD_LAUNCHED();
  // d a_in = g + conv(d raw, adjoint filter)
  return gwd_ref_diffuse_conv_dev(d_raw_ws, filt_bwd, g, d_a_in, stats2_ws + 2 * planes, B, heads, P, R, stream_);
}

extern "C" int gwd_subsample2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && y && C % 8 == 0 && B > 0 && H > 0 && W > 0, "gwd_subsample2: bad argument (C must be a multiple of 8)");
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  gwd_subsample2_kernel<<<grid_1d(static_cast<int64_t>(B) * h * w * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), B, H, W, h, w, C / 8);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_zero_stuff2(const void* s, const void* add, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(s && y && C % 8 == 0 && B > 0 && H > 0 && W > 0, "gwd_zero_stuff2: bad argument (C must be a multiple of 8)");
  const int h = (H + 1) / 2, w = (W + 1) / 2;
  gwd_zero_stuff2_kernel<<<grid_1d(static_cast<int64_t>(B) * H * W * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const uint4*>(s), static_cast<const uint4*>(add), static_cast<uint4*>(y), B, H, W, h, w, C / 8);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_scale_rows(float* g, const float* scale, int64_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(g && scale && n > 0, "gwd_scale_rows: bad argument");
  gwd_scale_rows_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(g, scale, n);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_fold_mirror(const float* p, const float* scale, void* mirror, int64_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(p && scale && mirror && n > 0, "gwd_fold_mirror: bad argument");
  gwd_fold_mirror_kernel<<<grid_1d(n, 256), 256, 0, stream>>>(p, scale, static_cast<bf16*>(mirror), n);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_dropout(const void* x, const void* res, void* out, int64_t n, const uint32_t* seed, uint32_t site, float p,
                           void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && seed && n > 0 && n % 8 == 0 && n < (1ll << 32), "gwd_dropout: null pointer / n must be a multiple of 8 below 2^32");
  GWD_CHECK_ARG(p >= 0.f && p < 1.f, "gwd_dropout: p must be in [0, 1)");
  const uint32_t threshold = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  gwd_dropout_kernel<<<grid_1d(n / 8, 256), 256, 0, stream>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(res),
                                                              static_cast<uint4*>(out), n / 8, seed, site, threshold, 1.f / (1.f - p));
  GWD_LAUNCHED();
  return GWD_OK;
}
