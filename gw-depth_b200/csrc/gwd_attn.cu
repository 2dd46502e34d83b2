// gwd_attn.cu -- fused softmax attention kernels (CUDA cores; head dims 4..32 are below any MMA tile
// for the window variants, and the DETR attention cores are 0.2% of the model FLOPs).
//
//   gwd_attention        generic multi-head softmax(QK^T*scale + bias + mask) V, K/V staged in shared memory
//                        replaces src/models/multi_head_attention.py:317-372 (DETR self / cross attention) and
//                        the window attention cores of src/models/multiscale_transformerr.py:311-328,539-556
//   gwd_token_attention  per-window class-token CHANNEL attention, multiscale_transformerr.py:561-578
//   gwd_ref_scores / gwd_ref_diffuse / gwd_ref_requery
//                        the line end-point ("glass structure") re-query of WindowAttention, :281-310
#include <algorithm>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

struct AttnParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  int items, heads, Lq, Lk, hd;
  int64_t q_is, q_rs, k_is, k_rs, v_is, v_rs, o_is, o_rs;
  const float* bias;        // [heads, Lq, Lk] or null
  const float* mask;        // [nW, Lq, Lk] or null (window = item % nW)
  int nW;
  const uint8_t* kpm;       // [items, Lk], 1 = key is padding, or null
  float scale;
  int q_tile;
};

constexpr int kAttnWarps = 4;

__global__ void __launch_bounds__(kAttnWarps * 32) gwd_attention_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int hd = p.hd, Lk = p.Lk;
  const int kstride = hd + 2;  // bf16 elements per padded K row (odd number of 32-bit words -> no bank conflicts)
  bf16* Ks = reinterpret_cast<bf16*>(smem);
  bf16* Vs = Ks + static_cast<size_t>(Lk) * kstride;
  float* Ps = reinterpret_cast<float*>(Vs + static_cast<size_t>(Lk) * hd + 8);
  Ps = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Ps) + 15) & ~uintptr_t(15));
  float* Qs = Ps + static_cast<size_t>(kAttnWarps) * Lk;

  const int item = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * p.q_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage K and V of this (item, head): bf16x2 granularity (hd is even)
  const bf16* kbase = p.k + item * p.k_is + head * hd;
  const bf16* vbase = p.v + item * p.v_is + head * hd;
  const int hw = hd >> 1;
  for (int idx = threadIdx.x; idx < Lk * hw; idx += blockDim.x) {
    int j = idx / hw, w = idx - j * hw;
    reinterpret_cast<uint32_t*>(Ks + static_cast<size_t>(j) * kstride)[w] =
        reinterpret_cast<const uint32_t*>(kbase + j * p.k_rs)[w];
    reinterpret_cast<uint32_t*>(Vs + static_cast<size_t>(j) * hd)[w] =
        reinterpret_cast<const uint32_t*>(vbase + j * p.v_rs)[w];
  }
  __syncthreads();

  float* ps = Ps + static_cast<size_t>(warp) * Lk;
  float* qs = Qs + warp * 32;
  const int G = 32 / hd;           // key groups in the PV pass
  const int d = lane % hd, g = lane / hd;
  const int q_end = min(q0 + p.q_tile, p.Lq);
  for (int qi = q0 + warp; qi < q_end; qi += kAttnWarps) {
    const bf16* qrow = p.q + item * p.q_is + static_cast<int64_t>(qi) * p.q_rs + head * hd;
    if (lane < hd) qs[lane] = __bfloat162float(qrow[lane]) * p.scale;
    __syncwarp();
    const float* brow = p.bias ? p.bias + (static_cast<int64_t>(head) * p.Lq + qi) * Lk : nullptr;
    const float* mrow = p.mask ? p.mask + (static_cast<int64_t>(item % p.nW) * p.Lq + qi) * Lk : nullptr;
    const uint8_t* kp = p.kpm ? p.kpm + static_cast<int64_t>(item) * Lk : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) {
      const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + static_cast<size_t>(j) * kstride);
      float s = 0.f;
#pragma unroll 4
      for (int w = 0; w < hw; ++w) {
        float2 kk = gwd_unpack_bf16x2(kr[w]);
        s = fmaf(qs[2 * w], kk.x, s);
        s = fmaf(qs[2 * w + 1], kk.y, s);
      }
      if (brow) s += brow[j];
      if (mrow) s += mrow[j];
      if (kp && kp[j]) s = -INFINITY;
      ps[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = gwd_warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      float e = __expf(ps[j] - mx);
      ps[j] = e;
      sum += e;
    }
    sum = gwd_warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int j = g; j < Lk; j += G) acc = fmaf(ps[j], __bfloat162float(Vs[static_cast<size_t>(j) * hd + d]), acc);
    for (int o = hd; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane < hd) {
      bf16* orow = p.o + item * p.o_is + static_cast<int64_t>(qi) * p.o_rs + head * hd;
      orow[lane] = __float2bfloat16(acc / sum);
    }
    __syncwarp();
  }
}

// -------------------------------------------------------------------------------------------------
// class-token channel attention (one CTA per window, one warp per head)
//   tq  [items, N, tdim]        projected (cls_dth_q / cls_seg_q) class tokens, two of them (depth, seg)
//   tk, tv [items, N, tC]       global_k / global_v of cat[x, depth_tok, seg_tok]
//   out[n][h*td + i] = sum_c softmax_c( scale * sum_n' tq[n'][h*td+i] tk[n'][h*tc+c] ) * tv[n][h*tc+c]
// -------------------------------------------------------------------------------------------------
struct TokAttnParams {
  const bf16* dq; const bf16* sq; const bf16* tk; const bf16* tv;
  bf16* dout; bf16* sout;
  int items, N, heads, td, tc;   // td = tdim/heads (4), tc = tC/heads
  int64_t q_rs, k_rs, v_rs, o_rs;  // row strides (elements); item stride = N * row stride
  float scale;
};

__global__ void gwd_token_attention_kernel(const TokAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  // per warp: attn[2][td][tc] floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  const int h = warp;
  if (h >= p.heads) return;
  float* A = reinterpret_cast<float*>(smem) + static_cast<size_t>(warp) * 2 * p.td * p.tc;
  const int N = p.N, td = p.td, tc = p.tc;
  const bf16* tk = p.tk + static_cast<int64_t>(item) * N * p.k_rs + h * tc;
  const bf16* tv = p.tv + static_cast<int64_t>(item) * N * p.v_rs + h * tc;
  // scores: (which, i, c) pairs distributed over lanes
  const int npairs = 2 * td * tc;
  for (int e = lane; e < npairs; e += 32) {
    int which = e / (td * tc);
    int r = e - which * td * tc;
    int i = r / tc, c = r - i * tc;
    const bf16* tq = (which == 0 ? p.dq : p.sq) + static_cast<int64_t>(item) * N * p.q_rs + h * td + i;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(__bfloat162float(tq[n * p.q_rs]), __bfloat162float(tk[n * p.k_rs + c]), s);
    A[e] = s * p.scale;
  }
  __syncwarp();
  // softmax over c for each (which, i)
  for (int r = lane; r < 2 * td; r += 32) {
    float* a = A + r * tc;
    float mx = -INFINITY;
    for (int c = 0; c < tc; ++c) mx = fmaxf(mx, a[c]);
    float sum = 0.f;
    for (int c = 0; c < tc; ++c) { a[c] = __expf(a[c] - mx); sum += a[c]; }
    float inv = 1.f / sum;
    for (int c = 0; c < tc; ++c) a[c] *= inv;
  }
  __syncwarp();
  // out[which][n][h*td+i] = sum_c A[which][i][c] * tv[n][c]
  for (int e = lane; e < 2 * N * td; e += 32) {
    int which = e / (N * td);
    int r = e - which * N * td;
    int n = r / td, i = r - n * td;
    const float* a = A + (which * td + i) * tc;
    float s = 0.f;
    for (int c = 0; c < tc; ++c) s = fmaf(a[c], __bfloat162float(tv[n * p.v_rs + c]), s);
    bf16* o = (which == 0 ? p.dout : p.sout) + (static_cast<int64_t>(item) * N + n) * p.o_rs + h * td + i;
    *o = __float2bfloat16(s);
  }
}

// -------------------------------------------------------------------------------------------------
// line end-point re-query (1/32 scale)
// -------------------------------------------------------------------------------------------------
// scores[b][h][w*N+n][r] = scale * q[(b*nW+w)*N+n][h*hd:] . refk[b*R+r][h*hd:]     (fp32 out)
__global__ void gwd_ref_scores_kernel(const bf16* __restrict__ q, int64_t q_rs, const float* __restrict__ refk,
                                      int64_t ref_rs, float* __restrict__ out, int B, int nW, int N, int heads, int hd,
                                      int R, float scale) {
  int64_t total = static_cast<int64_t>(B) * heads * nW * N * R;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int r = idx % R;
    int64_t t = idx / R;
    int tok = t % (nW * N);
    t /= (nW * N);
    int h = t % heads;
    int b = t / heads;
    const bf16* qr = q + (static_cast<int64_t>(b) * nW * N + tok) * q_rs + h * hd;
    const float* kr = refk + (static_cast<int64_t>(b) * R + r) * ref_rs + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(__bfloat162float(qr[d]) * scale, kr[d], s);
    out[idx] = s;
  }
}

// one diffusion step: a += gelu(layer_norm_over_image(conv3x3(a)))   a: [B][heads][P][R] fp32
// grid (heads_out, B); the conv output of one (b, oc) image is kept in shared memory (P*R floats)
__global__ void __launch_bounds__(256) gwd_ref_diffuse_kernel(const float* __restrict__ a_in, float* __restrict__ a_out,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              int heads, int P, int R) {
  extern __shared__ float img[];
  __shared__ float red[64];
  const int oc = blockIdx.x, b = blockIdx.y;
  const float* in_b = a_in + static_cast<int64_t>(b) * heads * P * R;
  const int n = P * R;
  float s = 0.f, ss = 0.f;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    int y = idx / R, x = idx - y * R;
    float acc = bias[oc];
    for (int ic = 0; ic < heads; ++ic) {
      const float* src = in_b + static_cast<int64_t>(ic) * n;
      const float* wk = w + (static_cast<int64_t>(oc) * heads + ic) * 9;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= P) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          int xx = x + dx;
          if (xx < 0 || xx >= R) continue;
          acc = fmaf(wk[(dy + 1) * 3 + dx + 1], src[yy * R + xx], acc);
        }
      }
    }
    img[idx] = acc;
    s += acc;
    ss += acc * acc;
  }
  s = gwd_warp_sum(s);
  ss = gwd_warp_sum(ss);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp] = s; red[32 + warp] = ss; }
  __syncthreads();
  if (warp == 0) {
    float a = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    float c = lane < (blockDim.x >> 5) ? red[32 + lane] : 0.f;
    a = gwd_warp_sum(a);
    c = gwd_warp_sum(c);
    if (lane == 0) { red[0] = a; red[32] = c; }
  }
  __syncthreads();
  float mean = red[0] / n;
  float var = fmaxf(red[32] / n - mean * mean, 0.f);
  float rstd = rsqrtf(var + 1e-5f);
  const float* res = in_b + static_cast<int64_t>(oc) * n;
  float* dst = a_out + (static_cast<int64_t>(b) * heads + oc) * n;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    float v = (img[idx] - mean) * rstd;
    dst[idx] = res[idx] + gwd_apply_act(v, GWD_ACT_GELU);
  }
}

// q_new[(b*nW+w)*N+n][h*hd+d] = scale * sum_r softmax_r(a[b][h][w*N+n][:])[r] * refv[b*R+r][h*hd+d]   (bf16 out)
// one warp per (b, h, token)
__global__ void gwd_ref_requery_kernel(const float* __restrict__ a, const float* __restrict__ refv, int64_t ref_rs,
                                       bf16* __restrict__ out, int64_t o_rs, int B, int nW, int N, int heads, int hd, int R,
                                       float scale) {
  int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t total = static_cast<int64_t>(B) * heads * nW * N;
  if (wid >= total) return;
  int tok = wid % (nW * N);
  int64_t t = wid / (nW * N);
  int h = t % heads;
  int b = t / heads;
  const float* row = a + wid * R;
  float mx = -INFINITY;
  for (int r = lane; r < R; r += 32) mx = fmaxf(mx, row[r]);
  mx = gwd_warp_max(mx);
  float sum = 0.f;
  for (int r = lane; r < R; r += 32) sum += __expf(row[r] - mx);
  sum = gwd_warp_sum(sum);
  // lane <-> d (hd <= 32)
  if (lane < hd) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r)
      acc = fmaf(__expf(row[r] - mx), refv[(static_cast<int64_t>(b) * R + r) * ref_rs + h * hd + lane], acc);
    out[(static_cast<int64_t>(b) * nW * N + tok) * o_rs + h * hd + lane] = __float2bfloat16(acc / sum * scale);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" int gwd_attention(const gwd_attn_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(d && d->q && d->k && d->v && d->o, "gwd_attention: null pointer");
  GWD_CHECK_ARG(d->hd == 4 || d->hd == 8 || d->hd == 16 || d->hd == 32, "gwd_attention: head dim %d unsupported", d->hd);
  GWD_CHECK_ARG(d->items > 0 && d->heads > 0 && d->Lq > 0 && d->Lk > 0, "gwd_attention: empty problem");
  GWD_CHECK_ARG(d->k_row_stride % 2 == 0 && d->v_row_stride % 2 == 0 && d->k_item_stride % 2 == 0 && d->v_item_stride % 2 == 0 &&
                    (reinterpret_cast<uintptr_t>(d->k) & 3) == 0 && (reinterpret_cast<uintptr_t>(d->v) & 3) == 0,
                "gwd_attention: K/V must be 4-byte aligned with even strides");
  AttnParams p;
  p.q = static_cast<const bf16*>(d->q); p.k = static_cast<const bf16*>(d->k); p.v = static_cast<const bf16*>(d->v);
  p.o = static_cast<bf16*>(d->o);
  p.items = d->items; p.heads = d->heads; p.Lq = d->Lq; p.Lk = d->Lk; p.hd = d->hd;
  p.q_is = d->q_item_stride; p.q_rs = d->q_row_stride; p.k_is = d->k_item_stride; p.k_rs = d->k_row_stride;
  p.v_is = d->v_item_stride; p.v_rs = d->v_row_stride; p.o_is = d->o_item_stride; p.o_rs = d->o_row_stride;
  p.bias = d->bias; p.mask = d->mask; p.nW = d->mask_windows > 0 ? d->mask_windows : 1;
  p.kpm = d->key_padding; p.scale = d->scale;
  p.q_tile = d->Lq <= 64 ? d->Lq : 32;
  size_t smem = static_cast<size_t>(d->Lk) * (d->hd + 2) * 2 + static_cast<size_t>(d->Lk) * d->hd * 2 + 16 + 16 +
                static_cast<size_t>(kAttnWarps) * d->Lk * 4 + kAttnWarps * 32 * 4;
  GWD_CHECK_ARG(smem <= 220 * 1024, "gwd_attention: Lk=%d does not fit shared memory", d->Lk);
  if (smem > 48 * 1024) {
    static size_t configured = 0;
    if (smem > configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured = 227 * 1024;
    }
  }
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(d->Lq, p.q_tile)), d->heads, d->items);
  gwd_attention_kernel<<<grid, kAttnWarps * 32, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_token_attention(const void* dq, const void* sq, const void* tk, const void* tv, void* dout, void* sout,
                                   int32_t items, int32_t N, int32_t heads, int32_t td, int32_t tc, int64_t q_rs,
                                   int64_t k_rs, int64_t v_rs, int64_t o_rs, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(dq && sq && tk && tv && dout && sout, "gwd_token_attention: null pointer");
  GWD_CHECK_ARG(heads > 0 && heads <= 32 && td > 0 && tc > 0 && items > 0, "gwd_token_attention: bad shape");
  TokAttnParams p;
  p.dq = static_cast<const bf16*>(dq); p.sq = static_cast<const bf16*>(sq);
  p.tk = static_cast<const bf16*>(tk); p.tv = static_cast<const bf16*>(tv);
  p.dout = static_cast<bf16*>(dout); p.sout = static_cast<bf16*>(sout);
  p.items = items; p.N = N; p.heads = heads; p.td = td; p.tc = tc;
  p.q_rs = q_rs; p.k_rs = k_rs; p.v_rs = v_rs; p.o_rs = o_rs; p.scale = scale;
  size_t smem = static_cast<size_t>(heads) * 2 * td * tc * sizeof(float);
  gwd_token_attention_kernel<<<items, heads * 32, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_scores(const void* q, int64_t q_rs, const float* refk, int64_t ref_rs, float* out, int32_t B,
                              int32_t nW, int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(q && refk && out, "gwd_ref_scores: null pointer");
  int64_t total = static_cast<int64_t>(B) * heads * nW * N * R;
  int blocks = static_cast<int>(std::min<int64_t>(gwd_ceil_div(total, 256), gwd_num_sms() * 8));
  gwd_ref_scores_kernel<<<blocks, 256, 0, stream>>>(static_cast<const bf16*>(q), q_rs, refk, ref_rs, out, B, nW, N, heads,
                                                    hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_diffuse(const float* a_in, float* a_out, const float* w, const float* bias, int32_t B, int32_t heads,
                               int32_t P, int32_t R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a_in && a_out && w && bias && a_in != a_out, "gwd_ref_diffuse: null / aliased pointer");
  size_t smem = static_cast<size_t>(P) * R * sizeof(float);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_ref_diffuse: image %dx%d too large", P, R);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_ref_diffuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  gwd_ref_diffuse_kernel<<<dim3(heads, B), 256, smem, stream>>>(a_in, a_out, w, bias, heads, P, R);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_requery(const float* a, const float* refv, int64_t ref_rs, void* out, int64_t o_rs, int32_t B,
                               int32_t nW, int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a && refv && out && hd <= 32, "gwd_ref_requery: bad argument");
  int64_t warps = static_cast<int64_t>(B) * heads * nW * N;
  int blocks = static_cast<int>(gwd_ceil_div(warps * 32, 256));
  gwd_ref_requery_kernel<<<blocks, 256, 0, stream>>>(a, refv, ref_rs, static_cast<bf16*>(out), o_rs, B, nW, N, heads, hd, R,
                                                     scale);
  GWD_LAUNCHED();
  return GWD_OK;
}
