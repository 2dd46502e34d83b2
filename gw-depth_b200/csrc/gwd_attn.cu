// gwd_attn.cu -- fused softmax attention kernels (CUDA cores; head dims 4..32 are below any MMA tile
// for the window variants, and the DETR attention cores are 0.2% of the model FLOPs).
//
//   gwd_attention        generic multi-head softmax(QK^T*scale + bias + mask) V, K/V staged in shared memory
//                        replaces src/models/multi_head_attention.py:317-372 (DETR self / cross attention) and
//                        the window attention cores of src/models/multiscale_transformerr.py:311-328,539-556
//   gwd_token_attention  per-window class-token CHANNEL attention, multiscale_transformerr.py:561-578
//   gwd_ref_scores / gwd_ref_diffuse / gwd_ref_requery
//                        the line end-point ("glass structure") re-query of WindowAttention, :281-310
#include <stdlib.h>
#include <string.h>
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


// -------------------------------------------------------------------------------------------------
// thread-per-query attention: one CTA per (item, head, query tile); K and V of the (item, head) are staged in shared
// memory as fp32 and every lane owns ONE query: it walks the keys with an online softmax, reading K/V rows as
// warp-wide broadcasts (no shuffles, no per-query synchronisation).  HD is a template parameter so q / acc stay in
// registers.
// -------------------------------------------------------------------------------------------------
// smem element type: fp32 (fast path) or bf16 (long key axes, e.g. L = 1200 at 960x1280, that would not fit as fp32)
template <typename T> struct KvRow;
template <> struct KvRow<float> {
  static __device__ __forceinline__ float4 load4(const float* row, int w) { return reinterpret_cast<const float4*>(row)[w]; }
  static __device__ __forceinline__ void store2(float* row, int w, float2 v) { row[2 * w] = v.x; row[2 * w + 1] = v.y; }
};
template <> struct KvRow<bf16> {
  static __device__ __forceinline__ float4 load4(const bf16* row, int w) {
    uint2 u = reinterpret_cast<const uint2*>(row)[w];
    float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store2(bf16* row, int w, float2 v) {
    reinterpret_cast<uint32_t*>(row)[w] = gwd_pack_bf16x2(v.x, v.y);
  }
};

template <int HD, typename T>
__global__ void __launch_bounds__(128) gwd_attention_tq_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  T* smf = reinterpret_cast<T*>(smraw);
  const int Lk = p.Lk;
  T* Ks = smf;                                  // [Lk][HD]
  T* Vs = smf + static_cast<size_t>(Lk) * HD;   // [Lk][HD]
  const int item = blockIdx.z, head = blockIdx.y;
  const bf16* kbase = p.k + item * p.k_is + head * HD;
  const bf16* vbase = p.v + item * p.v_is + head * HD;
  constexpr int HW = HD / 2;
  for (int idx = threadIdx.x; idx < Lk * HW; idx += blockDim.x) {
    int j = idx / HW, w = idx - j * HW;
    float2 kk = gwd_unpack_bf16x2(reinterpret_cast<const uint32_t*>(kbase + j * p.k_rs)[w]);
    float2 vv = gwd_unpack_bf16x2(reinterpret_cast<const uint32_t*>(vbase + j * p.v_rs)[w]);
    KvRow<T>::store2(Ks + j * HD, w, kk);
    KvRow<T>::store2(Vs + j * HD, w, vv);
  }
  __syncthreads();
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= p.Lq) return;
  float q[HD], acc[HD];
  const bf16* qrow = p.q + item * p.q_is + static_cast<int64_t>(qi) * p.q_rs + head * HD;
#pragma unroll
  for (int w = 0; w < HW; ++w) {
    float2 t = gwd_unpack_bf16x2(reinterpret_cast<const uint32_t*>(qrow)[w]);
    q[2 * w] = t.x * p.scale; q[2 * w + 1] = t.y * p.scale;
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  const float* brow = p.bias ? p.bias + (static_cast<int64_t>(head) * p.Lq + qi) * Lk : nullptr;
  const float* mrow = p.mask ? p.mask + (static_cast<int64_t>(item % p.nW) * p.Lq + qi) * Lk : nullptr;
  const uint8_t* kp = p.kpm ? p.kpm + static_cast<int64_t>(item) * Lk : nullptr;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < Lk; ++j) {
    const T* kr = Ks + j * HD;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < HD / 4; ++w) {
      float4 k4 = KvRow<T>::load4(kr, w);
      s = fmaf(q[4 * w], k4.x, s); s = fmaf(q[4 * w + 1], k4.y, s);
      s = fmaf(q[4 * w + 2], k4.z, s); s = fmaf(q[4 * w + 3], k4.w, s);
    }
    if (brow) s += __ldg(brow + j);
    if (mrow) s += __ldg(mrow + j);
    if (kp && kp[j]) continue;
    float mn = fmaxf(m, s);
    float corr = __expf(m - mn), pj = __expf(s - mn);
    m = mn;
    l = l * corr + pj;
    const T* vr = Vs + j * HD;
#pragma unroll
    for (int w = 0; w < HD / 4; ++w) {
      float4 v4 = KvRow<T>::load4(vr, w);
      acc[4 * w] = fmaf(acc[4 * w], corr, pj * v4.x); acc[4 * w + 1] = fmaf(acc[4 * w + 1], corr, pj * v4.y);
      acc[4 * w + 2] = fmaf(acc[4 * w + 2], corr, pj * v4.z); acc[4 * w + 3] = fmaf(acc[4 * w + 3], corr, pj * v4.w);
    }
  }
  float inv = 1.f / l;
  bf16* orow = p.o + item * p.o_is + static_cast<int64_t>(qi) * p.o_rs + head * HD;
#pragma unroll
  for (int w = 0; w < HW; ++w)
    reinterpret_cast<uint32_t*>(orow)[w] = gwd_pack_bf16x2(acc[2 * w] * inv, acc[2 * w + 1] * inv);
}

template <int HD, typename T>
static int launch_attention_tq_as(const AttnParams& p, cudaStream_t stream) {
  size_t smem = static_cast<size_t>(p.Lk) * HD * 2 * sizeof(T);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_attention: Lk=%d does not fit shared memory", p.Lk);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_attention_tq_kernel<HD, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  int threads = p.Lq <= 64 ? 64 : 128;
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(p.Lq, threads)), p.heads, p.items);
  gwd_attention_tq_kernel<HD, T><<<grid, threads, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}
template <int HD>
static int launch_attention_tq(const AttnParams& p, cudaStream_t stream) {
  if (static_cast<size_t>(p.Lk) * HD * 2 * sizeof(float) <= 160 * 1024) return launch_attention_tq_as<HD, float>(p, stream);
  return launch_attention_tq_as<HD, bf16>(p, stream);
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
  const int N = p.N, td = p.td, tc = p.tc, heads = p.heads;
  const int TQ = heads * td, TC = heads * tc;          // row widths of the query / key-value tensors
  bf16* sk = reinterpret_cast<bf16*>(smem);            // [N][TC]
  bf16* sv = sk + static_cast<size_t>(N) * TC;         // [N][TC]
  bf16* sdq = sv + static_cast<size_t>(N) * TC;        // [N][TQ]
  bf16* ssq = sdq + static_cast<size_t>(N) * TQ;       // [N][TQ]
  float* Aall = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ssq + static_cast<size_t>(N) * TQ) + 15) & ~uintptr_t(15));
  const int item = blockIdx.x;
  {  // coalesced staging of the window's rows (32-bit words)
    const int kw = TC >> 1, qw = TQ >> 1;
    const bf16* gk = p.tk + static_cast<int64_t>(item) * N * p.k_rs;
    const bf16* gv = p.tv + static_cast<int64_t>(item) * N * p.v_rs;
    for (int i = threadIdx.x; i < N * kw; i += blockDim.x) {
      int n = i / kw, w = i - n * kw;
      reinterpret_cast<uint32_t*>(sk)[i] = reinterpret_cast<const uint32_t*>(gk + n * p.k_rs)[w];
      reinterpret_cast<uint32_t*>(sv)[i] = reinterpret_cast<const uint32_t*>(gv + n * p.v_rs)[w];
    }
    const bf16* gd = p.dq + static_cast<int64_t>(item) * N * p.q_rs;
    const bf16* gs = p.sq + static_cast<int64_t>(item) * N * p.q_rs;
    for (int i = threadIdx.x; i < N * qw; i += blockDim.x) {
      int n = i / qw, w = i - n * qw;
      reinterpret_cast<uint32_t*>(sdq)[i] = reinterpret_cast<const uint32_t*>(gd + n * p.q_rs)[w];
      reinterpret_cast<uint32_t*>(ssq)[i] = reinterpret_cast<const uint32_t*>(gs + n * p.q_rs)[w];
    }
  }
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= heads) return;
  float* A = Aall + static_cast<size_t>(h) * 2 * td * tc;
  // scores A[which][i][c] = scale * sum_n tq[n][h*td+i] * tk[n][h*tc+c]
  const int npairs = 2 * td * tc;
  for (int e = lane; e < npairs; e += 32) {
    int which = e / (td * tc);
    int r = e - which * td * tc;
    int i = r / tc, c = r - i * tc;
    const bf16* tq = (which == 0 ? sdq : ssq) + h * td + i;
    const bf16* tk = sk + h * tc + c;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(__bfloat162float(tq[n * TQ]), __bfloat162float(tk[n * TC]), s);
    A[e] = s * p.scale;
  }
  __syncwarp();
  for (int r = lane; r < 2 * td; r += 32) {   // softmax over c
    float* a = A + r * tc;
    float mx = -INFINITY;
    for (int c = 0; c < tc; ++c) mx = fmaxf(mx, a[c]);
    float sum = 0.f;
    for (int c = 0; c < tc; ++c) { a[c] = __expf(a[c] - mx); sum += a[c]; }
    float inv = 1.f / sum;
    for (int c = 0; c < tc; ++c) a[c] *= inv;
  }
  __syncwarp();
  // out[which][n][h*td+i] = sum_c A[which][i][c] * tv[n][h*tc+c]
  for (int e = lane; e < 2 * N * td; e += 32) {
    int which = e / (N * td);
    int r = e - which * N * td;
    int n = r / td, i = r - n * td;
    const float* a = A + (which * td + i) * tc;
    const bf16* tv = sv + static_cast<size_t>(n) * TC + h * tc;
    float s = 0.f;
    for (int c = 0; c < tc; ++c) s = fmaf(a[c], __bfloat162float(tv[c]), s);
    bf16* o = (which == 0 ? p.dout : p.sout) + (static_cast<int64_t>(item) * N + n) * p.o_rs + h * td + i;
    *o = __float2bfloat16(s);
  }
}

// -------------------------------------------------------------------------------------------------
// line end-point re-query (1/32 scale)
// -------------------------------------------------------------------------------------------------
// scores[b][h][tok][r] = scale * q[b*T+tok][h*hd:] . refk[b*R+r][h*hd:]     (fp32 out, T = nW*N tokens per image)
// grid (token tiles, heads, B): ref_k of one (image, head) sits in shared memory, one thread per token
__global__ void __launch_bounds__(128) gwd_ref_scores_kernel(const bf16* __restrict__ q, int64_t q_rs,
                                                            const float* __restrict__ refk, int64_t ref_rs,
                                                            float* __restrict__ out, int T, int heads, int hd, int R,
                                                            float scale) {
  extern __shared__ __align__(16) float rk[];  // [R][hd] reference keys, then [128][R + 1] staged scores
  float* stile = rk + R * hd;
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    int r = i / hd, dd = i - r * hd;
    rk[i] = refk[(static_cast<int64_t>(b) * R + r) * ref_rs + h * hd + dd] * scale;
  }
  __syncthreads();
  const int tok0 = blockIdx.x * blockDim.x, tok = tok0 + threadIdx.x;
  if (tok < T) {
    float qv[32];
    const bf16* qr = q + (static_cast<int64_t>(b) * T + tok) * q_rs + h * hd;
    if ((hd & 7) == 0 && (reinterpret_cast<uintptr_t>(qr) & 15) == 0) {   // 16-byte pieces of the query row
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (8 * w < hd) u = __ldg(reinterpret_cast<const uint4*>(qr) + w);
        float2 f0 = gwd_unpack_bf16x2(u.x), f1 = gwd_unpack_bf16x2(u.y), f2 = gwd_unpack_bf16x2(u.z), f3 = gwd_unpack_bf16x2(u.w);
        qv[8 * w + 0] = f0.x; qv[8 * w + 1] = f0.y; qv[8 * w + 2] = f1.x; qv[8 * w + 3] = f1.y;
        qv[8 * w + 4] = f2.x; qv[8 * w + 5] = f2.y; qv[8 * w + 6] = f3.x; qv[8 * w + 7] = f3.y;
      }
    } else {
#pragma unroll
      for (int dd = 0; dd < 32; ++dd) qv[dd] = dd < hd ? __bfloat162float(qr[dd]) : 0.f;
    }
    float* srow = stile + threadIdx.x * (R + 1);
    for (int r = 0; r < R; ++r) {
      float sacc = 0.f;
#pragma unroll
      for (int dd = 0; dd < 32; ++dd)
        if (dd < hd) sacc = fmaf(qv[dd], rk[r * hd + dd], sacc);
      srow[r] = sacc;
    }
  }
  __syncthreads();
  // the tile's scores are one contiguous block of the output: consecutive threads write consecutive floats
  const int ntok = min(static_cast<int>(blockDim.x), T - tok0);
  float* o = out + ((static_cast<int64_t>(b) * heads + h) * T + tok0) * R;
  for (int i = threadIdx.x; i < ntok * R; i += blockDim.x) {
    const int tl = i / R, r = i - tl * R;
    o[i] = stile[tl * (R + 1) + r];
  }
}

constexpr int kDiffBand = 7;
constexpr int kDiffHeads = 16;
struct DiffuseFilter {
  float w[kDiffHeads * kDiffHeads * 9];
  float b[kDiffHeads];
};
// Direct convolution, one pixel x 4 output channels per thread: lanes map to consecutive pixels (stride-1, conflict-free
// shared-memory reads of the input window), the 9 taps of the thread's 4 output channels come as float4 broadcasts
// from a [ic][tap][oc] copy of the filter.  8 warps: warp w -> channel group w%4, pixel half w/4.
// grid (row bands of kDiffBand rows, B), block 256.
__global__ void __launch_bounds__(256) gwd_ref_diffuse_conv_kernel(const float* __restrict__ a, float* __restrict__ raw,
                                                                   const __grid_constant__ DiffuseFilter flt,
                                                                   double* __restrict__ stats, int P, int R,
                                                                   const float* __restrict__ fdev, const float* __restrict__ add) {
  extern __shared__ __align__(16) float dsm[];
  // training: the filter is a parameter that the optimizer updates on the device -> read it from `fdev`
  // ([oc][ic][ky][kx] then the 16 biases) instead of the by-value copy of the inference plan
  const float* fw = fdev ? fdev : flt.w;
  const float* fb = fdev ? fdev + kDiffHeads * kDiffHeads * 9 : flt.b;
  constexpr int heads = kDiffHeads;
  const int y0 = blockIdx.x * kDiffBand, b = blockIdx.y;
  const int rows = min(kDiffBand, P - y0);
  const int TR = kDiffBand + 2, TC = R + 2;
  float* wsm = dsm;                                       // [heads ic][9][heads oc]
  float* tile = dsm + heads * 9 * heads;                  // [heads][TR][TC] zero padded
  __shared__ float red[8][8];
  for (int i = threadIdx.x; i < heads * heads * 9; i += blockDim.x) {
    int oc = i / (heads * 9), rem = i - oc * heads * 9;
    int ic = rem / 9, k = rem - ic * 9;
    wsm[(ic * 9 + k) * heads + oc] = fw[i];
  }
  for (int rowid = threadIdx.x >> 5; rowid < heads * TR; rowid += blockDim.x >> 5) {
    int ic = rowid / TR, ty = rowid - ic * TR;
    int y = y0 + ty - 1;
    bool yok = y >= 0 && y < P && ty < rows + 2;
    const float* src = a + ((static_cast<int64_t>(b) * heads + ic) * P + (yok ? y : 0)) * R;
    for (int tx = threadIdx.x & 31; tx < TC; tx += 32) {
      int x = tx - 1;
      tile[rowid * TC + tx] = (yok && x >= 0 && x < R) ? src[x] : 0.f;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ocg = warp & 3, half = warp >> 2;
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (int pix = half * 32 + lane; pix < rows * R; pix += 64) {
    int ty = pix / R, tx = pix - ty * R;
    float acc[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = fb[ocg * 4 + o];
#pragma unroll 2
    for (int ic = 0; ic < heads; ++ic) {
      const float* t = tile + (ic * TR + ty) * TC + tx;
      const float* wk = wsm + ic * 9 * heads + ocg * 4;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        float v = t[(k / 3) * TC + (k % 3)];
        float4 w4 = *reinterpret_cast<const float4*>(wk + k * heads);
        acc[0] = fmaf(w4.x, v, acc[0]);
        acc[1] = fmaf(w4.y, v, acc[1]);
        acc[2] = fmaf(w4.z, v, acc[2]);
        acc[3] = fmaf(w4.w, v, acc[3]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int64_t oi = ((static_cast<int64_t>(b) * heads + ocg * 4 + o) * P + y0 + ty) * R + tx;
      if (add) acc[o] += add[oi];
      raw[oi] = acc[o];
      s[o] += acc[o];
      ss[o] += acc[o] * acc[o];
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    s[o] = gwd_warp_sum(s[o]);
    ss[o] = gwd_warp_sum(ss[o]);
  }
  if (lane == 0) {
#pragma unroll
    for (int o = 0; o < 4; ++o) { red[warp][o] = s[o]; red[warp][4 + o] = ss[o]; }
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // 16 channels x {sum, sumsq}: one fp64 atomic per CTA each
    int oc = threadIdx.x >> 1, which = threadIdx.x & 1;
    int g = oc >> 2, o = oc & 3;
    double v = static_cast<double>(red[g][which * 4 + o]) + static_cast<double>(red[4 + g][which * 4 + o]);
    atomicAdd(&stats[(static_cast<int64_t>(b) * heads + oc) * 2 + which], v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The same convolution as an implicit GEMM on warp-level tensor-core MMAs (m16n8k8, TF32 operands, fp32 accumulate):
// M = the band's pixels, N = 16 output channels, K = 9 taps x 16 input channels.  The scores feed a soft-max over the
// reference axis after three diffusion rounds, so fp32-level accuracy is kept with the 3xTF32 split
// (a_hi b_hi + a_lo b_hi + a_hi b_lo); the MMA count is irrelevant here (1.3 GFLOP per launch), the win is that one
// A fragment of 4 shared-memory loads feeds 6 MMAs instead of 2 loads per 4 FMAs.  (direct kernel: 103 us, LSU bound.)
// The weight fragments are laid out once per CTA as [k-step][n-tile][lane] float4 = {b0_hi, b1_hi, b0_lo, b1_lo}.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kDiffWarps = 9;
__global__ void __launch_bounds__(kDiffWarps * 32) gwd_ref_diffuse_mma_kernel(const float* __restrict__ a, float* __restrict__ raw,
                                                                           const __grid_constant__ DiffuseFilter flt,
                                                                           double* __restrict__ stats, int P, int R,
                                                                           int plane, int terms, int nimg,
                                                                           const float* __restrict__ fdev, const float* __restrict__ add) {
  extern __shared__ __align__(16) float dsm[];
  const float* fw = fdev ? fdev : flt.w;      // see gwd_ref_diffuse_conv_kernel
  const float* fb = fdev ? fdev + kDiffHeads * kDiffHeads * 9 : flt.b;
  constexpr int heads = kDiffHeads;
  constexpr int TR = kDiffBand + 2;
  const int TC = R + 2;
  float4* wfrag = reinterpret_cast<float4*>(dsm);             // [18 k-steps][2 n-tiles][32 lanes]
  float* tile = dsm + 18 * 2 * 32 * 4;                         // [heads][plane]  (plane >= TR*TC, == 8 mod 32: no bank conflicts)
  __shared__ float red[kDiffWarps][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // the filter leaves the parameter bank with consecutive lanes on consecutive words (a per-lane gather out of the
  // constant bank is serialised address by address), parked in the tile area, then re-laid-out as fragments
  for (int i = tid; i < heads * heads * 9; i += blockDim.x) tile[i] = fw[i];
  __syncthreads();
  // weight fragments: k-step ks = tap * 2 + kb covers input channels 8 kb .. 8 kb + 7 of tap `tap`
  for (int i = tid; i < 18 * 2 * 32; i += blockDim.x) {
    const int l = i & 31, nt = (i >> 5) & 1, ks = i >> 6;
    const int tap = ks >> 1, kb = ks & 1;
    const int oc = 8 * nt + (l >> 2), ic = 8 * kb + (l & 3);
    const float w0 = tile[(oc * heads + ic) * 9 + tap], w1 = tile[(oc * heads + ic + 4) * 9 + tap];
    const float h0 = __uint_as_float(__float_as_uint(w0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(w1) & 0xffffe000u);
    wfrag[i] = make_float4(h0, h1, w0 - h0, w1 - h1);
  }
  __syncthreads();
  // persistent CTAs: the fragments above are built once, then the CTA walks (band, image) items
  const int bands = (P + kDiffBand - 1) / kDiffBand;
  for (int item = blockIdx.x; item < bands * nimg; item += gridDim.x) {
  const int b = item / bands, y0 = (item - b * bands) * kDiffBand;
  const int rows = min(kDiffBand, P - y0);
  // input band with its halo, zero padded (cp.async with a zero source size): a warp per (channel, row), every request
  // of the thread in flight before the first one is waited for
  {
    const float* img = a + static_cast<int64_t>(b) * heads * P * R;
    for (int rowid = warp; rowid < heads * TR; rowid += kDiffWarps) {
      const int ic = rowid / TR, ty = rowid - ic * TR;
      const int y = y0 + ty - 1;
      const bool yok = y >= 0 && y < P && ty < rows + 2;
      const float* srow = img + (static_cast<int64_t>(ic) * P + (yok ? y : 0)) * R;
      float* drow = tile + ic * plane + ty * TC;
      for (int tx = lane; tx < TC; tx += 32) {
        const int x = tx - 1;
        const bool ok = yok && x >= 0 && x < R;
        const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(drow + tx));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(ok ? srow + x : img), "r"(ok ? 4 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int npix = rows * R;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};   // this lane's channels: 2t, 2t+1, 8+2t, 8+2t+1
  for (int mt = warp; mt * 16 < npix; mt += kDiffWarps) {
    int p0 = mt * 16 + g, p1 = p0 + 8;
    const bool ok0 = p0 < npix, ok1 = p1 < npix;
    if (!ok0) p0 = npix - 1;
    if (!ok1) p1 = npix - 1;
    const int ty0 = p0 / R, tx0 = p0 - ty0 * R, ty1 = p1 / R, tx1 = p1 - ty1 * R;
    const float* base0 = tile + t * plane + ty0 * TC + tx0;
    const float* base1 = tile + t * plane + ty1 * TC + tx1;
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      acc[nt][0] = fb[8 * nt + 2 * t];
      acc[nt][1] = fb[8 * nt + 2 * t + 1];
      acc[nt][2] = acc[nt][0];
      acc[nt][3] = acc[nt][1];
    }
    // operand split without conversions: the tensor core reads only the TF32 part (sign, exponent, 10 mantissa bits)
    // of a 32-bit operand, so hi = v as is and lo = v - trunc_tf32(v) (exact in fp32) give the 3xTF32 terms with one
    // LOP3 + one FADD per element (cvt.rna.tf32 expands to ~5 instructions here)
    const int p4 = 4 * plane;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float* r0 = base0 + dy * TC;
      const float* r1 = base1 + dy * TC;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const float v[4] = {r0[2 * kb * p4 + dx], r1[2 * kb * p4 + dx], r0[(2 * kb + 1) * p4 + dx], r1[(2 * kb + 1) * p4 + dx]};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            hi[i] = __float_as_uint(v[i]);
            lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i] & 0xffffe000u));
          }
          const int ks = (dy * 3 + dx) * 2 + kb;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float4 w = wfrag[(ks * 2 + nt) * 32 + lane];
            if (terms > 1) mma_tf32(acc[nt], lo, __float_as_uint(w.x), __float_as_uint(w.y));
            if (terms > 2) mma_tf32(acc[nt], hi, __float_as_uint(w.z), __float_as_uint(w.w));
            mma_tf32(acc[nt], hi, __float_as_uint(w.x), __float_as_uint(w.y));
          }
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int oc = 8 * nt + 2 * t + j;
        float* dst = raw + ((static_cast<int64_t>(b) * heads + oc) * P + y0) * R;
        if (add) {
          const float* ad = add + ((static_cast<int64_t>(b) * heads + oc) * P + y0) * R;
          if (ok0) acc[nt][j] += ad[ty0 * R + tx0];
          if (ok1) acc[nt][2 + j] += ad[ty1 * R + tx1];
        }
        if (ok0) {
          dst[ty0 * R + tx0] = acc[nt][j];
          s[2 * nt + j] += acc[nt][j];
          ss[2 * nt + j] = fmaf(acc[nt][j], acc[nt][j], ss[2 * nt + j]);
        }
        if (ok1) {
          dst[ty1 * R + tx1] = acc[nt][2 + j];
          s[2 * nt + j] += acc[nt][2 + j];
          ss[2 * nt + j] = fmaf(acc[nt][2 + j], acc[nt][2 + j], ss[2 * nt + j]);
        }
      }
    }
  }
  // per-channel sums: over the 8 pixel lanes (xor 4, 8, 16), then over the warps, then one fp64 atomic per CTA
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
      ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
    }
  }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oc = 8 * (i >> 1) + 2 * t + (i & 1);
      red[warp][oc * 2] = s[i];
      red[warp][oc * 2 + 1] = ss[i];
    }
  }
  __syncthreads();
  if (tid < 32) {
    double v = 0.0;
    for (int w = 0; w < kDiffWarps; ++w) v += static_cast<double>(red[w][tid]);
    atomicAdd(&stats[(static_cast<int64_t>(b) * heads) * 2 + tid], v);
  }
  __syncthreads();   // `red` and the tile are reused by the next item
  }
}

// phase 2: a_out = a_in + gelu((raw - mean) * rstd)   with mean / var over the whole [P,R] image of (b, channel)
// grid (chunks, B * channels): the statistics are read once per thread, the plane is walked with 16-byte vectors
__global__ void __launch_bounds__(256) gwd_ref_diffuse_norm_kernel(const float* __restrict__ a_in, const float* __restrict__ raw,
                                                                   const double* __restrict__ stats, float* __restrict__ a_out,
                                                                   int per_img) {
  const int img = blockIdx.y;
  const double mean_d = stats[img * 2] / per_img;
  const double var = stats[img * 2 + 1] / per_img - mean_d * mean_d;
  const float mean = static_cast<float>(mean_d);
  const float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
  const int64_t off = static_cast<int64_t>(img) * per_img;
  const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if ((per_img & 3) == 0) {
    const float4* ai = reinterpret_cast<const float4*>(a_in + off);
    const float4* ri = reinterpret_cast<const float4*>(raw + off);
    float4* ao = reinterpret_cast<float4*>(a_out + off);
    for (int i = i0; i < (per_img >> 2); i += stride) {
      const float4 x = ai[i], r = ri[i];
      float4 y;
      y.x = x.x + gwd_gelu((r.x - mean) * rstd);
      y.y = x.y + gwd_gelu((r.y - mean) * rstd);
      y.z = x.z + gwd_gelu((r.z - mean) * rstd);
      y.w = x.w + gwd_gelu((r.w - mean) * rstd);
      ao[i] = y;
    }
  } else {
    for (int i = i0; i < per_img; i += stride) a_out[off + i] = a_in[off + i] + gwd_gelu((raw[off + i] - mean) * rstd);
  }
}

// q_new[b*T+tok][h*hd+d] = scale * sum_r softmax_r(a[b][h][tok][:])[r] * refv[b*R+r][h*hd+d]   (bf16 out)
// grid (token tiles, heads, B), one thread per token, ref_v of the (image, head) in shared memory
__global__ void __launch_bounds__(128) gwd_ref_requery_kernel(const float* __restrict__ a, const float* __restrict__ refv,
                                                             int64_t ref_rs, bf16* __restrict__ out, int64_t o_rs, int T,
                                                             int heads, int hd, int R, float scale) {
  extern __shared__ __align__(16) float rv[];  // [R][hd] reference values, then [128][R + 1] staged score rows
  float* stile = rv + R * hd;
  const int b = blockIdx.z, h = blockIdx.y;
  const int tok0 = blockIdx.x * blockDim.x, tok = tok0 + threadIdx.x;
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    int r = i / hd, dd = i - r * hd;
    rv[i] = refv[(static_cast<int64_t>(b) * R + r) * ref_rs + h * hd + dd];
  }
  {  // the tile's score rows are one contiguous block: coalesced read, one padded row per thread afterwards
    const int ntok = min(static_cast<int>(blockDim.x), T - tok0);
    const float* src = a + ((static_cast<int64_t>(b) * heads + h) * T + tok0) * R;
    for (int i = threadIdx.x; i < ntok * R; i += blockDim.x) {
      const int tl = i / R, r = i - tl * R;
      stile[tl * (R + 1) + r] = __ldg(src + i);
    }
  }
  __syncthreads();
  if (tok >= T) return;
  const float* row = stile + threadIdx.x * (R + 1);
  float mx = -INFINITY;
  for (int r = 0; r < R; ++r) mx = fmaxf(mx, row[r]);
  float acc[32];
#pragma unroll
  for (int dd = 0; dd < 32; ++dd) acc[dd] = 0.f;
  float sum = 0.f;
  for (int r = 0; r < R; ++r) {
    float e = __expf(row[r] - mx);
    sum += e;
#pragma unroll
    for (int dd = 0; dd < 32; ++dd)
      if (dd < hd) acc[dd] = fmaf(e, rv[r * hd + dd], acc[dd]);
  }
  float inv = scale / sum;
  bf16* o = out + (static_cast<int64_t>(b) * T + tok) * o_rs + h * hd;
  if ((hd & 7) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      if (8 * w < hd) {
        uint4 u;
        u.x = gwd_pack_bf16x2(acc[8 * w + 0] * inv, acc[8 * w + 1] * inv);
        u.y = gwd_pack_bf16x2(acc[8 * w + 2] * inv, acc[8 * w + 3] * inv);
        u.z = gwd_pack_bf16x2(acc[8 * w + 4] * inv, acc[8 * w + 5] * inv);
        u.w = gwd_pack_bf16x2(acc[8 * w + 6] * inv, acc[8 * w + 7] * inv);
        reinterpret_cast<uint4*>(o)[w] = u;
      }
    }
  } else {
#pragma unroll
    for (int dd = 0; dd < 32; dd += 2)
      if (dd < hd) *reinterpret_cast<uint32_t*>(o + dd) = gwd_pack_bf16x2(acc[dd] * inv, acc[dd + 1] * inv);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
int gwd_attention_tc_try(const gwd_attn_desc* d, cudaStream_t stream);       // gwd_attn_tc.cu (tcgen05 path)
int gwd_attention_window_try(const gwd_attn_desc* d, cudaStream_t stream);   // gwd_attn_win.cu (biased windows)
int gwd_token_attention_mma_try(const void* dq, const void* sq, const void* tk, const void* tv, void* dout, void* sout,
                                int items, int N, int heads, int td, int tc, int64_t q_rs, int64_t k_rs, int64_t v_rs,
                                int64_t o_rs, float scale, cudaStream_t stream);   // gwd_attn_win.cu (mma.sync path)

extern "C" int gwd_attention(const gwd_attn_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(d && d->q && d->k && d->v && d->o, "gwd_attention: null pointer");
  GWD_CHECK_ARG(d->hd == 4 || d->hd == 8 || d->hd == 16 || d->hd == 32, "gwd_attention: head dim %d unsupported", d->hd);
  GWD_CHECK_ARG(d->items > 0 && d->heads > 0 && d->Lq > 0 && d->Lk > 0, "gwd_attention: empty problem");
  GWD_CHECK_ARG(d->k_row_stride % 2 == 0 && d->v_row_stride % 2 == 0 && d->k_item_stride % 2 == 0 && d->v_item_stride % 2 == 0 &&
                    (reinterpret_cast<uintptr_t>(d->k) & 3) == 0 && (reinterpret_cast<uintptr_t>(d->v) & 3) == 0,
                "gwd_attention: K/V must be 4-byte aligned with even strides");
  {  // DETR-shaped problems (head_dim 32, no bias / window mask, Lk <= 480) run on the tensor cores
    static const bool tc_enabled = []() { const char* e = getenv("GWD_ATTN_TC"); return !(e && e[0] == '0'); }();
    if (tc_enabled) {
      int rc = gwd_attention_tc_try(d, stream);
      if (rc <= 0) return rc;   // 0 = launched, < 0 = error, 1 = not eligible
    }
  }
  GWD_CHECK_ARG(!(d->dropout_seed != nullptr && d->dropout_p > 0.f),
                "gwd_attention: dropout is built on the tcgen05 path only (head dim 32, scale 1, no bias / window mask)");
  {  // biased (shifted-)window attention, N <= 64 tokens: persistent-CTA kernel with the bias table in shared memory
    static const bool win_enabled = []() { const char* e = getenv("GWD_ATTN_WINDOW"); return !(e && e[0] == '0'); }();
    if (win_enabled) {
      int rc = gwd_attention_window_try(d, stream);
      if (rc <= 0) return rc;
    }
  }
  AttnParams p;
  p.q = static_cast<const bf16*>(d->q); p.k = static_cast<const bf16*>(d->k); p.v = static_cast<const bf16*>(d->v);
  p.o = static_cast<bf16*>(d->o);
  p.items = d->items; p.heads = d->heads; p.Lq = d->Lq; p.Lk = d->Lk; p.hd = d->hd;
  p.q_is = d->q_item_stride; p.q_rs = d->q_row_stride; p.k_is = d->k_item_stride; p.k_rs = d->k_row_stride;
  p.v_is = d->v_item_stride; p.v_rs = d->v_row_stride; p.o_is = d->o_item_stride; p.o_rs = d->o_row_stride;
  p.bias = d->bias; p.mask = d->mask; p.nW = d->mask_windows > 0 ? d->mask_windows : 1;
  p.kpm = d->key_padding; p.scale = d->scale;
  p.q_tile = 0;
  GWD_CHECK_ARG(d->q_row_stride % 2 == 0 && d->q_item_stride % 2 == 0 && d->o_row_stride % 2 == 0 &&
                    d->o_item_stride % 2 == 0 && (reinterpret_cast<uintptr_t>(d->q) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->o) & 3) == 0,
                "gwd_attention: Q/O must be 4-byte aligned with even strides");
  switch (d->hd) {
    case 4: return launch_attention_tq<4>(p, stream);
    case 8: return launch_attention_tq<8>(p, stream);
    case 16: return launch_attention_tq<16>(p, stream);
    default: return launch_attention_tq<32>(p, stream);
  }
}

extern "C" int gwd_token_attention(const void* dq, const void* sq, const void* tk, const void* tv, void* dout, void* sout,
                                   int32_t items, int32_t N, int32_t heads, int32_t td, int32_t tc, int64_t q_rs,
                                   int64_t k_rs, int64_t v_rs, int64_t o_rs, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(dq && sq && tk && tv && dout && sout, "gwd_token_attention: null pointer");
  GWD_CHECK_ARG(heads > 0 && heads <= 32 && td > 0 && tc > 0 && items > 0, "gwd_token_attention: bad shape");
  {
    static const bool mma_enabled = []() { const char* e = getenv("GWD_TOKEN_MMA"); return !(e && e[0] == '0'); }();
    if (mma_enabled) {
      int rc = gwd_token_attention_mma_try(dq, sq, tk, tv, dout, sout, items, N, heads, td, tc, q_rs, k_rs, v_rs, o_rs, scale,
                                           stream);
      if (rc <= 0) return rc;
    }
  }
  TokAttnParams p;
  p.dq = static_cast<const bf16*>(dq); p.sq = static_cast<const bf16*>(sq);
  p.tk = static_cast<const bf16*>(tk); p.tv = static_cast<const bf16*>(tv);
  p.dout = static_cast<bf16*>(dout); p.sout = static_cast<bf16*>(sout);
  p.items = items; p.N = N; p.heads = heads; p.td = td; p.tc = tc;
  p.q_rs = q_rs; p.k_rs = k_rs; p.v_rs = v_rs; p.o_rs = o_rs; p.scale = scale;
  GWD_CHECK_ARG((heads * td) % 2 == 0 && (heads * tc) % 2 == 0 && q_rs % 2 == 0 && k_rs % 2 == 0 && v_rs % 2 == 0,
                "gwd_token_attention: widths / strides must be even");
  size_t smem = (static_cast<size_t>(N) * heads * tc * 2 + static_cast<size_t>(N) * heads * td * 2) * 2 + 16 +
                static_cast<size_t>(heads) * 2 * td * tc * sizeof(float);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_token_attention: window does not fit shared memory");
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_token_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  gwd_token_attention_kernel<<<items, heads * 32, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_scores(const void* q, int64_t q_rs, const float* refk, int64_t ref_rs, float* out, int32_t B,
                              int32_t nW, int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(q && refk && out && hd <= 32 && hd % 2 == 0, "gwd_ref_scores: bad argument");
  const int T = nW * N;
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(T, 128)), heads, B);
  gwd_ref_scores_kernel<<<grid, 128, (static_cast<size_t>(R) * hd + 128 * (R + 1)) * sizeof(float), stream>>>(
      static_cast<const bf16*>(q), q_rs, refk, ref_rs, out, T, heads, hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}

// conv phase of one diffusion round: raw = conv3x3(a_in) (+ add), per-(image, channel) sum / sum of squares into stats.
// The filter is either the by-value host copy `flt` (inference plan) or the device array `fdev` (training).
static int launch_diffuse_conv(const float* a_in, const DiffuseFilter& flt, const float* fdev, const float* add, float* raw,
                               double* stats_ws, int B, int heads, int P, int R, cudaStream_t stream) {
  GWD_CHECK_ARG(heads == kDiffHeads, "gwd_ref_diffuse: built for %d heads (got %d)", kDiffHeads, heads);
  GWD_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B * heads, stream));
  static const bool use_mma = []() { const char* e = getenv("GWD_DIFFUSE_MMA"); return !(e && e[0] == '0'); }();
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(P, kDiffBand)), B);
  int plane = (kDiffBand + 2) * (R + 2);
  plane += ((8 - plane % 32) + 32) % 32;       // plane stride == 8 (mod 32): the 4 channels x 8 pixels of a fragment load hit 32 banks
  // the MMA kernel parks the filter in its tile area while it builds the fragments: very few reference points (R < 14) do not
  // give it the room -> direct kernel
  if (use_mma && heads * plane >= heads * heads * 9) {
    size_t smem = (static_cast<size_t>(heads) * plane + 18 * 2 * 32 * 4) * sizeof(float);
    GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_ref_diffuse: %d reference points do not fit shared memory", R);
    if (smem > 48 * 1024) {
      static bool configured = false;
      if (!configured) {
        GWD_CUDA(cudaFuncSetAttribute(gwd_ref_diffuse_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
      }
    }
    static const int terms = []() { const char* e = getenv("GWD_DIFFUSE_TERMS"); return e ? atoi(e) : 3; }();
    int per_sm = static_cast<int>((200 * 1024) / (smem + 1024));
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    unsigned ctas = static_cast<unsigned>(per_sm * gwd_num_sms());
    if (ctas > grid.x * grid.y) ctas = grid.x * grid.y;
    gwd_ref_diffuse_mma_kernel<<<ctas, kDiffWarps * 32, smem, stream>>>(a_in, raw, flt, stats_ws, P, R, plane, terms, B, fdev, add);
  } else {
    size_t smem = (static_cast<size_t>(heads) * (kDiffBand + 2) * (R + 2) + static_cast<size_t>(heads) * heads * 9) * sizeof(float);
    GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_ref_diffuse: %d reference points do not fit shared memory", R);
    if (smem > 48 * 1024) {
      static bool configured = false;
      if (!configured) {
        GWD_CUDA(cudaFuncSetAttribute(gwd_ref_diffuse_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
      }
    }
    gwd_ref_diffuse_conv_kernel<<<grid, 256, smem, stream>>>(a_in, raw, flt, stats_ws, P, R, fdev, add);
  }
  GWD_LAUNCHED();
  return GWD_OK;
}

static int launch_diffuse_norm(const float* a_in, const float* raw, const double* stats_ws, float* a_out, int B, int heads, int P,
                               int R, cudaStream_t stream) {
  const int per_img = P * R;
  int chunks = static_cast<int>(gwd_ceil_div(per_img, 256 * 4 * 4));
  if (chunks < 1) chunks = 1;
  gwd_ref_diffuse_norm_kernel<<<dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(B * heads)), 256, 0, stream>>>(
      a_in, raw, stats_ws, a_out, per_img);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_diffuse(const float* a_in, float* a_out, const float* w_host, const float* bias_host, float* raw_ws,
                               double* stats_ws, int32_t B, int32_t heads, int32_t P, int32_t R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a_in && a_out && w_host && bias_host && raw_ws && stats_ws && a_in != a_out,
                "gwd_ref_diffuse: null / aliased pointer");
  DiffuseFilter flt;
  memcpy(flt.w, w_host, sizeof(flt.w));
  memcpy(flt.b, bias_host, sizeof(flt.b));
  int rc = launch_diffuse_conv(a_in, flt, nullptr, nullptr, raw_ws, stats_ws, B, heads, P, R, stream);
  if (rc != GWD_OK) return rc;
  return launch_diffuse_norm(a_in, raw_ws, stats_ws, a_out, B, heads, P, R, stream);
}

// training forward: the same round with the filter read from DEVICE memory (filt_dev = [oc][ic][ky][kx] fp32 then 16
// biases, produced by gwd_diffuse_filter_pack); raw / stats are kept by the caller for gwd_ref_diffuse_bwd
extern "C" int gwd_ref_diffuse_dev(const float* a_in, float* a_out, const float* filt_dev, float* raw, double* stats,
                                   int32_t B, int32_t heads, int32_t P, int32_t R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a_in && a_out && filt_dev && raw && stats && a_in != a_out, "gwd_ref_diffuse_dev: null / aliased pointer");
  static const DiffuseFilter none = {};
  int rc = launch_diffuse_conv(a_in, none, filt_dev, nullptr, raw, stats, B, heads, P, R, stream);
  if (rc != GWD_OK) return rc;
  return launch_diffuse_norm(a_in, raw, stats, a_out, B, heads, P, R, stream);
}

// out = add + conv3x3(a_in) with a device filter: the data-gradient convolution of the diffusion backward (the caller
// passes the flipped / transposed filter with zero biases); stats_scratch: fp64 [2 * B * heads] scratch
extern "C" int gwd_ref_diffuse_conv_dev(const float* a_in, const float* filt_dev, const float* add, float* out,
                                        double* stats_scratch, int32_t B, int32_t heads, int32_t P, int32_t R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a_in && filt_dev && out && stats_scratch && a_in != out, "gwd_ref_diffuse_conv_dev: null / aliased pointer");
  static const DiffuseFilter none = {};
  return launch_diffuse_conv(a_in, none, filt_dev, add, out, stats_scratch, B, heads, P, R, stream);
}

extern "C" int gwd_ref_requery(const float* a, const float* refv, int64_t ref_rs, void* out, int64_t o_rs, int32_t B,
                               int32_t nW, int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a && refv && out && hd <= 32 && hd % 2 == 0 && o_rs % 2 == 0, "gwd_ref_requery: bad argument");
  const int T = nW * N;
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(T, 128)), heads, B);
  gwd_ref_requery_kernel<<<grid, 128, (static_cast<size_t>(R) * hd + 128 * (R + 1)) * sizeof(float), stream>>>(
      a, refv, ref_rs, static_cast<bf16*>(out), o_rs, T, heads, hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}
