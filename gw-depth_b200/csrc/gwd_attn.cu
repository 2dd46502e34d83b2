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
// window attention: one CTA per window, one warp per head.  The window's K and V rows ([N, C], all heads) are
// staged in shared memory with coalesced 16-byte loads; K rows are padded per head to kill bank conflicts.
// -------------------------------------------------------------------------------------------------
struct WinAttnParams {
  const bf16* q; const bf16* k; const bf16* v; bf16* o;
  int windows, heads, N, hd;
  int64_t q_rs, k_rs, v_rs, o_rs;      // row strides; window stride = N * row stride
  const float* bias;                   // [heads, N, N]
  const float* mask;                   // [nW, N, N] or null
  int nW;
};

__global__ void gwd_window_attention_kernel(const WinAttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int N = p.N, hd = p.hd, heads = p.heads, C = heads * hd;
  const int kst = hd + 2;                                 // padded per-head K row (bf16 elements)
  bf16* Ks = reinterpret_cast<bf16*>(smem);               // [heads][N][kst]
  bf16* Vs = Ks + static_cast<size_t>(heads) * N * kst;   // [N][C]
  float* Ps = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Vs + static_cast<size_t>(N) * C) + 15) & ~uintptr_t(15));
  float* Qs = Ps + static_cast<size_t>(heads) * 64;       // [heads][64] probabilities, then [heads][32] query
  const int win = blockIdx.x;
  const bf16* kb = p.k + static_cast<int64_t>(win) * N * p.k_rs;
  const bf16* vb = p.v + static_cast<int64_t>(win) * N * p.v_rs;
  const int cw = C >> 1;                                  // 32-bit words per row
  for (int idx = threadIdx.x; idx < N * cw; idx += blockDim.x) {
    int j = idx / cw, w = idx - j * cw;
    uint32_t kk = reinterpret_cast<const uint32_t*>(kb + j * p.k_rs)[w];
    uint32_t vv = reinterpret_cast<const uint32_t*>(vb + j * p.v_rs)[w];
    int h = (2 * w) / hd, dd = 2 * w - h * hd;
    *reinterpret_cast<uint32_t*>(Ks + (static_cast<size_t>(h) * N + j) * kst + dd) = kk;
    reinterpret_cast<uint32_t*>(Vs + static_cast<size_t>(j) * C)[w] = vv;
  }
  __syncthreads();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= heads) return;
  float* ps = Ps + h * 64;
  float* qs = Qs + h * 32;
  const bf16* Kh = Ks + static_cast<size_t>(h) * N * kst;
  const int hw = hd >> 1;
  const int G = 32 / hd, d = lane % hd, g = lane / hd;
  const float* mbase = p.mask ? p.mask + static_cast<int64_t>(win % p.nW) * N * N : nullptr;
  for (int qi = 0; qi < N; ++qi) {
    const bf16* qrow = p.q + (static_cast<int64_t>(win) * N + qi) * p.q_rs + h * hd;
    if (lane < hd) qs[lane] = __bfloat162float(qrow[lane]);
    __syncwarp();
    const float* brow = p.bias + (static_cast<int64_t>(h) * N + qi) * N;
    float sc[2];
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int j = lane + 32 * r;
      float s = -INFINITY;
      if (j < N) {
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(Kh + static_cast<size_t>(j) * kst);
        s = 0.f;
        for (int w = 0; w < hw; ++w) {
          float2 kk = gwd_unpack_bf16x2(kr[w]);
          s = fmaf(qs[2 * w], kk.x, s);
          s = fmaf(qs[2 * w + 1], kk.y, s);
        }
        s += brow[j];
        if (mbase) s += mbase[qi * N + j];
      }
      sc[r] = s;
      mx = fmaxf(mx, s);
    }
    mx = gwd_warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int j = lane + 32 * r;
      float e = j < N ? __expf(sc[r] - mx) : 0.f;
      ps[j] = e;
      sum += e;
    }
    sum = gwd_warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int j = g; j < N; j += G) acc = fmaf(ps[j], __bfloat162float(Vs[static_cast<size_t>(j) * C + h * hd + d]), acc);
    for (int o = hd; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane < hd) p.o[(static_cast<int64_t>(win) * N + qi) * p.o_rs + h * hd + lane] = __float2bfloat16(acc / sum);
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
  extern __shared__ float rk[];  // [R][hd]
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    int r = i / hd, dd = i - r * hd;
    rk[i] = refk[(static_cast<int64_t>(b) * R + r) * ref_rs + h * hd + dd] * scale;
  }
  __syncthreads();
  int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= T) return;
  float qv[32];
  const bf16* qr = q + (static_cast<int64_t>(b) * T + tok) * q_rs + h * hd;
#pragma unroll
  for (int dd = 0; dd < 32; ++dd) qv[dd] = dd < hd ? __bfloat162float(qr[dd]) : 0.f;
  float* o = out + ((static_cast<int64_t>(b) * heads + h) * T + tok) * R;
  for (int r = 0; r < R; ++r) {
    float sacc = 0.f;
#pragma unroll
    for (int dd = 0; dd < 32; ++dd)
      if (dd < hd) sacc = fmaf(qv[dd], rk[r * hd + dd], sacc);
    o[r] = sacc;
  }
}

// diffusion step, phase 1: raw = conv3x3_{heads->heads}(a) for a band of rows of one image, all output channels;
// per-(image, channel) sum / sum of squares accumulated in fp64 for the whole-image LayerNorm of phase 2.
// a: [B][heads][P][R] fp32.  grid (row bands, B), block 256.
constexpr int kDiffBand = 21;
__global__ void __launch_bounds__(256) gwd_ref_diffuse_conv_kernel(const float* __restrict__ a, float* __restrict__ raw,
                                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                                   double* __restrict__ stats, int heads, int P, int R) {
  extern __shared__ float sm[];
  float* tile = sm;                                       // [heads][band+2][R+2] zero padded
  const int y0 = blockIdx.x * kDiffBand, b = blockIdx.y;
  const int rows = min(kDiffBand, P - y0);
  const int TR = kDiffBand + 2, TC = R + 2;
  float* wsm = tile + static_cast<size_t>(heads) * TR * TC;  // [heads*heads*9]
  for (int i = threadIdx.x; i < heads * heads * 9; i += blockDim.x) wsm[i] = w[i];
  for (int i = threadIdx.x; i < heads * TR * TC; i += blockDim.x) {
    int ic = i / (TR * TC);
    int rem = i - ic * TR * TC;
    int ty = rem / TC, tx = rem - ty * TC;
    int y = y0 + ty - 1, x = tx - 1;
    float v = 0.f;
    if (y >= 0 && y < P && x >= 0 && x < R && ty < rows + 2) v = a[((static_cast<int64_t>(b) * heads + ic) * P + y) * R + x];
    tile[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // each warp owns output channels oc = warp, warp + 8, ...; lanes stride over the band's pixels
  for (int oc = warp; oc < heads; oc += 8) {
    float s = 0.f, ss = 0.f;
    for (int pix = lane; pix < rows * R; pix += 32) {
      int ty = pix / R, tx = pix - ty * R;
      float acc = bias[oc];
      for (int ic = 0; ic < heads; ++ic) {
        const float* t = tile + (static_cast<size_t>(ic) * TR + ty) * TC + tx;
        const float* wk = wsm + (oc * heads + ic) * 9;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) acc = fmaf(wk[dy * 3 + dx], t[dy * TC + dx], acc);
      }
      raw[((static_cast<int64_t>(b) * heads + oc) * P + y0 + ty) * R + tx] = acc;
      s += acc;
      ss += acc * acc;
    }
    double ds = s, dss = ss;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
      dss += __shfl_xor_sync(0xffffffffu, dss, o);
    }
    if (lane == 0) {
      atomicAdd(&stats[(static_cast<int64_t>(b) * heads + oc) * 2], ds);
      atomicAdd(&stats[(static_cast<int64_t>(b) * heads + oc) * 2 + 1], dss);
    }
  }
}

// phase 2: a_out = a_in + gelu((raw - mean) * rstd)   with mean / var over the whole [P,R] image of (b, channel)
__global__ void gwd_ref_diffuse_norm_kernel(const float* __restrict__ a_in, const float* __restrict__ raw,
                                            const double* __restrict__ stats, float* __restrict__ a_out, int64_t per_img,
                                            int64_t total) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t img = i / per_img;
    double mean = stats[img * 2] / per_img;
    double var = stats[img * 2 + 1] / per_img - mean * mean;
    float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-5f);
    float v = (raw[i] - static_cast<float>(mean)) * rstd;
    a_out[i] = a_in[i] + gwd_apply_act(v, GWD_ACT_GELU);
  }
}

// q_new[b*T+tok][h*hd+d] = scale * sum_r softmax_r(a[b][h][tok][:])[r] * refv[b*R+r][h*hd+d]   (bf16 out)
// grid (token tiles, heads, B), one thread per token, ref_v of the (image, head) in shared memory
__global__ void __launch_bounds__(128) gwd_ref_requery_kernel(const float* __restrict__ a, const float* __restrict__ refv,
                                                             int64_t ref_rs, bf16* __restrict__ out, int64_t o_rs, int T,
                                                             int heads, int hd, int R, float scale) {
  extern __shared__ float rv[];  // [R][hd]
  const int b = blockIdx.z, h = blockIdx.y;
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    int r = i / hd, dd = i - r * hd;
    rv[i] = refv[(static_cast<int64_t>(b) * R + r) * ref_rs + h * hd + dd];
  }
  __syncthreads();
  int tok = blockIdx.x * blockDim.x + threadIdx.x;
  if (tok >= T) return;
  const float* row = a + ((static_cast<int64_t>(b) * heads + h) * T + tok) * R;
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
#pragma unroll
  for (int dd = 0; dd < 32; dd += 2)
    if (dd < hd) *reinterpret_cast<uint32_t*>(o + dd) = gwd_pack_bf16x2(acc[dd] * inv, acc[dd + 1] * inv);
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
  if (d->Lq == d->Lk && d->Lk <= 64 && d->bias != nullptr && d->key_padding == nullptr && d->scale == 1.0f &&
      d->heads <= 32 && d->q_item_stride == d->Lq * d->q_row_stride && d->k_item_stride == d->Lk * d->k_row_stride &&
      d->v_item_stride == d->Lk * d->v_row_stride && d->o_item_stride == d->Lq * d->o_row_stride) {
    WinAttnParams w;
    w.q = static_cast<const bf16*>(d->q); w.k = static_cast<const bf16*>(d->k); w.v = static_cast<const bf16*>(d->v);
    w.o = static_cast<bf16*>(d->o);
    w.windows = d->items; w.heads = d->heads; w.N = d->Lq; w.hd = d->hd;
    w.q_rs = d->q_row_stride; w.k_rs = d->k_row_stride; w.v_rs = d->v_row_stride; w.o_rs = d->o_row_stride;
    w.bias = d->bias; w.mask = d->mask; w.nW = d->mask_windows > 0 ? d->mask_windows : 1;
    const int C = d->heads * d->hd;
    size_t smem = static_cast<size_t>(d->heads) * d->Lk * (d->hd + 2) * 2 + static_cast<size_t>(d->Lk) * C * 2 + 16 +
                  static_cast<size_t>(d->heads) * (64 + 32) * 4;
    if (smem > 48 * 1024) {
      static bool configured = false;
      if (!configured) {
        GWD_CUDA(cudaFuncSetAttribute(gwd_window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
      }
    }
    gwd_window_attention_kernel<<<d->items, d->heads * 32, smem, stream>>>(w);
    GWD_LAUNCHED();
    return GWD_OK;
  }
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
  gwd_ref_scores_kernel<<<grid, 128, static_cast<size_t>(R) * hd * sizeof(float), stream>>>(
      static_cast<const bf16*>(q), q_rs, refk, ref_rs, out, T, heads, hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_diffuse(const float* a_in, float* a_out, const float* w, const float* bias, float* raw_ws,
                               double* stats_ws, int32_t B, int32_t heads, int32_t P, int32_t R, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a_in && a_out && w && bias && raw_ws && stats_ws && a_in != a_out, "gwd_ref_diffuse: null / aliased pointer");
  GWD_CHECK_ARG(heads <= 16, "gwd_ref_diffuse: at most 16 heads");
  GWD_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B * heads, stream));
  size_t smem = (static_cast<size_t>(heads) * (kDiffBand + 2) * (R + 2) + static_cast<size_t>(heads) * heads * 9) * sizeof(float);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_ref_diffuse: %d reference points do not fit shared memory", R);
  if (smem > 48 * 1024) {
    static bool configured = false;
    if (!configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_ref_diffuse_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(P, kDiffBand)), B);
  gwd_ref_diffuse_conv_kernel<<<grid, 256, smem, stream>>>(a_in, raw_ws, w, bias, stats_ws, heads, P, R);
  GWD_LAUNCHED();
  int64_t per_img = static_cast<int64_t>(P) * R, total = per_img * B * heads;
  int blocks = static_cast<int>(std::min<int64_t>(gwd_ceil_div(total, 256), gwd_num_sms() * 8));
  gwd_ref_diffuse_norm_kernel<<<blocks, 256, 0, stream>>>(a_in, raw_ws, stats_ws, a_out, per_img, total);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_ref_requery(const float* a, const float* refv, int64_t ref_rs, void* out, int64_t o_rs, int32_t B,
                               int32_t nW, int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(a && refv && out && hd <= 32 && hd % 2 == 0 && o_rs % 2 == 0, "gwd_ref_requery: bad argument");
  const int T = nW * N;
  dim3 grid(static_cast<unsigned>(gwd_ceil_div(T, 128)), heads, B);
  gwd_ref_requery_kernel<<<grid, 128, static_cast<size_t>(R) * hd * sizeof(float), stream>>>(
      a, refv, ref_rs, static_cast<bf16*>(out), o_rs, T, heads, hd, R, scale);
  GWD_LAUNCHED();
  return GWD_OK;
}
