// gwd_elem.cu -- bandwidth-bound kernels of the GW-Depth forward: LayerNorm, row-broadcast add, the Swin
// window gather / merge (LayerNorm + pad + cyclic shift + partition folded into one pass each way),
// nearest / bilinear resampling, average pooling, point sampling and the anchor-mixture depth read-out.
// All activations are channels-last bf16; statistics and small depth maps are fp32.
// One warp owns one token (pixel) row: 16-byte vector loads, shuffle reductions, no shared memory.
#include <stdlib.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y), c = gwd_unpack_bf16x2(u.z), d = gwd_unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  u.x = gwd_pack_bf16x2(f[0], f[1]); u.y = gwd_pack_bf16x2(f[2], f[3]);
  u.z = gwd_pack_bf16x2(f[4], f[5]); u.w = gwd_pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// A token row is held by a GROUP of G lanes (G = 8, 16 or 32, a power of two >= C/8 capped at 32), so narrow rows
// (C = 64 at the 1/4 scale) keep all 32 lanes of a warp busy: a warp owns 32/G rows.  Vector v of sub-lane s covers
// channels (v*G + s)*8 .. +8.
struct RowGroup {
  int G, sub;        // group width, lane index inside the group
  int64_t row;       // row owned by this lane's group
};
__device__ __forceinline__ RowGroup row_group(int C) {
  RowGroup rg;
  int need = C >> 3;
  rg.G = need <= 8 ? 8 : (need <= 16 ? 16 : 32);
  int lane = threadIdx.x & 31;
  rg.sub = lane & (rg.G - 1);
  int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  rg.row = warp * (32 / rg.G) + lane / rg.G;
  return rg;
}
__host__ __device__ inline int rows_per_warp(int C) {
  int need = C >> 3;
  return need <= 8 ? 4 : (need <= 16 ? 2 : 1);
}
__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
struct WarpRow {
  float f[NV][8];
};

template <int NV>
__device__ __forceinline__ void row_load(WarpRow<NV>& r, const bf16* src, int C, const RowGroup& g, bool valid) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C && valid) load8(src + c, r.f[v]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) r.f[v][i] = 0.f;
    }
  }
}
template <int NV>
__device__ __forceinline__ void row_add(WarpRow<NV>& r, const bf16* src, int C, const RowGroup& g) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C) {
      float t[8];
      load8(src + c, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) r.f[v][i] += t[i];
    }
  }
}
template <int NV>
__device__ __forceinline__ void row_store(const WarpRow<NV>& r, bf16* dst, int C, const RowGroup& g) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C) store8(dst + c, r.f[v]);
  }
}
// LayerNorm over the first n channels (channels >= n are padding and come out as 0).  Every lane of the warp must call.
template <int NV>
__device__ __forceinline__ void row_layernorm(WarpRow<NV>& r, int C, int n, const RowGroup& g, const float* gam, const float* bet,
                                              float eps) {
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (c + i < n) s += r.f[v][i];
    }
  }
  float mean = group_sum(s, g.G) / n;
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (c + i < n) { float dlt = r.f[v][i] - mean; ss += dlt * dlt; }
    }
  }
  float rstd = rsqrtf(group_sum(ss, g.G) / n + eps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    int c = (v * g.G + g.sub) * 8;
    if (c < C) {
      // gamma / beta as 16-byte loads (padded to the physical width by the callers), (x - mean) * rstd * g + b as 2 FMAs
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gam + c)), g1 = __ldg(reinterpret_cast<const float4*>(gam + c) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bet + c)), b1 = __ldg(reinterpret_cast<const float4*>(bet + c) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float shift = -mean * rstd;
#pragma unroll
      for (int i = 0; i < 8; ++i) r.f[v][i] = (c + i < n) ? fmaf(fmaf(r.f[v][i], rstd, shift), gg[i], bb[i]) : 0.f;
    }
  }
}
template <int NV>
__device__ __forceinline__ void row_act(WarpRow<NV>& r, int act, int C, const RowGroup& g) {
  if (act == GWD_ACT_NONE) return;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    if ((v * g.G + g.sub) * 8 < C) {
      if (act == GWD_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.f[v][i] = gwd_gelu(r.f[v][i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.f[v][i] = gwd_apply_act(r.f[v][i], act);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// out[row] = act(LN(x[row] + res[row]))     (res, LN optional)
// ------------------------------------------------------------------------------------------------
template <int NV>
__global__ void gwd_layernorm_kernel(const bf16* x, int64_t x_rs, const bf16* res, int64_t res_rs, const float* g,
                                     const float* b, float eps, int act, bf16* out, int64_t out_rs, int64_t rows, int C,
                                     int n) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  RowGroup rg = row_group(C);
  bool live = rg.row < rows;
  int64_t row = live ? rg.row : 0;
  WarpRow<NV> r;
  row_load(r, x + row * x_rs, C, rg, live);
  if (res && live) row_add(r, res + row * res_rs, C, rg);
  if (g) row_layernorm(r, C, n, rg, g, b, eps);
  row_act(r, act, C, rg);
  if (live) row_store(r, out + row * out_rs, C, rg);
}

// the same for row widths that fill the power-of-two lane groups badly (C = 320: 40 vectors on 32 lanes x 2 slots = 37 % idle lanes,
// and the kernel is bound by instruction issue, not by HBM): groups of 8 lanes x NV = C / 64 slots, four rows per warp, no idle lane
template <int NV>
__global__ void gwd_layernorm_g8_kernel(const bf16* x, int64_t x_rs, const bf16* res, int64_t res_rs, const float* g,
                                        const float* b, float eps, int act, bf16* out, int64_t out_rs, int64_t rows, int C,
                                        int n) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  RowGroup rg;
  const int lane = threadIdx.x & 31;
  rg.G = 8;
  rg.sub = lane & 7;
  rg.row = ((blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5) * 4 + (lane >> 3);
  bool live = rg.row < rows;
  int64_t row = live ? rg.row : 0;
  WarpRow<NV> r;
  row_load(r, x + row * x_rs, C, rg, live);
  if (res && live) row_add(r, res + row * res_rs, C, rg);
  if (g) row_layernorm(r, C, n, rg, g, b, eps);
  row_act(r, act, C, rg);
  if (live) row_store(r, out + row * out_rs, C, rg);
}

// out[row] = x[row] + addend[row % period]
template <int NV>
__global__ void gwd_add_rows_kernel(const bf16* x, int64_t x_rs, const bf16* addend, int64_t a_rs, int64_t period,
                                    bf16* out, int64_t out_rs, int64_t rows, int C) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  RowGroup rg = row_group(C);
  if (rg.row >= rows) return;
  WarpRow<NV> r;
  row_load(r, x + rg.row * x_rs, C, rg, true);
  row_add(r, addend + (rg.row % period) * a_rs, C, rg);
  row_store(r, out + rg.row * out_rs, C, rg);
}

// ------------------------------------------------------------------------------------------------
// Swin window gather: out[(b*nW + w)*N + t] = LN(x[b, y, x]) (0 in the bottom/right padding), where the window
// grid lives on the cyclically shifted, padded map (multiscale_transformerr.py:659-707)
// ------------------------------------------------------------------------------------------------
struct WinGeom {
  int B, H, W, Hp, Wp, ws, shift;
};
// grid (row groups of one map row, padded map rows, images): the window arithmetic is one 32-bit division per thread (a linear row
// index decoded with 64-bit div / mod chains cost ~500 instructions per 16-byte vector and capped these kernels at 1.5 TB/s)
template <int NV>
__global__ void gwd_window_gather_kernel(const bf16* x, int64_t x_rs, const float* g, const float* b, float eps, bf16* out,
                                         int64_t out_rs, WinGeom gm, int C, int n) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  RowGroup rg = row_group(C);
  const int xs = static_cast<int>(rg.row), ys = blockIdx.y, bb = blockIdx.z;     // position on the shifted, padded map
  const bool live = xs < gm.Wp;
  const int nWx = gm.Wp / gm.ws, nWy = gm.Hp / gm.ws;
  const int wy = ys / gm.ws, wx = xs / gm.ws;
  const int64_t orow = ((static_cast<int64_t>(bb) * nWy + wy) * nWx + wx) * (gm.ws * gm.ws) + (ys - wy * gm.ws) * gm.ws + (xs - wx * gm.ws);
  int y = ys + gm.shift, xx = xs + gm.shift;
  if (y >= gm.Hp) y -= gm.Hp;
  if (xx >= gm.Wp) xx -= gm.Wp;
  const bool valid = live && y < gm.H && xx < gm.W;
  WarpRow<NV> r;
  row_load(r, x + ((static_cast<int64_t>(bb) * gm.H + (valid ? y : 0)) * gm.W + (valid ? xx : 0)) * x_rs, C, rg, valid);
  if (g) row_layernorm(r, C, n, rg, g, b, eps);
  if (!valid) {   // padding tokens are exact zeros, not LN(0)  (zero padding happens after norm1, :659-671)
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) r.f[v][i] = 0.f;
  }
  if (live) row_store(r, out + orow * out_rs, C, rg);
}

// Swin window merge: y[b,y,x] = shortcut[b,y,x] + win[row(b,y,x)]; optionally y_ln = LN(y)   (:731-755)
// grid (row groups of one map row, map rows, images).  (Several positions per lane group with their loads issued together were
// measured: 4 % slower on the whole forward -- these kernels are not bound by load latency.)
template <int NV>
__global__ void gwd_window_merge_kernel(const bf16* win, int64_t win_rs, const bf16* shortcut, int64_t sc_rs, bf16* out,
                                        int64_t out_rs, const float* g, const float* b, float eps, bf16* out_ln,
                                        int64_t ln_rs, WinGeom gm, int C, int n) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  RowGroup rg = row_group(C);
  const int x = static_cast<int>(rg.row), y = blockIdx.y, bb = blockIdx.z;
  const bool live = x < gm.W;
  const int64_t pix = (static_cast<int64_t>(bb) * gm.H + y) * gm.W + (live ? x : 0);
  int ys = y - gm.shift, xs = (live ? x : 0) - gm.shift;
  if (ys < 0) ys += gm.Hp;
  if (xs < 0) xs += gm.Wp;
  const int nWx = gm.Wp / gm.ws, nWy = gm.Hp / gm.ws;
  const int wy = ys / gm.ws, wx = xs / gm.ws;
  const int64_t wrow = ((static_cast<int64_t>(bb) * nWy + wy) * nWx + wx) * (gm.ws * gm.ws) + (ys - wy * gm.ws) * gm.ws + (xs - wx * gm.ws);
  WarpRow<NV> r;
  row_load(r, win + wrow * win_rs, C, rg, live);
  if (live) {
    row_add(r, shortcut + pix * sc_rs, C, rg);
    row_store(r, out + pix * out_rs, C, rg);
  }
  if (out_ln) {
    row_layernorm(r, C, n, rg, g, b, eps);
    if (live) row_store(r, out_ln + pix * ln_rs, C, rg);
  }
}

// ------------------------------------------------------------------------------------------------
// resampling (8-channel vectors, one thread per output vector)
// ------------------------------------------------------------------------------------------------
// nearest (legacy F.interpolate(mode='nearest'): src = floor(dst * in / out)); optional add of a second same-size map
__global__ void gwd_upsample_nearest_kernel(const bf16* x, int64_t x_rs, int B, int h, int w, bf16* out, int64_t out_rs,
                                            int H, int W, int C, const bf16* add, int64_t add_rs) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  // grid (vectors of one output row, rows, images): one division per thread (64-bit div / mod chains made these
  // resampling kernels instruction bound)
  const int cv = C / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * cv) return;
  const int X = t / cv, c = (t - X * cv) * 8;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int64_t pix = (static_cast<int64_t>(b) * H + Y) * W + X;
  int sy = min(static_cast<int>((static_cast<int64_t>(Y) * h) / H), h - 1);
  int sx = min(static_cast<int>((static_cast<int64_t>(X) * w) / W), w - 1);
  float f[8];
  load8(x + ((static_cast<int64_t>(b) * h + sy) * w + sx) * x_rs + c, f);
  if (add) {
    float u[8];
    load8(add + pix * add_rs + c, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += u[i];
  }
  store8(out + pix * out_rs + c, f);
}

// nn.AvgPool2d(k, stride=k), floor mode
__global__ void gwd_avgpool_kernel(const bf16* x, int64_t x_rs, int B, int H, int W, int k, bf16* out, int64_t out_rs,
                                   int C) {
  gwd_pdl_wait();      // launched programmatically (gwd_launch): nothing touches global memory before the kernel in front has completed
  gwd_pdl_trigger();
  const int cv = C / 8;
  const int oh = H / k, ow = W / k;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ow * cv) return;
  const int X = t / cv, c = (t - X * cv) * 8;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int64_t pix = (static_cast<int64_t>(b) * oh + Y) * ow + X;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int dy = 0; dy < k; ++dy)
    for (int dx = 0; dx < k; ++dx) {
      float f[8];
      load8(x + ((static_cast<int64_t>(b) * H + Y * k + dy) * W + X * k + dx) * x_rs + c, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  float inv = 1.f / (k * k);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= inv;
  store8(out + pix * out_rs + c, acc);
}

// nn.AvgPool2d(k, k), floor mode, for k = 2, 4, 8 and 16 of the same map in ONE pass over it: the four pooling branches of
// PyramidLayer (points_sample.py:61-75,115-121).  A CTA owns a 16 x 16 pixel block: 2 x 2 sums from global memory (fp32, kept in
// shared memory), then 4 x 4, 8 x 8 and 16 x 16 sums as a tree over them.  A cell of size k at (Y, X) exists iff Y < H/k and
// X < W/k, and then all of its sub-cells exist, so edge blocks need no special case beyond that test.
// shared memory: (64 + 16 + 4) x C floats.
__global__ void __launch_bounds__(256)
gwd_avgpool_pyramid_kernel(const bf16* __restrict__ x, int64_t x_rs, int H, int W, bf16* __restrict__ o2, bf16* __restrict__ o4,
                           bf16* __restrict__ o8, bf16* __restrict__ o16, int C) {
  extern __shared__ __align__(16) float pyr_s[];
  float* s2 = pyr_s;              // [8][8][C]
  float* s4 = s2 + 64 * C;        // [4][4][C]
  float* s8 = s4 + 16 * C;        // [2][2][C]
  const int cv = C / 8;
  const int bx = blockIdx.x, by = blockIdx.y, b = blockIdx.z;
  const bf16* xb = x + static_cast<int64_t>(b) * H * W * x_rs;
  // ---- 2 x 2 ----
  {
    const int h2 = H / 2, w2 = W / 2;
    for (int it = threadIdx.x; it < 64 * cv; it += blockDim.x) {
      const int cell = it / cv, c = (it - cell * cv) * 8;
      const int Y = 8 * by + (cell >> 3), X = 8 * bx + (cell & 7);
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (Y < h2 && X < w2) {
        const bf16* p0 = xb + (static_cast<int64_t>(2 * Y) * W + 2 * X) * x_rs + c;
        float f0[8], f1[8], f2[8], f3[8];
        load8(p0, f0); load8(p0 + x_rs, f1); load8(p0 + W * x_rs, f2); load8(p0 + (W + 1) * x_rs, f3);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = ((f0[i] + f1[i]) + f2[i]) + f3[i];      // the order of the k = 2 loop of gwd_avgpool_kernel
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[i] * 0.25f;
        store8(o2 + ((static_cast<int64_t>(b) * h2 + Y) * w2 + X) * C + c, o);
      }
      float4* d = reinterpret_cast<float4*>(s2 + cell * C + c);
      d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  __syncthreads();
  // ---- 4 x 4, 8 x 8, 16 x 16: each level sums the 2 x 2 cells of the level below ----
  const float* src = s2;
  float* dst = s4;
  bf16* outs[3] = {o4, o8, o16};
#pragma unroll
  for (int lv = 0; lv < 3; ++lv) {
    const int k = 4 << lv, side = 4 >> lv;      // cells per block side at this level: 4, 2, 1
    const int hk = H / k, wk = W / k;
    const float inv = 1.f / (k * k);
    for (int it = threadIdx.x; it < side * side * cv; it += blockDim.x) {
      const int cell = it / cv, c = (it - cell * cv) * 8;
      const int cy = cell / side, cx = cell - cy * side;
      const int Y = side * by + cy, X = side * bx + cx;
      const float* q = src + ((2 * cy) * (2 * side) + 2 * cx) * C + c;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = (q[i] + q[C + i]) + (q[2 * side * C + i] + q[(2 * side + 1) * C + i]);
      if (lv < 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[cell * C + c + i] = acc[i];
      }
      if (Y < hk && X < wk) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[i] * inv;
        store8(outs[lv] + ((static_cast<int64_t>(b) * hk + Y) * wk + X) * C + c, o);
      }
    }
    __syncthreads();
    src = dst;
    dst = s8;
  }
}

// F.interpolate(mode='bilinear', align_corners=True) of FOUR maps into four adjacent channel slices of one concat buffer
// (the branches of PyramidLayer, points_sample.py:115-121): a pixel's 4 x C channels leave as one contiguous run.
struct Up4 {
  const bf16* src[4];
  int h[4], w[4];
};
__global__ void __launch_bounds__(256)
gwd_bilinear_ac4_kernel(const Up4 u, bf16* __restrict__ out, int64_t out_rs, int H, int W, int C) {
  // four lanes per (pixel, branch): the source coordinates and weights are computed once and shared by the lanes' C / 32 vectors each
  // (this kernel is bound by instruction issue: 89 instructions per 16-byte vector at HBM speed is the budget on this part)
  const int cv = C / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int pb = t >> 2, q = t & 3;
  if (pb >= W * 4) return;
  const int X = pb >> 2, j = pb & 3;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int h = u.h[j], w = u.w[j];
  const float ry = (H > 1) ? static_cast<float>(h - 1) / (H - 1) : 0.f;
  const float rx = (W > 1) ? static_cast<float>(w - 1) / (W - 1) : 0.f;
  const float fy = ry * Y, fx = rx * X;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const bf16* base = u.src[j] + static_cast<int64_t>(b) * h * w * C;
  const bf16* p00 = base + (static_cast<int64_t>(y0) * w + x0) * C;
  const bf16* p01 = base + (static_cast<int64_t>(y0) * w + x1) * C;
  const bf16* p10 = base + (static_cast<int64_t>(y1) * w + x0) * C;
  const bf16* p11 = base + (static_cast<int64_t>(y1) * w + x1) * C;
  bf16* dst = out + ((static_cast<int64_t>(b) * H + Y) * W + X) * out_rs + j * C;
  for (int v = q; v < cv; v += 4) {
    const int c = v * 8;
    float f00[8], f01[8], f10[8], f11[8], o[8];
    load8(p00 + c, f00);
    load8(p01 + c, f01);
    load8(p10 + c, f10);
    load8(p11 + c, f11);
#pragma unroll
    for (int i = 0; i < 8; ++i)     // the expression of gwd_bilinear_ac_kernel, bit for bit
      o[i] = (1.f - ly) * ((1.f - lx) * f00[i] + lx * f01[i]) + ly * ((1.f - lx) * f10[i] + lx * f11[i]);
    store8(dst + c, o);
  }
}

// F.interpolate(mode='bilinear', align_corners=True)
__global__ void gwd_bilinear_ac_kernel(const bf16* x, int64_t x_rs, int B, int h, int w, bf16* out, int64_t out_rs, int H,
                                       int W, int C) {
  const int cv = C / 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * cv) return;
  const int X = t / cv, c = (t - X * cv) * 8;
  const int Y = blockIdx.y, b = blockIdx.z;
  const int64_t pix = (static_cast<int64_t>(b) * H + Y) * W + X;
  float ry = (H > 1) ? static_cast<float>(h - 1) / (H - 1) : 0.f;
  float rx = (W > 1) ? static_cast<float>(w - 1) / (W - 1) : 0.f;
  float fy = ry * Y, fx = rx * X;
  int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  float ly = fy - y0, lx = fx - x0;
  const bf16* base = x + static_cast<int64_t>(b) * h * w * x_rs + c;
  float f00[8], f01[8], f10[8], f11[8], o[8];
  load8(base + (static_cast<int64_t>(y0) * w + x0) * x_rs, f00);
  load8(base + (static_cast<int64_t>(y0) * w + x1) * x_rs, f01);
  load8(base + (static_cast<int64_t>(y1) * w + x0) * x_rs, f10);
  load8(base + (static_cast<int64_t>(y1) * w + x1) * x_rs, f11);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    o[i] = (1.f - ly) * ((1.f - lx) * f00[i] + lx * f01[i]) + ly * ((1.f - lx) * f10[i] + lx * f11[i]);
  store8(out + pix * out_rs + c, o);
}

// ------------------------------------------------------------------------------------------------
// point sampling  (F.grid_sample, align_corners=False, zero padding)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float unnorm(float g, int size) { return ((g + 1.f) * size - 1.f) * 0.5f; }

// bilinear sample of a bf16 map (+ an fp32 [H,W,C] table, e.g. the sine position code) at K points per image:
// out fp32 [B,K,C]
__global__ void gwd_sample_bilinear_kernel(const bf16* x, int64_t x_rs, int x_coff, const float* table, int64_t table_bs, int B,
                                           int H, int W, int C, const float* coords, int K, float* out) {
  int64_t total = static_cast<int64_t>(B) * K * C;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int c = idx % C;
    int k = (idx / C) % K;
    int b = idx / (static_cast<int64_t>(C) * K);
    float gx = coords[(static_cast<int64_t>(b) * K + k) * 2], gy = coords[(static_cast<int64_t>(b) * K + k) * 2 + 1];
    float fx = unnorm(gx, W), fy = unnorm(gy, H);
    int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
    float lx = fx - x0, ly = fy - y0;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        int xx = x0 + dx, yy = y0 + dy;
        if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
        float wgt = (dx ? lx : 1.f - lx) * (dy ? ly : 1.f - ly);
        float v = 0.f;
        if (x) v += __bfloat162float(x[((static_cast<int64_t>(b) * H + yy) * W + xx) * x_rs + x_coff + c]);
        if (table) v += table[b * table_bs + (static_cast<int64_t>(yy) * W + xx) * C + c];
        acc = fmaf(wgt, v, acc);
      }
    out[idx] = acc;
  }
}

// bilinear sample of a single-channel fp32 map [B,H,W] at K points: out [B,K]
__global__ void gwd_sample_scalar_kernel(const float* x, int B, int H, int W, const float* coords, int K, float* out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  int b = idx / K;
  float fx = unnorm(coords[idx * 2], W), fy = unnorm(coords[idx * 2 + 1], H);
  int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
  float lx = fx - x0, ly = fy - y0, acc = 0.f;
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      int xx = x0 + dx, yy = y0 + dy;
      if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
      acc += (dx ? lx : 1.f - lx) * (dy ? ly : 1.f - ly) * x[(static_cast<int64_t>(b) * H + yy) * W + xx];
    }
  out[idx] = acc;
}

// reference tokens of the 1/32 line-window attention (multiscale_transformerr.py:676-701): nearest sample of the
// windowed (LayerNorm'ed, padded, shifted) feature map + nearest sample of the shifted position table at R points.
// win: [B*nW*N, C] (window layout produced by gwd_window_gather), pos: fp32 [H,W,C] (un-shifted), out bf16 [B,R,C]
__global__ void gwd_line_ref_gather_kernel(const bf16* win, int64_t win_rs, const float* pos, int64_t pos_bs, const float* coords,
                                           int R, bf16* out, int64_t out_rs, WinGeom gm, int C) {
  gwd_pdl_trigger();   // a programmatically launched dependent may start its prologue (gwd_common.cuh, PDL)
  int64_t wid = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wid >= static_cast<int64_t>(gm.B) * R) return;
  int b = wid / R;
  float gx = coords[wid * 2], gy = coords[wid * 2 + 1];
  if (gm.shift > 0) {  // roll the coordinates with the feature map, reflecting what crosses -1 (:680-684)
    gx -= (static_cast<float>(gm.shift) / (gm.Wp - 1)) * 2.f;
    gy -= (static_cast<float>(gm.shift) / (gm.Hp - 1)) * 2.f;
    if (gx < -1.f) gx = -1.f - (1.f + gx);
    if (gy < -1.f) gy = -1.f - (1.f + gy);
  }
  // feature sample on the padded, shifted map
  int ix = static_cast<int>(nearbyintf(unnorm(gx, gm.Wp))), iy = static_cast<int>(nearbyintf(unnorm(gy, gm.Hp)));
  bool fvalid = ix >= 0 && ix < gm.Wp && iy >= 0 && iy < gm.Hp;
  int nWx = gm.Wp / gm.ws;
  int64_t wrow = 0;
  if (fvalid)
    wrow = ((static_cast<int64_t>(b) * (gm.Hp / gm.ws) + iy / gm.ws) * nWx + ix / gm.ws) * (gm.ws * gm.ws) +
           (iy % gm.ws) * gm.ws + ix % gm.ws;
  // position sample on the un-padded, shifted table
  int px = static_cast<int>(nearbyintf(unnorm(gx, gm.W))), py = static_cast<int>(nearbyintf(unnorm(gy, gm.H)));
  bool pvalid = px >= 0 && px < gm.W && py >= 0 && py < gm.H;
  int sy = pvalid ? (py + gm.shift) % gm.H : 0, sx = pvalid ? (px + gm.shift) % gm.W : 0;
  for (int c = lane * 8; c < C; c += 256) {
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (fvalid) load8(win + wrow * win_rs + c, f);
    if (pvalid) {
      const float* pp = pos + b * pos_bs + (static_cast<int64_t>(sy) * gm.W + sx) * C + c;
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += pp[i];
    }
    store8(out + wid * out_rs + c, f);
  }
}

// adjoint of the FEATURE part of gwd_line_ref_gather (the position table carries no gradient, and grid_sample(nearest) has
// none w.r.t. the coordinates): d_win[row(b, r)] += d_ref[b, r].  One CTA per image walks its R points in order (several
// points may hit the same token), a thread owns 8 channels: deterministic, no atomics.
__global__ void gwd_line_ref_scatter_kernel(const bf16* d_ref, int64_t dref_rs, const float* coords, int R, bf16* d_win,
                                            int64_t win_rs, WinGeom gm, int C) {
  const int b = blockIdx.x;
  const int nWx = gm.Wp / gm.ws;
  for (int r = 0; r < R; ++r) {
    const int64_t wid = static_cast<int64_t>(b) * R + r;
    float gx = coords[wid * 2], gy = coords[wid * 2 + 1];
    if (gm.shift > 0) {
      gx -= (static_cast<float>(gm.shift) / (gm.Wp - 1)) * 2.f;
      gy -= (static_cast<float>(gm.shift) / (gm.Hp - 1)) * 2.f;
      if (gx < -1.f) gx = -1.f - (1.f + gx);
      if (gy < -1.f) gy = -1.f - (1.f + gy);
    }
    const int ix = static_cast<int>(nearbyintf(unnorm(gx, gm.Wp))), iy = static_cast<int>(nearbyintf(unnorm(gy, gm.Hp)));
    if (!(ix >= 0 && ix < gm.Wp && iy >= 0 && iy < gm.Hp)) continue;      // uniform over the CTA
    const int64_t wrow = ((static_cast<int64_t>(b) * (gm.Hp / gm.ws) + iy / gm.ws) * nWx + ix / gm.ws) * (gm.ws * gm.ws) +
                         (iy % gm.ws) * gm.ws + ix % gm.ws;
    for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
      float f[8], g[8];
      load8(d_win + wrow * win_rs + c, f);
      load8(d_ref + wid * dref_rs + c, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
      store8(d_win + wrow * win_rs + c, f);
    }
  }
}

// depth[p] = sum_k softmax_k(logits[p, :K]) * anchor[b, k]        (points_sample.py:277-279)
__global__ void gwd_anchor_mix_kernel(const bf16* logits, int64_t l_rs, const float* anchor, int B, int64_t HW, int K,
                                      float* out) {
  int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (pix >= B * HW) return;
  int b = pix / HW;
  const bf16* row = logits + pix * l_rs;
  const float* an = anchor + static_cast<int64_t>(b) * K;
  float mx = -INFINITY;
  for (int k = 0; k < K; k += 8) {
    float f[8];
    load8(row + k, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (k + i < K) mx = fmaxf(mx, f[i]);
  }
  float sum = 0.f, acc = 0.f;
  for (int k = 0; k < K; k += 8) {
    float f[8];
    load8(row + k, f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (k + i < K) {
        float e = __expf(f[i] - mx);
        sum += e;
        acc = fmaf(e, an[k + i], acc);
      }
  }
  out[pix] = acc / sum;
}

// image NCHW fp32 -> NHWC bf16 with channel padding (backbone stem input)  and  NHWC bf16 -> NCHW fp32 (outputs)
__global__ void gwd_nchw_to_nhwc_kernel(const float* x, int B, int C, int64_t HW, bf16* out, int Cp) {
  int64_t total = B * HW * Cp;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int c = idx % Cp;
    int64_t p = (idx / Cp) % HW;
    int b = idx / (Cp * HW);
    out[idx] = __float2bfloat16(c < C ? x[(static_cast<int64_t>(b) * C + c) * HW + p] : 0.f);
  }
}

// number of 8-channel vector slots a lane holds for a C-channel row
inline int slots_for(int C) {
  int need = C >> 3;
  int G = need <= 8 ? 8 : (need <= 16 ? 16 : 32);
  return (need + G - 1) / G;
}
#define GWD_ROW_DISPATCH(KERNEL, C, GRID, STREAM, ...)                                   \
  do {                                                                                  \
    int nv__ = slots_for(C);                                                            \
    /* every kernel dispatched here starts with gwd_pdl_wait(): programmatic launch (gwd_common.cuh) */ \
    if (nv__ <= 1) GWD_CUDA(gwd_launch(KERNEL<1>, dim3(GRID), dim3(256), 0, STREAM, 1, __VA_ARGS__));      \
    else if (nv__ <= 2) GWD_CUDA(gwd_launch(KERNEL<2>, dim3(GRID), dim3(256), 0, STREAM, 1, __VA_ARGS__)); \
    else if (nv__ <= 4) GWD_CUDA(gwd_launch(KERNEL<4>, dim3(GRID), dim3(256), 0, STREAM, 1, __VA_ARGS__)); \
    else GWD_CUDA(gwd_launch(KERNEL<8>, dim3(GRID), dim3(256), 0, STREAM, 1, __VA_ARGS__));                \
  } while (0)

int grid_for(int64_t total, int threads) {
  int64_t blocks = gwd_ceil_div(total, threads);
  int64_t cap = static_cast<int64_t>(gwd_num_sms()) * 16;
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}


// ------------------------------------------------------------------------------------------------
// Input builder: B uint8 HWC images (possibly of different sizes) -> the normalised, zero-padded fp32 NCHW batch and its
// padding mask in one pass: ToTensor (/255), Normalize ((v - mean) / std) and nested_tensor_from_tensor_list of
// src/datasets/coco.py:77-78, src/datasets/transforms_depth.py:618-637, src/util/misc.py:291-313.  IEEE division in the
// reference's operation order, so the result is bit-identical to torchvision's.
// ------------------------------------------------------------------------------------------------
struct ImgNorm { float mean[3], std[3]; };

__global__ void __launch_bounds__(256)
gwd_images_to_batch_kernel(const int64_t* __restrict__ table, int B, int H, int W, ImgNorm nm, float* __restrict__ out,
                           uint8_t* __restrict__ mask) {
  const int64_t HW = static_cast<int64_t>(H) * W, total = HW * B;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / HW);
    const int64_t p = i - b * HW;
    const int y = static_cast<int>(p / W), x = static_cast<int>(p - static_cast<int64_t>(y) * W);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(table[3 * b]);
    const int h = static_cast<int>(table[3 * b + 1]), w = static_cast<int>(table[3 * b + 2]);
    const bool in = y < h && x < w;
    float v[3] = {0.f, 0.f, 0.f};
    if (in) {
      const uint8_t* px = src + (static_cast<int64_t>(y) * w + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(px[c]), 255.f), nm.mean[c]), nm.std[c]);
    }
    float* o = out + static_cast<int64_t>(b) * 3 * HW + p;
    o[0] = v[0]; o[HW] = v[1]; o[2 * HW] = v[2];
    if (mask != nullptr) mask[i] = in ? 0 : 1;
  }
}

}  // namespace

#define GWD_STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)
#define GWD_ALIGN8(v) ((v) % 8 == 0)

extern "C" int gwd_layernorm(const void* x, int64_t x_rs, const void* res, int64_t res_rs, const float* gamma,
                             const float* beta, float eps, int32_t act, void* out, int64_t out_rs, int64_t rows, int32_t C,
                             int32_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && rows > 0, "gwd_layernorm: null pointer / empty");
  GWD_CHECK_ARG(GWD_ALIGN8(C) && C <= 2048 && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs) && GWD_ALIGN8(res_rs),
                "gwd_layernorm: C and strides must be multiples of 8, C <= 2048");
  const int nv64 = C % 64 == 0 ? C / 64 : 0;
  if (nv64 == 3 || nv64 == 5 || nv64 == 6 || nv64 == 7) {
    const unsigned grid = static_cast<unsigned>(gwd_ceil_div(gwd_ceil_div(rows, 4) * 32, 256));
#define GWD_LN_G8(NV)                                                                                                  \
  GWD_CUDA(gwd_launch(gwd_layernorm_g8_kernel<NV>, dim3(grid), dim3(256), 0, stream, 1, static_cast<const bf16*>(x), x_rs,         \
                      static_cast<const bf16*>(res), res_rs, gamma, beta, eps, act, static_cast<bf16*>(out), out_rs, rows, C, n))
    if (nv64 == 3) GWD_LN_G8(3);
    else if (nv64 == 5) GWD_LN_G8(5);
    else if (nv64 == 6) GWD_LN_G8(6);
    else GWD_LN_G8(7);
#undef GWD_LN_G8
  } else {
    GWD_ROW_DISPATCH(gwd_layernorm_kernel, C, static_cast<unsigned>(gwd_ceil_div(gwd_ceil_div(rows, rows_per_warp(C)) * 32, 256)), stream,
                     static_cast<const bf16*>(x), x_rs, static_cast<const bf16*>(res), res_rs, gamma, beta, eps, act,
                     static_cast<bf16*>(out), out_rs, rows, C, n);
  }
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_add_rows(const void* x, int64_t x_rs, const void* addend, int64_t a_rs, int64_t period, void* out,
                            int64_t out_rs, int64_t rows, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && addend && out && rows > 0 && period > 0, "gwd_add_rows: bad argument");
  GWD_CHECK_ARG(GWD_ALIGN8(C) && C <= 2048 && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs) && GWD_ALIGN8(a_rs),
                "gwd_add_rows: alignment");
  GWD_ROW_DISPATCH(gwd_add_rows_kernel, C, static_cast<unsigned>(gwd_ceil_div(gwd_ceil_div(rows, rows_per_warp(C)) * 32, 256)), stream,
                   static_cast<const bf16*>(x), x_rs, static_cast<const bf16*>(addend), a_rs, period, static_cast<bf16*>(out), out_rs,
      rows, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

static int make_geom(WinGeom& gm, int B, int H, int W, int ws, int shift) {
  gm.B = B; gm.H = H; gm.W = W; gm.ws = ws; gm.shift = shift;
  gm.Hp = static_cast<int>(gwd_ceil_div(H, ws)) * ws;
  gm.Wp = static_cast<int>(gwd_ceil_div(W, ws)) * ws;
  return 0;
}

extern "C" int gwd_window_gather(const void* x, int64_t x_rs, const float* gamma, const float* beta, float eps, void* out,
                                 int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C,
                                 int32_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && ws > 0 && shift >= 0 && shift < ws, "gwd_window_gather: bad argument");
  GWD_CHECK_ARG(GWD_ALIGN8(C) && C <= 2048 && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs), "gwd_window_gather: alignment");
  WinGeom gm;
  make_geom(gm, B, H, W, ws, shift);
  GWD_CHECK_ARG(B > 0 && B <= 65535 && gm.Hp <= 65535, "gwd_window_gather: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(gwd_ceil_div(gm.Wp, rows_per_warp(C)) * 32, 256)), gm.Hp, B);
  GWD_ROW_DISPATCH(gwd_window_gather_kernel, C, grid, stream,
                   static_cast<const bf16*>(x), x_rs, gamma, beta, eps, static_cast<bf16*>(out), out_rs, gm, C, n);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_window_merge(const void* win, int64_t win_rs, const void* shortcut, int64_t sc_rs, void* out,
                                int64_t out_rs, const float* gamma, const float* beta, float eps, void* out_ln,
                                int64_t ln_rs, int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C,
                                int32_t n, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(win && shortcut && out, "gwd_window_merge: null pointer");
  GWD_CHECK_ARG(GWD_ALIGN8(C) && C <= 2048 && GWD_ALIGN8(win_rs) && GWD_ALIGN8(sc_rs) && GWD_ALIGN8(out_rs) &&
                    GWD_ALIGN8(ln_rs),
                "gwd_window_merge: alignment");
  WinGeom gm;
  make_geom(gm, B, H, W, ws, shift);
  GWD_CHECK_ARG(B > 0 && B <= 65535 && H <= 65535, "gwd_window_merge: bad extents");
  const dim3 grid(static_cast<unsigned>(gwd_ceil_div(gwd_ceil_div(W, rows_per_warp(C)) * 32, 256)), H, B);
  GWD_ROW_DISPATCH(gwd_window_merge_kernel, C, grid, stream,
                   static_cast<const bf16*>(win), win_rs, static_cast<const bf16*>(shortcut), sc_rs, static_cast<bf16*>(out), out_rs,
      gamma, beta, eps, static_cast<bf16*>(out_ln), ln_rs, gm, C, n);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_upsample_nearest(const void* x, int64_t x_rs, int32_t B, int32_t h, int32_t w, void* out, int64_t out_rs,
                                    int32_t H, int32_t W, int32_t C, const void* add, int64_t add_rs, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && GWD_ALIGN8(C) && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs) && GWD_ALIGN8(add_rs),
                "gwd_upsample_nearest: bad argument");
  int64_t total = static_cast<int64_t>(B) * H * W * (C / 8);
  GWD_CHECK_ARG(total > 0 && H <= 65535 && B <= 65535, "gwd_upsample_nearest: bad extents");
  GWD_CUDA(gwd_launch(gwd_upsample_nearest_kernel, dim3(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W) * (C / 8), 256)), H, B),
                      dim3(256), 0, stream, 1, static_cast<const bf16*>(x), x_rs, B, h, w, static_cast<bf16*>(out), out_rs, H, W, C,
                      static_cast<const bf16*>(add), add_rs));
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_avgpool(const void* x, int64_t x_rs, int32_t B, int32_t H, int32_t W, int32_t k, void* out,
                           int64_t out_rs, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && k > 0 && H >= k && W >= k && GWD_ALIGN8(C) && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs),
                "gwd_avgpool: bad argument");
  int64_t total = static_cast<int64_t>(B) * (H / k) * (W / k) * (C / 8);
  GWD_CHECK_ARG(total > 0 && H / k <= 65535 && B <= 65535, "gwd_avgpool: bad extents");
  GWD_CUDA(gwd_launch(gwd_avgpool_kernel, dim3(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W / k) * (C / 8), 256)), H / k, B),
                      dim3(256), 0, stream, 1, static_cast<const bf16*>(x), x_rs, B, H, W, k, static_cast<bf16*>(out), out_rs, C));
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_avgpool_pyramid(const void* x, int64_t x_rs, int32_t B, int32_t H, int32_t W, void* o2, void* o4, void* o8,
                                   void* o16, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && o2 && o4 && o8 && o16 && H >= 16 && W >= 16 && GWD_ALIGN8(C) && GWD_ALIGN8(x_rs), "gwd_avgpool_pyramid: bad argument");
  const size_t smem = static_cast<size_t>(84) * C * sizeof(float);
  GWD_CHECK_ARG(C > 0 && smem <= 200 * 1024 && B > 0 && B <= 65535 && (H + 15) / 16 <= 65535, "gwd_avgpool_pyramid: bad extents");
  static size_t attr = 0;
  if (smem > attr) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_avgpool_pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  gwd_avgpool_pyramid_kernel<<<dim3((W + 15) / 16, (H + 15) / 16, B), 256, smem, stream>>>(
      static_cast<const bf16*>(x), x_rs, H, W, static_cast<bf16*>(o2), static_cast<bf16*>(o4), static_cast<bf16*>(o8),
      static_cast<bf16*>(o16), C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_bilinear_up4(const void* x0, const void* x1, const void* x2, const void* x3, const int32_t* hw, int32_t B,
                                void* out, int64_t out_rs, int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x0 && x1 && x2 && x3 && hw && out && GWD_ALIGN8(C) && GWD_ALIGN8(out_rs), "gwd_bilinear_up4: bad argument");
  GWD_CHECK_ARG(B > 0 && H > 0 && W > 0 && H <= 65535 && B <= 65535, "gwd_bilinear_up4: bad extents");
  Up4 u;
  const void* xs[4] = {x0, x1, x2, x3};
  for (int j = 0; j < 4; ++j) {
    u.src[j] = static_cast<const bf16*>(xs[j]);
    u.h[j] = hw[2 * j];
    u.w[j] = hw[2 * j + 1];
    GWD_CHECK_ARG(u.h[j] > 0 && u.w[j] > 0, "gwd_bilinear_up4: empty source");
  }
  gwd_bilinear_ac4_kernel<<<dim3(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W) * 16, 256)), H, B), 256, 0, stream>>>(
      u, static_cast<bf16*>(out), out_rs, H, W, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_bilinear_up(const void* x, int64_t x_rs, int32_t B, int32_t h, int32_t w, void* out, int64_t out_rs,
                               int32_t H, int32_t W, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && GWD_ALIGN8(C) && GWD_ALIGN8(x_rs) && GWD_ALIGN8(out_rs), "gwd_bilinear_up: bad argument");
  int64_t total = static_cast<int64_t>(B) * H * W * (C / 8);
  GWD_CHECK_ARG(total > 0 && H <= 65535 && B <= 65535, "gwd_bilinear_up: bad extents");
  gwd_bilinear_ac_kernel<<<dim3(static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(W) * (C / 8), 256)), H, B), 256, 0, stream>>>(static_cast<const bf16*>(x), x_rs, B, h, w,
                                                                   static_cast<bf16*>(out), out_rs, H, W, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_sample_bilinear(const void* x, int64_t x_rs, int32_t x_coff, const float* table, int64_t table_bstride,
                                   int32_t B, int32_t H, int32_t W, int32_t C, const float* coords, int32_t K, float* out,
                                   void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG((x || table) && coords && out, "gwd_sample_bilinear: null pointer");
  int64_t total = static_cast<int64_t>(B) * K * C;
  gwd_sample_bilinear_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const bf16*>(x), x_rs, x_coff, table,
                                                                       table_bstride, B, H, W, C, coords, K, out);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_sample_scalar(const float* x, int32_t B, int32_t H, int32_t W, const float* coords, int32_t K,
                                 float* out, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && coords && out, "gwd_sample_scalar: null pointer");
  gwd_sample_scalar_kernel<<<static_cast<unsigned>(gwd_ceil_div(static_cast<int64_t>(B) * K, 128)), 128, 0, stream>>>(
      x, B, H, W, coords, K, out);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_line_ref_gather(const void* win, int64_t win_rs, const float* pos, int64_t pos_bstride, const float* coords,
                                   int32_t R, void* out, int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift,
                                   int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(win && pos && coords && out && GWD_ALIGN8(C) && GWD_ALIGN8(win_rs) && GWD_ALIGN8(out_rs),
                "gwd_line_ref_gather: bad argument");
  WinGeom gm;
  make_geom(gm, B, H, W, ws, shift);
  int64_t warps = static_cast<int64_t>(B) * R;
  gwd_line_ref_gather_kernel<<<static_cast<unsigned>(gwd_ceil_div(warps * 32, 128)), 128, 0, stream>>>(
      static_cast<const bf16*>(win), win_rs, pos, pos_bstride, coords, R, static_cast<bf16*>(out), out_rs, gm, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_line_ref_scatter(const void* d_ref, int64_t dref_rs, const float* coords, int32_t R, void* d_win, int64_t win_rs,
                                    int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(d_ref && coords && d_win && GWD_ALIGN8(C) && GWD_ALIGN8(win_rs) && GWD_ALIGN8(dref_rs),
                "gwd_line_ref_scatter: bad argument");
  WinGeom gm;
  make_geom(gm, B, H, W, ws, shift);
  gwd_line_ref_scatter_kernel<<<static_cast<unsigned>(B), 64, 0, stream>>>(static_cast<const bf16*>(d_ref), dref_rs, coords, R,
                                                                          static_cast<bf16*>(d_win), win_rs, gm, C);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_anchor_mix(const void* logits, int64_t l_rs, const float* anchor, int32_t B, int64_t HW, int32_t K,
                              float* out, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(logits && anchor && out && GWD_ALIGN8(l_rs), "gwd_anchor_mix: bad argument");
  gwd_anchor_mix_kernel<<<static_cast<unsigned>(gwd_ceil_div(B * HW, 256)), 256, 0, stream>>>(
      static_cast<const bf16*>(logits), l_rs, anchor, B, HW, K, out);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_nchw_to_nhwc(const float* x, int32_t B, int32_t C, int64_t HW, void* out, int32_t Cp, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(x && out && Cp >= C, "gwd_nchw_to_nhwc: bad argument");
  gwd_nchw_to_nhwc_kernel<<<grid_for(B * HW * Cp, 256), 256, 0, stream>>>(x, B, C, HW, static_cast<bf16*>(out), Cp);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_images_to_batch(const int64_t* table, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3,
                                   float* out, uint8_t* mask, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(table && mean3 && std3 && out && B > 0 && H > 0 && W > 0, "gwd_images_to_batch: bad argument");
  ImgNorm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
  gwd_images_to_batch_kernel<<<grid_for(static_cast<int64_t>(B) * H * W, 256), 256, 0, stream>>>(table, B, H, W, nm, out, mask);
  GWD_LAUNCHED();
  return GWD_OK;
}
