// gwd_backbone.cu -- the stride-2 3x3 convolutions of the ResNet-50 bottlenecks (layer2.0 / layer3.0 / layer4.0 conv2,
// torchvision v1.5 as wrapped by src/models/backbone.py:58-92) as im2col + the tcgen05 GEMM, forward and backward:
//   gwd_im2col3x3_s2 : x bf16 [B,H,W,C] -> col bf16 [B,ho,wo,9C], col[.., (ky*3+kx)*C + c] = x[2oy+ky-1, 2ox+kx-1, c] (zero padding 1)
//   gwd_col2im3x3_s2 : its adjoint, as a GATHER over the input pixels (1, 2 or 4 taps land on a pixel; fp32 sums, no atomics)
// With the filter stored as [N, 9C] the convolution, its data gradient (dcol = dY W) and its weight gradient (dW = dY^T col) are
// plain Linears on gwd_conv_gemm / gwd_linear_wgrad: exactly the stride-2 FLOPs (a stride-1 convolution followed by a
// sub-sampling, or a zero-stuffed gradient, would do 4x the tensor work).  Both kernels are HBM bound: 9C * 2 B written per
// output pixel (im2col), 9C/4 * 2 B read per input pixel (col2im), 16-byte vectors.
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

__global__ void gwd_im2col3x3_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ col, int B, int H, int W, int ho, int wo, int C8) {
  gwd_pdl_trigger();   // a programmatically launched dependent may start its prologue (gwd_common.cuh, PDL)
  const int64_t total = static_cast<int64_t>(B) * ho * wo * 9 * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    int64_t r = i / C8;
    const int tap = r % 9;
    r /= 9;
    const int ox = r % wo, oy = (r / wo) % ho, b = r / (static_cast<int64_t>(wo) * ho);
    const int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((static_cast<int64_t>(b) * H + iy) * W + ix) * C8 + c];
    col[i] = v;
  }
}

__device__ __forceinline__ void acc8(float (&f)[8], const uint4& u) {
  const float2 a = gwd_unpack_bf16x2(u.x), b = gwd_unpack_bf16x2(u.y), c = gwd_unpack_bf16x2(u.z), d = gwd_unpack_bf16x2(u.w);
  f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y; f[4] += c.x; f[5] += c.y; f[6] += d.x; f[7] += d.y;
}

__global__ void gwd_col2im3x3_s2_kernel(const uint4* __restrict__ dcol, const uint4* __restrict__ add, uint4* __restrict__ dx, int B, int H,
                                        int W, int ho, int wo, int C8) {
  const int64_t total = static_cast<int64_t>(B) * H * W * C8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = i % C8;
    const int64_t p = i / C8;
    const int ix = p % W, iy = (p / W) % H, b = p / (static_cast<int64_t>(W) * H);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (add) acc8(f, add[i]);
    for (int ky = (iy + 1) & 1; ky < 3; ky += 2) {        // 2 oy = iy + 1 - ky
      const int oy = (iy + 1 - ky) >> 1;
      if (oy < 0 || oy >= ho) continue;
      for (int kx = (ix + 1) & 1; kx < 3; kx += 2) {
        const int ox = (ix + 1 - kx) >> 1;
        if (ox < 0 || ox >= wo) continue;
        acc8(f, dcol[(((static_cast<int64_t>(b) * ho + oy) * wo + ox) * 9 + ky * 3 + kx) * C8 + c]);
      }
    }
    uint4 o;
    o.x = gwd_pack_bf16x2(f[0], f[1]); o.y = gwd_pack_bf16x2(f[2], f[3]);
    o.z = gwd_pack_bf16x2(f[4], f[5]); o.w = gwd_pack_bf16x2(f[6], f[7]);
    dx[i] = o;
  }
}

unsigned grid_1d(int64_t n, int block) {
  int64_t g = gwd_ceil_div(n, block);
  const int64_t cap = static_cast<int64_t>(gwd_num_sms()) * 16;
  return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int gwd_im2col3x3_s2(const void* x, void* col, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(x && col && C % 8 == 0 && B > 0 && H > 0 && W > 0, "gwd_im2col3x3_s2: bad argument (C must be a multiple of 8)");
  const int ho = (H - 1) / 2 + 1, wo = (W - 1) / 2 + 1;
  gwd_im2col3x3_s2_kernel<<<grid_1d(static_cast<int64_t>(B) * ho * wo * 9 * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(col), B, H, W, ho, wo, C / 8);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_col2im3x3_s2(const void* dcol, const void* add, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(dcol && dx && C % 8 == 0 && B > 0 && H > 0 && W > 0, "gwd_col2im3x3_s2: bad argument (C must be a multiple of 8)");
  const int ho = (H - 1) / 2 + 1, wo = (W - 1) / 2 + 1;
  gwd_col2im3x3_s2_kernel<<<grid_1d(static_cast<int64_t>(B) * H * W * (C / 8), 256), 256, 0, stream>>>(
      static_cast<const uint4*>(dcol), static_cast<const uint4*>(add), static_cast<uint4*>(dx), B, H, W, ho, wo, C / 8);
  GWD_LAUNCHED();
  return GWD_OK;
}
