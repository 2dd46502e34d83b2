// gwd_data.cu -- the pixel side of the training data path on the GPU: what the reference does to
// PIL images in its DataLoader workers (src/datasets/transforms_depth.py hflip :206, vflip :234, resize :316, crop :59,
// ColorJitter :551; composed by src/datasets/coco.py:74-103), bit for bit:
//   * BILINEAR resize of the 8-bit RGB image = Pillow's ImagingResample (separable antialiased triangle filter, 22-bit fixed-point
//     coefficients, 8-bit intermediate after the horizontal pass); the coefficient tables are computed on the HOST in double, in
//     Pillow's operation order (gwd_pil_bilinear_coeffs), the passes are integer arithmetic on the device;
//   * NEAREST resize / flip / crop of the auxiliary maps (depth, segmentation) = one index gather (gwd_gather2d) with Pillow's
//     ImagingScaleAffine index rule (gwd_pil_nearest_index: running double sum);
//   * ColorJitter = ImageEnhance.Brightness / Contrast / Color (ImagingBlend, single precision, truncation) and torchvision's hue
//     shift through Pillow's RGB <-> HSV conversions, applied per pixel in registers in the drawn order (gwd_jitter_u8).
// All kernels are bandwidth trivial (a 480 x 640 RGB image is 0.9 MB); they exist so that decoded uint8 images go to the GPU once
// and the fp32 batch is built there (gwd_images_to_batch), instead of 2 DataLoader workers feeding 8 GPUs.
#include <math.h>
#include "gwd_common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

// one pass of ImagingResample along x (AXIS 1: dst [H, out, C]) or y (AXIS 0: dst [out, W, C]); flip mirrors the SOURCE index along
// the resampled axis (the reference flips first, then resizes)
template <int AXIS>
__global__ void gwd_resample_u8_kernel(const uint8_t* __restrict__ src, int64_t src_rs, int H, int W, int C, uint8_t* __restrict__ dst,
                                       int out_size, const int32_t* __restrict__ xmin, const int32_t* __restrict__ cnt,
                                       const int32_t* __restrict__ kk, int ksize, int flip) {
  const int64_t total = AXIS == 1 ? static_cast<int64_t>(H) * out_size * C : static_cast<int64_t>(out_size) * W * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int c, o, other;
    if (AXIS == 1) {
      c = static_cast<int>(i % C);
      o = static_cast<int>((i / C) % out_size);
      other = static_cast<int>(i / (static_cast<int64_t>(C) * out_size));      // y
    } else {
      const int wc = W * C;
      other = static_cast<int>(i % wc);                                        // x * C + c
      o = static_cast<int>(i / wc);
      c = 0;
    }
    const int x0 = xmin[o], n = cnt[o];
    const int32_t* k = kk + static_cast<int64_t>(o) * ksize;
    int32_t acc = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < n; ++x) {
      const int s = x0 + x;
      uint8_t v;
      if (AXIS == 1) {
        const int sx = flip ? W - 1 - s : s;
        v = src[other * src_rs + static_cast<int64_t>(sx) * C + c];
      } else {
        const int sy = flip ? H - 1 - s : s;
        v = src[sy * src_rs + other];
      }
      acc += static_cast<int32_t>(v) * k[x];
    }
    acc >>= kPrecisionBits;
    dst[i] = static_cast<uint8_t>(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

// dst[y, x] = src[sy(y), sx(x)], element size ES bytes; iy / ix = source index tables (null = identity), flips mirror the source index
template <typename T>
__global__ void gwd_gather2d_kernel(const T* __restrict__ src, int64_t src_rs, int H, int W, T* __restrict__ dst, int oh, int ow,
                                    const int32_t* __restrict__ iy, const int32_t* __restrict__ ix, int flip_h, int flip_v) {
  const int64_t total = static_cast<int64_t>(oh) * ow;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % ow), y = static_cast<int>(i / ow);
    int sy = iy ? iy[y] : y, sx = ix ? ix[x] : x;
    if (flip_v) sy = H - 1 - sy;
    if (flip_h) sx = W - 1 - sx;
    dst[i] = src[sy * src_rs + sx];
  }
}

struct JitterOps {
  int n;
  int op[4];          // 0 brightness, 1 contrast, 2 saturation, 3 hue
  float factor[4];
  int hue_shift[4];   // uint8(int32(factor * 255)) of a hue op
};

__device__ __forceinline__ int gray_of(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// ImagingBlend(degenerate, image, alpha) for one byte: (UINT8)(in1 + alpha * (in2 - in1)) in single precision, no contraction
__device__ __forceinline__ int blend_u8(int deg, int v, float alpha, bool inside) {
  const float t = __fadd_rn(static_cast<float>(deg), __fmul_rn(alpha, static_cast<float>(v - deg)));
  if (inside) return static_cast<int>(t) & 0xFF;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : static_cast<int>(t));
}

// Pillow Convert.c rgb2hsv_row / hsv2rgb (the literal 2.0 / 4.0 / 6.0 / 255.0 are doubles there: the mixed precision below is theirs)
__device__ __forceinline__ void rgb_to_hsv(int r, int g, int b, int& uh, int& us, int& uv) {
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  uv = maxc;
  if (minc == maxc) { uh = 0; us = 0; return; }
  const float cr = static_cast<float>(maxc - minc);
  const float s = __fdiv_rn(cr, static_cast<float>(maxc));
  const float rc = __fdiv_rn(static_cast<float>(maxc - r), cr), gc = __fdiv_rn(static_cast<float>(maxc - g), cr),
              bc = __fdiv_rn(static_cast<float>(maxc - b), cr);
  float h;
  if (r == maxc) h = __fsub_rn(bc, gc);
  else if (g == maxc) h = static_cast<float>(__dsub_rn(__dadd_rn(2.0, static_cast<double>(rc)), static_cast<double>(bc)));
  else h = static_cast<float>(__dsub_rn(__dadd_rn(4.0, static_cast<double>(gc)), static_cast<double>(rc)));
  h = static_cast<float>(fmod(__dadd_rn(__ddiv_rn(static_cast<double>(h), 6.0), 1.0), 1.0));
  const int ih = static_cast<int>(__dmul_rn(static_cast<double>(h), 255.0)), is = static_cast<int>(__dmul_rn(static_cast<double>(s), 255.0));
  uh = min(max(ih, 0), 255);
  us = min(max(is, 0), 255);
}
__device__ __forceinline__ int round_clip8(float x) {      // C round(): half away from zero
  const int v = static_cast<int>(floor(static_cast<double>(x) + 0.5));
  return min(max(v, 0), 255);
}
__device__ __forceinline__ void hsv_to_rgb(int h, int s, int v, int& r, int& g, int& b) {
  if (s == 0) { r = g = b = v; return; }
  const float fh = __fdiv_rn(__fmul_rn(static_cast<float>(h), 6.0f), 255.0f);
  const int i = static_cast<int>(floorf(fh));
  const float f = __fsub_rn(fh, static_cast<float>(i));
  const float fs = __fdiv_rn(static_cast<float>(s), 255.0f);
  const float vf = static_cast<float>(v);
  const int p = round_clip8(__fmul_rn(vf, __fsub_rn(1.0f, fs)));
  const int q = round_clip8(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, f))));
  const int t = round_clip8(__fmul_rn(vf, __fsub_rn(1.0f, __fmul_rn(fs, __fsub_rn(1.0f, f)))));
  switch (i % 6) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// in place on [npix, 3] uint8.  A contrast op must be the FIRST op of a launch: it blends with int(mean(L) + 0.5) of the image as it
// is at that point, whose L sum arrives in gray_in (written by the previous launch's gray_out).
__global__ void gwd_jitter_u8_kernel(uint8_t* __restrict__ img, int64_t npix, const JitterOps ops,
                                     const unsigned long long* __restrict__ gray_in, unsigned long long* __restrict__ gray_out) {
  int mean = 0;
  if (gray_in != nullptr) mean = static_cast<int>(static_cast<double>(*gray_in) / static_cast<double>(npix) + 0.5);
  unsigned long long local = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < npix; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int r = img[3 * i], g = img[3 * i + 1], b = img[3 * i + 2];
    for (int k = 0; k < ops.n; ++k) {
      const float a = ops.factor[k];
      const bool inside = a >= 0.f && a <= 1.0f;
      switch (ops.op[k]) {
        case 0: r = blend_u8(0, r, a, inside); g = blend_u8(0, g, a, inside); b = blend_u8(0, b, a, inside); break;
        case 1: r = blend_u8(mean, r, a, inside); g = blend_u8(mean, g, a, inside); b = blend_u8(mean, b, a, inside); break;
        case 2: {
          const int l = gray_of(r, g, b);
          r = blend_u8(l, r, a, inside); g = blend_u8(l, g, a, inside); b = blend_u8(l, b, a, inside);
          break;
        }
        default: {
          int h, s, v;
          rgb_to_hsv(r, g, b, h, s, v);
          h = (h + ops.hue_shift[k]) & 0xFF;
          hsv_to_rgb(h, s, v, r, g, b);
          break;
        }
      }
    }
    img[3 * i] = static_cast<uint8_t>(r); img[3 * i + 1] = static_cast<uint8_t>(g); img[3 * i + 2] = static_cast<uint8_t>(b);
    if (gray_out != nullptr) local += static_cast<unsigned long long>(gray_of(r, g, b));
  }
  if (gray_out != nullptr) {     // integer sum: exact and order independent
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(gray_out, local);
  }
}

int grid_of(int64_t total) {
  int64_t blocks = gwd_ceil_div(total, 256);
  const int64_t cap = static_cast<int64_t>(gwd_num_sms()) * 8;
  return static_cast<int>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

#define GWD_STREAM cudaStream_t stream = static_cast<cudaStream_t>(stream_)

// ---------------------------------------------------------------------------------------------- host: Pillow's index arithmetic
extern "C" int gwd_pil_bilinear_ksize(int32_t in_size, int32_t out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  return static_cast<int>(ceil(1.0 * filterscale)) * 2 + 1;
}

extern "C" int gwd_pil_bilinear_coeffs(int32_t in_size, int32_t out_size, int32_t* xmin_out, int32_t* cnt_out, int32_t* kk_out) {
  GWD_CHECK_ARG(in_size > 0 && out_size > 0 && xmin_out && cnt_out && kk_out, "gwd_pil_bilinear_coeffs: bad argument");
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  const double ss = 1.0 / filterscale;
  double* w = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double v = (x + xmin - center + 0.5) * ss;
      if (v < 0.0) v = -v;
      w[x] = v < 1.0 ? 1.0 - v : 0.0;
      ww += w[x];
    }
    int32_t* k = kk_out + static_cast<int64_t>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      double kv = 0.0;
      if (x < xmax) kv = ww != 0.0 ? w[x] / ww : w[x];
      k[x] = kv < 0 ? static_cast<int32_t>(-0.5 + kv * (1 << kPrecisionBits)) : static_cast<int32_t>(0.5 + kv * (1 << kPrecisionBits));
    }
    xmin_out[xx] = xmin;
    cnt_out[xx] = xmax;
  }
  delete[] w;
  return GWD_OK;
}

extern "C" int gwd_pil_nearest_index(int32_t in_size, int32_t out_size, int32_t* idx_out) {
  GWD_CHECK_ARG(in_size > 0 && out_size > 0 && idx_out, "gwd_pil_nearest_index: bad argument");
  const double a = static_cast<double>(in_size) / out_size;
  double xo = 0.0 + a * 0.5;
  for (int x = 0; x < out_size; ++x) {
    int v = xo < 0.0 ? -1 : static_cast<int>(xo);
    if (v > in_size - 1) v = in_size - 1;
    idx_out[x] = v;
    xo += a;
  }
  return GWD_OK;
}

// ---------------------------------------------------------------------------------------------- device entry points
extern "C" int gwd_resample_u8(const void* src, int64_t src_rs, int32_t H, int32_t W, int32_t C, void* dst, int32_t out_size, int32_t axis,
                               const int32_t* xmin, const int32_t* cnt, const int32_t* kk, int32_t ksize, int32_t flip, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(src && dst && xmin && cnt && kk && H > 0 && W > 0 && C > 0 && out_size > 0 && ksize > 0 && (axis == 0 || axis == 1),
                "gwd_resample_u8: bad argument");
  const int64_t total = axis == 1 ? static_cast<int64_t>(H) * out_size * C : static_cast<int64_t>(out_size) * W * C;
  if (axis == 1)
    gwd_resample_u8_kernel<1><<<grid_of(total), 256, 0, stream>>>(static_cast<const uint8_t*>(src), src_rs, H, W, C, static_cast<uint8_t*>(dst),
                                                                  out_size, xmin, cnt, kk, ksize, flip);
  else
    gwd_resample_u8_kernel<0><<<grid_of(total), 256, 0, stream>>>(static_cast<const uint8_t*>(src), src_rs, H, W, C, static_cast<uint8_t*>(dst),
                                                                  out_size, xmin, cnt, kk, ksize, flip);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_gather2d(const void* src, int64_t src_rs, int32_t elem_bytes, int32_t H, int32_t W, void* dst, int32_t oh, int32_t ow,
                            const int32_t* iy, const int32_t* ix, int32_t flip_h, int32_t flip_v, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(src && dst && H > 0 && W > 0 && oh > 0 && ow > 0, "gwd_gather2d: bad argument");
  GWD_CHECK_ARG((iy != nullptr || oh <= H) && (ix != nullptr || ow <= W), "gwd_gather2d: identity index beyond the source");
  const int g = grid_of(static_cast<int64_t>(oh) * ow);
#define GWD_GATHER(T)                                                                                                          \
  gwd_gather2d_kernel<T><<<g, 256, 0, stream>>>(static_cast<const T*>(src), src_rs, H, W, static_cast<T*>(dst), oh, ow, iy, ix, flip_h, \
                                                flip_v)
  switch (elem_bytes) {      // src_rs is in ELEMENTS
    case 1: GWD_GATHER(uint8_t); break;
    case 2: GWD_GATHER(uint16_t); break;
    case 4: GWD_GATHER(uint32_t); break;
    case 8: GWD_GATHER(uint64_t); break;
    default: gwd_set_error("gwd_gather2d: element size %d not supported", elem_bytes); return GWD_ERR_ARG;
  }
#undef GWD_GATHER
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_jitter_u8(void* img, int64_t npix, int32_t n_ops, const int32_t* ops, const float* factors, const void* gray_in,
                             void* gray_out, void* stream_) {
  GWD_STREAM;
  GWD_CHECK_ARG(img && npix > 0 && n_ops >= 0 && n_ops <= 4 && (n_ops == 0 || (ops && factors)), "gwd_jitter_u8: bad argument");
  JitterOps jo;
  jo.n = n_ops;
  for (int k = 0; k < 4; ++k) { jo.op[k] = 0; jo.factor[k] = 1.f; jo.hue_shift[k] = 0; }
  for (int k = 0; k < n_ops; ++k) {
    GWD_CHECK_ARG(ops[k] >= 0 && ops[k] <= 3, "gwd_jitter_u8: unknown op %d", ops[k]);
    GWD_CHECK_ARG(ops[k] != 1 || (k == 0 && gray_in != nullptr), "gwd_jitter_u8: contrast must be the first op of a launch and needs gray_in");
    jo.op[k] = ops[k];
    jo.factor[k] = factors[k];
    if (ops[k] == 3) {
      GWD_CHECK_ARG(factors[k] >= -0.5f && factors[k] <= 0.5f, "gwd_jitter_u8: hue factor outside [-0.5, 0.5]");
      jo.hue_shift[k] = static_cast<int>(static_cast<int32_t>(static_cast<double>(factors[k]) * 255.0)) & 0xFF;   // np.int32(f * 255).astype(uint8)
    }
  }
  if (gray_out != nullptr) GWD_CUDA(cudaMemsetAsync(gray_out, 0, sizeof(unsigned long long), stream));
  gwd_jitter_u8_kernel<<<grid_of(npix), 256, 0, stream>>>(static_cast<uint8_t*>(img), npix, jo, static_cast<const unsigned long long*>(gray_in),
                                                          static_cast<unsigned long long*>(gray_out));
  GWD_LAUNCHED();
  return GWD_OK;
}
