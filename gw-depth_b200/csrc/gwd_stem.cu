// gwd_stem.cu -- the ResNet stem in one kernel: 7x7 stride-2 convolution (3 -> 64) + folded FrozenBatchNorm + ReLU +
// 3x3 stride-2 max-pool, fp32 NCHW images in, bf16 channels-last C1 map out.
//
// Replaces conv1 / bn1 / relu / maxpool of the torchvision ResNet-50 body the reference wraps in
// src/models/backbone.py:58-92 (IntermediateLayerGetter over resnet50, FrozenBatchNorm2d :19-55).
//
// Why its own kernel: with 3 input channels the convolution is an im2col GEMM with K = 147, far too ragged for the
// TMA/tcgen05 tile kernel (gwd_gemm.cu needs channel counts that are multiples of 16), and the library path wrote the
// 157 MB full-resolution 64-channel map, re-read it for the ReLU and again for the pool (0.9 ms of a 20 ms step).
// Here one CTA produces an 8x16 tile of POOLED pixels: it stages the 39x71x3 input patch as bf16 (zero padded), forms
// the 17x33 convolution pixels the pool needs with warp-level mma.sync m16n8k16 (A fragments are read straight out of
// the patch: for a fixed filter row the 7x3 taps of a pixel are 21 contiguous patch elements), adds the folded BN shift,
// applies the ReLU, parks the bf16 tile in shared memory (XOR-swizzled 16-byte chunks) and max-pools it from there.
// Only the pooled map (1/16 of the bytes) goes to HBM.
#include <stdlib.h>
#include <string.h>
#include "gwd_common.cuh"

namespace {

typedef __nv_bfloat16 bf16;

constexpr int kPH = 8, kPW = 16;                    // pooled pixels per CTA
constexpr int kCH = 2 * kPH + 1, kCW = 2 * kPW + 1;  // convolution pixels the pool needs: 17 x 33
constexpr int kNPix = kCH * kCW;                     // 561
constexpr int kMT = (kNPix + 15) / 16;               // 36 row tiles of 16 convolution pixels
constexpr int kIH = 4 * kPH + 7, kIW = 4 * kPW + 7;  // input patch 39 x 71
constexpr int kRS = 216;                             // patch row stride in bf16 elements (71*3 = 213, even, padded)
constexpr int kPatchRows = kIH + 1;                  // one spare row: the zero-weight K padding reads into it
constexpr int kK = 160;                              // 7 filter rows x 22 (21 taps + 1 pad) = 154, padded to 10 k-steps
constexpr int kWS = 168;                             // weight row stride (elements): 84 words -> conflict-free fragments
constexpr int kWarps = 9;                            // 36 row tiles = 9 warps x 2 rounds x 2 tiles
constexpr int kThreads = kWarps * 32;
static_assert(kThreads % 72 == 0 && 72 * 3 == kRS, "patch loader: one column per thread");

struct StemParams {
  const float* img;      // [B,3,H,W] fp32
  const bf16* w;         // [64][kK] packed: k = ky*22 + kx*3 + c
  const float* bias;     // [64] folded BN shift
  bf16* out;             // [B, PH_all, PW_all, 64]
  int B, H, W, CH_all, CW_all, PH_all, PW_all, tiles_x, tiles_y;
};

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t bf2max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(kThreads, 2) gwd_stem_kernel(const StemParams p) {
  extern __shared__ __align__(16) uint8_t smraw[];
  uint32_t* patch = reinterpret_cast<uint32_t*>(smraw);                       // [kPatchRows][kRS/2] words
  uint32_t* wsm = patch + kPatchRows * (kRS / 2);                             // [64][kWS/2] words
  uint32_t* conv = wsm + 64 * (kWS / 2);                                      // [kNPix][32] words, 16-byte chunks swizzled
  float* bias_s = reinterpret_cast<float*>(conv + kNPix * 32);                // [64]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y, b = tile / p.tiles_y;
  const int py0 = ty * kPH, px0 = tx * kPW;        // pooled origin
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // convolution-pixel origin (max-pool padding 1)
  const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;  // input origin (convolution padding 3)

  // ---- weights (bf16, 16-byte pieces) and bias ----
  for (int i = tid; i < 64 * (kK / 8); i += kThreads) {
    const int n = i / (kK / 8), q = i - n * (kK / 8);
    *reinterpret_cast<uint4*>(wsm + n * (kWS / 2) + q * 4) = __ldg(reinterpret_cast<const uint4*>(p.w + n * kK) + q);
  }
  if (tid < 64) bias_s[tid] = p.bias[tid];
  // ---- input patch: fp32 planar -> bf16 [row][col][c]; out-of-image pixels are the convolution's zero padding ----
  {
    // kThreads = 4 * 72: every thread keeps one patch column and walks the (row, channel) planes four at a time, so a
    // warp reads consecutive floats of one image row
    bf16* pb = reinterpret_cast<bf16*>(patch);
    const int col = tid % 72, ix = ix0 + col;
    const bool col_ok = col < kIW && ix >= 0 && ix < p.W;
    const float* plane0 = p.img + static_cast<int64_t>(b) * 3 * p.H * p.W + ix;
    for (int rc = tid / 72; rc < kPatchRows * 3; rc += kThreads / 72) {
      const int r = rc / 3, c = rc - 3 * r;
      const int iy = iy0 + r;
      float v = 0.f;
      if (col_ok && r < kIH && iy >= 0 && iy < p.H) v = __ldg(plane0 + (static_cast<int64_t>(c) * p.H + iy) * p.W);
      pb[r * kRS + col * 3 + c] = __float2bfloat16(v);
    }
  }
  // per-thread K offsets (32-bit words inside the patch) of the fragment columns k0 = 16 ks + 2 t (+ 8)
  int koff[10][2];
#pragma unroll
  for (int ks = 0; ks < 10; ++ks) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k0 = 16 * ks + 8 * h + 2 * t;
      const int ky = k0 / 22, j = k0 - ky * 22;
      koff[ks][h] = (ky * kRS + j) >> 1;
    }
  }
  __syncthreads();

  // ---- convolution pixels: each warp takes pairs of 16-pixel row tiles (the weight fragments serve both) ----
  for (int pair = warp; pair < kMT / 2; pair += kWarps) {
    int pbase[4];   // patch word offsets of this lane's four pixel rows: tile 0 rows g, g+8; tile 1 rows g, g+8
    int pidx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int pi = 32 * pair + 16 * (i >> 1) + 8 * (i & 1) + g;
      pidx[i] = pi;
      if (pi >= kNPix) pi = kNPix - 1;
      const int cy = pi / kCW, cx = pi - cy * kCW;
      pbase[i] = cy * kRS + 3 * cx;   // (2 cy * kRS + 6 cx) / 2
    }
    float acc[2][8][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) acc[m][nt][0] = acc[m][nt][1] = acc[m][nt][2] = acc[m][nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 10; ++ks) {
      uint32_t a0[4], a1[4];
      a0[0] = patch[pbase[0] + koff[ks][0]];
      a0[1] = patch[pbase[1] + koff[ks][0]];
      a0[2] = patch[pbase[0] + koff[ks][1]];
      a0[3] = patch[pbase[1] + koff[ks][1]];
      a1[0] = patch[pbase[2] + koff[ks][0]];
      a1[1] = patch[pbase[3] + koff[ks][0]];
      a1[2] = patch[pbase[2] + koff[ks][1]];
      a1[3] = patch[pbase[3] + koff[ks][1]];
      const uint32_t* wrow = wsm + g * (kWS / 2) + 8 * ks + t;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0 = wrow[nt * 8 * (kWS / 2)], b1 = wrow[nt * 8 * (kWS / 2) + 4];
        mma_16816(acc[0][nt], a0, b0, b1);
        mma_16816(acc[1][nt], a1, b0, b1);
      }
    }
    // folded BN shift + ReLU -> bf16 tile; convolution pixels outside the image become 0 (ReLU outputs are >= 0 and
    // every pooling window holds a real pixel, so 0 is as good as the pool's -inf padding)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pi = pidx[i];
      if (pi < kNPix) {
        const int cy = pi / kCW, cx = pi - cy * kCW;
        const bool inside = cy0 + cy >= 0 && cy0 + cy < p.CH_all && cx0 + cx >= 0 && cx0 + cx < p.CW_all;
        const int m = i >> 1, hi = i & 1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float v0 = inside ? fmaxf(acc[m][nt][2 * hi] + bias_s[8 * nt + 2 * t], 0.f) : 0.f;
          const float v1 = inside ? fmaxf(acc[m][nt][2 * hi + 1] + bias_s[8 * nt + 2 * t + 1], 0.f) : 0.f;
          conv[pi * 32 + ((nt ^ (pi & 7)) << 2) + t] = gwd_pack_bf16x2(v0, v1);
        }
      }
    }
  }
  __syncthreads();

  // ---- 3x3 stride-2 max-pool out of shared memory; one 16-byte (8-channel) piece per work item ----
  for (int i = tid; i < kPH * kPW * 8; i += kThreads) {
    const int c8 = i & 7, pp = i >> 3;
    const int py = pp / kPW, px = pp - py * kPW;
    if (py0 + py >= p.PH_all || px0 + px >= p.PW_all) continue;
    uint4 m = make_uint4(0u, 0u, 0u, 0u);   // bf16 +0: ReLU outputs are non-negative
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int pi = (2 * py + dy) * kCW + 2 * px + dx;
        const uint4 v = *reinterpret_cast<const uint4*>(conv + pi * 32 + ((c8 ^ (pi & 7)) << 2));
        m.x = bf2max(m.x, v.x); m.y = bf2max(m.y, v.y); m.z = bf2max(m.z, v.z); m.w = bf2max(m.w, v.w);
      }
    }
    bf16* dst = p.out + ((static_cast<int64_t>(b) * p.PH_all + py0 + py) * p.PW_all + px0 + px) * 64 + c8 * 8;
    *reinterpret_cast<uint4*>(dst) = m;
  }
}

}  // namespace

extern "C" int gwd_stem_conv_pool(const float* images, const void* w_packed, const float* bias, void* out, int32_t B,
                                  int32_t H, int32_t W, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(images && w_packed && bias && out, "gwd_stem_conv_pool: null pointer");
  GWD_CHECK_ARG(B > 0 && H >= 8 && W >= 8, "gwd_stem_conv_pool: bad shape");
  GWD_CHECK_ARG((reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "gwd_stem_conv_pool: weights / output must be 16-byte aligned");
  StemParams p;
  p.img = images; p.w = static_cast<const bf16*>(w_packed); p.bias = bias; p.out = static_cast<bf16*>(out);
  p.B = B; p.H = H; p.W = W;
  p.CH_all = (H + 2 * 3 - 7) / 2 + 1;          // convolution output size (7x7, stride 2, padding 3)
  p.CW_all = (W + 2 * 3 - 7) / 2 + 1;
  p.PH_all = (p.CH_all + 2 * 1 - 3) / 2 + 1;   // pooled size (3x3, stride 2, padding 1)
  p.PW_all = (p.CW_all + 2 * 1 - 3) / 2 + 1;
  p.tiles_x = (p.PW_all + kPW - 1) / kPW;
  p.tiles_y = (p.PH_all + kPH - 1) / kPH;
  const size_t smem = static_cast<size_t>(kPatchRows) * kRS * 2 + 64 * kWS * 2 + static_cast<size_t>(kNPix) * 128 + 64 * 4;
  static bool configured = false;
  if (!configured) {
    GWD_CUDA(cudaFuncSetAttribute(gwd_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = true;
  }
  const int64_t tiles = static_cast<int64_t>(B) * p.tiles_x * p.tiles_y;
  gwd_stem_kernel<<<static_cast<unsigned>(tiles), kThreads, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}
