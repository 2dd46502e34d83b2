// gwd_common.cuh -- shared host/device helpers for libgwd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <utility>
#include "../../include/gwd_b200.h"

#ifndef GWD_PDL_DEFAULT
#define GWD_PDL_DEFAULT 1
#endif

// ----------------------------------------------------------------------------------------------
// host side: error reporting + launch accounting
// ----------------------------------------------------------------------------------------------
void gwd_set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_gwd_launches;

#define GWD_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      gwd_set_error(__VA_ARGS__);           \
      return GWD_ERR_ARG;                   \
    }                                       \
  } while (0)

#define GWD_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      gwd_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GWD_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define GWD_LAUNCHED()                                                                   \
  do {                                                                                   \
    g_gwd_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      gwd_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GWD_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

int gwd_num_sms();

static inline int64_t gwd_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL).  The model is ~500 (forward) / ~1 700 (training step) short kernels replayed from CUDA graphs;
// a tcgen05 GEMM spends 1-2 us before it touches global memory (kernel parameters, barrier initialisation, tensor-map prefetch, TMEM
// allocation, one CTA-wide barrier).  Launched with cudaLaunchAttributeProgrammaticStreamSerialization that prologue runs while the
// kernel in front of it on the stream is still finishing: the front kernel executes gwd_pdl_trigger() (griddepcontrol.
// launch_dependents) once its CTAs are running, the dependent executes gwd_pdl_wait() (griddepcontrol.wait: returns when the front
// grid has COMPLETED and its memory operations are visible) before its first global-memory access.  Rules kept by every kernel that
// is launched through gwd_launch(): nothing is read from or written to global memory ahead of gwd_pdl_wait().  Both instructions
// are no-ops in a kernel that was launched without the attribute / has no programmatic dependent.  GWD_PDL=0 switches the launch
// attribute off (the instructions stay, as no-ops).
// ----------------------------------------------------------------------------------------------
static inline bool gwd_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("GWD_PDL"); return e ? atoi(e) != 0 : GWD_PDL_DEFAULT != 0; }();
  return on;
}

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t gwd_launch(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = static_cast<unsigned>(cluster_x); at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (gwd_pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kfn, std::forward<Args>(args)...);
}
#endif

// ----------------------------------------------------------------------------------------------
// device side: activations, packing, warp reductions
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void gwd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gwd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// GELU.  The reference uses the erf form (nn.GELU()); here 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) on the MUFU
// tanh unit: 5 FP instructions + 1 MUFU instead of 15 + 2 for an erf polynomial.  |tanh form - erf form| <= 4.8e-4
// absolute (3e-4 relative where it peaks, |x| ~ 2.2) and tanh.approx adds ~5e-4 relative, both an order of magnitude
// below the bf16 rounding (3.9e-3) every GELU output goes through; the GELU epilogues were FP-issue bound.
__device__ __forceinline__ float gwd_tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gwd_gelu(float v) {
  const float u = v * fmaf(0.0356774081f, v * v, 0.7978845608f);   // sqrt(2/pi) (v + 0.044715 v^3)
  const float h = 0.5f * v;
  return fmaf(h, gwd_tanh_approx(u), h);
}
// d gelu(v) / dv of the tanh form above: 0.5 (1 + t) + 0.5 v (1 - t^2) u' with u = sqrt(2/pi) (v + 0.044715 v^3), t = tanh u.
// 9 FP instructions + 1 MUFU instead of ~45 for erff + expf; within 9e-4 of the erf form's derivative (nn.GELU) everywhere, and
// it is the exact derivative of what the forward kernels compute.  The LayerNorm / activation backward kernels evaluate it per
// element and are bound by instruction issue.
__device__ __forceinline__ float gwd_gelu_grad(float v) {
  const float v2 = v * v;
  const float t = gwd_tanh_approx(v * fmaf(0.0356774081f, v2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, v2, 0.7978845608f);
  return fmaf(0.5f * v * fmaf(-t, t, 1.f), du, fmaf(0.5f, t, 0.5f));
}
// d act / d (input) as a factor: from the activation's OUTPUT yy (from_input == 0: ReLU, ELU, sigmoid are invertible there) or from
// its INPUT yy (from_input != 0; GELU needs it).  Shared by gwd_act_bwd, gwd_layernorm_bwd and the activation-gradient epilogue of
// the data-gradient GEMMs (gwd_conv_gemm, res_mode = GWD_RES_MUL_ACTGRAD).
__device__ __forceinline__ float gwd_act_grad(float v, int act) {
  switch (act) {
    case GWD_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case GWD_ACT_GELU: return gwd_gelu_grad(v);
    case GWD_ACT_ELU: return v > 0.f ? 1.f : __expf(v);
    case GWD_ACT_SIGMOID: { const float s = 1.f / (1.f + __expf(-v)); return s * (1.f - s); }
    default: return 1.f;
  }
}
__device__ __forceinline__ float gwd_act_factor(float yy, int act, int from_input) {
  if (from_input) return gwd_act_grad(yy, act);
  switch (act) {
    case GWD_ACT_RELU: return yy > 0.f ? 1.f : 0.f;
    case GWD_ACT_ELU: return yy > 0.f ? 1.f : yy + 1.f;
    case GWD_ACT_SIGMOID: return yy * (1.f - yy);
    default: return 1.f;
  }
}
// erf(x) by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7), kept for callers that need erf itself
__device__ __forceinline__ float gwd_erf(float x) {
  float ax = fabsf(x);
  float t = __frcp_rn(fmaf(0.3275911f, ax, 1.f));
  float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  float r = 1.f - poly * __expf(-ax * ax);
  return copysignf(r, x);
}

__device__ __forceinline__ float gwd_apply_act(float v, int act) {
  switch (act) {
    case GWD_ACT_RELU: return fmaxf(v, 0.f);
    case GWD_ACT_GELU: return gwd_gelu(v);
    case GWD_ACT_ELU: return fmaxf(v, 0.f) + (__expf(fminf(v, 0.f)) - 1.f);
    case GWD_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

__device__ __forceinline__ uint32_t gwd_pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float2 gwd_unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}

// ----------------------------------------------------------------------------------------------
// dropout masks (train-mode nn.Dropout of the DETR layers, src/models/transformer.py:149-162,212-233 and the attention-
// probability dropout of src/models/multi_head_attention.py:368).  The mask of an element is a pure function of (step seed in
// device memory, site id, element index): the forward and the backward kernels regenerate it instead of storing it, and the
// seed lives in device memory so that a replayed CUDA graph draws new masks every step.  32-bit integer mixer (two
// multiply-xorshift rounds, full avalanche); keep probability = 1 - p to 2^-32.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gwd_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// stream key of one dropout site (and, for attention, one (image, head)): mixed once per thread, then one mix per element
__device__ __forceinline__ uint32_t gwd_drop_key(uint32_t seed, uint32_t site, uint32_t sub) {
  return gwd_mix32(seed ^ gwd_mix32(site * 0x9E3779B9U + sub * 0x85EBCA6BU + 0x6A09E667U));
}
__device__ __forceinline__ bool gwd_drop_keep(uint32_t key, uint32_t idx, uint32_t threshold) {
  return gwd_mix32(idx ^ key) >= threshold;          // threshold = p * 2^32
}

__device__ __forceinline__ float gwd_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float gwd_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__
