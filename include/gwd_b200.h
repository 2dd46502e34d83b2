/*
 * gwd_b200.h -- C ABI of libgwd_b200.so, the sm_100a kernel library behind the
 * GW-Depth model forward path (BASELINE.json north_star).
 *
 * The reference (ViktorLiang/GW-Depth) ships no native code: every entry point
 * below replaces a chain of PyTorch library calls at the cited reference call
 * site (paths relative to the reference root).  All pointers are CALLER-OWNED
 * DEVICE memory unless a parameter says "host"; sizes are element counts; no
 * hidden allocation; every function enqueues on `stream` (a cudaStream_t passed
 * as void*) and returns 0 on success or a negative gwd_status, with a
 * thread-local message available from gwd_last_error().
 *
 * Layout convention: activations are channels-last ("NHWC"): a [B,H,W,C] map or
 * a [rows, C] token matrix (B=1,H=1,W=rows), bf16 unless stated, with an
 * explicit channel stride so that a tensor can be a channel slice of a wider
 * buffer (used to build concatenations without a copy).
 */
#ifndef GWD_B200_H_
#define GWD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum gwd_status {
  GWD_OK = 0,
  GWD_ERR_ARG = -1,      /* bad argument (shape / alignment / null) */
  GWD_ERR_CUDA = -2,     /* CUDA runtime / driver error, see gwd_last_error() */
  GWD_ERR_NODEVICE = -3  /* no sm_100 device */
};

enum gwd_act { GWD_ACT_NONE = 0, GWD_ACT_RELU = 1, GWD_ACT_GELU = 2, GWD_ACT_ELU = 3, GWD_ACT_SIGMOID = 4 };
enum gwd_res { GWD_RES_NONE = 0, GWD_RES_BEFORE_NORM = 1, GWD_RES_AFTER = 2 };

const char* gwd_last_error(void);
int gwd_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t gwd_launch_count(void);
void gwd_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * gwd_conv_gemm: the tensor-core workhorse (tcgen05.mma + TMEM accumulators, operands staged
 * by TMA).  Computes, for every output pixel p and output channel n,
 *     acc[p,n] = sum_{tap, c} x[p + tap_offset, x_coff + c] * w[tap][n][c]       (fp32 accumulate)
 * with taps = 1 (a Linear / 1x1 conv) or taps = 9 (3x3, stride 1, zero padding 1), followed by the
 * fused epilogue, in this order:
 *     v = acc + bias[n];  v = pre_act(v);  if res_mode==BEFORE_NORM: v += res[p,n]
 *     if y_raw: y_raw[p,n] = v
 *     if ln_g:  v = (v - mean_n(v)) * rsqrt(var_n(v) + ln_eps) * ln_g[n] + ln_b[n]   (over the n logical channels)
 *     v = post_act(v) * out_scale;  if res_mode==AFTER: v += res[p,n];  y[p,n] = v
 * Replaces: nn.Linear/F.linear + bias + activation (+ residual + nn.LayerNorm) chains of
 *   src/models/transformer.py:149-162,212-233, src/models/multi_head_attention.py:236-254,373,
 *   src/models/multiscale_transformerr.py:67-73,275,329,755, src/models/glassrgbd.py:87-90,101;
 * and cuDNN conv2d -> permute -> LayerNorm -> permute -> GELU (+ residual) chains of
 *   src/models/points/points_sample.py:12-43,106-125, src/models/dense_upsample.py:82-90,160-182,
 *   src/models/multiscale_transformerr.py:104-118.
 * Constraints: cin % 16 == 0, n_pad % 16 == 0, n <= n_pad, channel strides/offsets % 8 == 0,
 *   LayerNorm only when n_pad <= 256.  w is packed bf16 [taps][n_pad][cin] with tap index
 *   dx*3+dy for taps = 9 (see gwd_pack_conv3x3_weight in the Python host code).
 * ------------------------------------------------------------------------------------------ */
typedef struct gwd_gemm_desc {
  const void* x;        /* bf16 [B,H,W,x_cstride] */
  int32_t B, H, W;
  int32_t x_cstride;    /* channel stride of x (elements) */
  int32_t x_coff;       /* first channel consumed */
  int32_t cin;          /* channels consumed per tap (K per tap) */
  const void* w;        /* bf16 packed weights [taps][n_pad][cin] */
  int32_t taps;         /* 1 or 9 */
  int32_t n_pad;        /* physical output channels (multiple of 16) */
  int32_t n;            /* logical output channels (LayerNorm width, store width) */
  const float* bias;    /* [n_pad] or NULL */
  const float* ln_g;    /* [n_pad] or NULL -> no LayerNorm */
  const float* ln_b;    /* [n_pad] */
  float ln_eps;
  int32_t pre_act;      /* gwd_act applied before the norm */
  int32_t post_act;     /* gwd_act applied after the norm */
  float out_scale;      /* multiplies the activated value (1.0f for none) */
  const void* res;      /* bf16 [B,H,W,res_cstride] or NULL */
  int32_t res_cstride, res_coff;
  int32_t res_mode;     /* gwd_res */
  void* y;              /* output [B,H,W,y_cstride], bf16 or f32 */
  int32_t y_cstride, y_coff;
  int32_t y_f32;        /* 1: y is float32 */
  void* y_raw;          /* optional bf16 copy of the pre-norm value, or NULL */
  int32_t yraw_cstride, yraw_coff;
  int32_t store_n;      /* channels stored (<= n_pad); pad channels written as 0 when store_n > n */
} gwd_gemm_desc;

int gwd_conv_gemm(const gwd_gemm_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWD_B200_H_ */
