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
enum gwd_res { GWD_RES_NONE = 0, GWD_RES_BEFORE_NORM = 1, GWD_RES_AFTER = 2,
               GWD_RES_MUL_ACTGRAD = 3 /* training: `res` is a SAVED ACTIVATION (output, or input with ag_from_input); the result is
                                          multiplied by d act / d input there and by ag_scale: the activation backward fused into
                                          the data-gradient GEMM that produces its operand.  No LayerNorm / activation / y_f32. */ };

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
  int32_t w_per_image;  /* 1: w is [B][taps][n_pad][cin], image b of x uses w[b] (PointBasedPred correlation,
                           src/models/points/points_sample.py:272) */
  int32_t upsample2;    /* 1: `upconv` (src/models/dense_upsample.py:82-90): the 3x3 conv acts on the nearest x2 up-sampled
                           input WITHOUT materialising it.  w holds 4 phase filters stacked along N (n = 4*Cout, phase
                           2*oy+ox): output pixel (2y+oy, 2x+ox) of the [B,2H,2W,y_cstride] output takes channels
                           [phase*Cout, (phase+1)*Cout).  A LayerNorm epilogue then normalises each phase group. */
  int64_t x_wstride, x_hstride, x_bstride;   /* pixel strides of x along W / H / B in PIXELS (0, 0, 0 = dense [B,H,W]); a
                           stride-2 1x1 projection (ResNet downsample, torchvision Bottleneck) is a Linear over the
                           view x[:, ::2, ::2, :]: H, W = the view's extents, strides = (2, 2*W0, H0*W0).  taps == 1 only. */
  int32_t ag_act, ag_from_input;             /* GWD_RES_MUL_ACTGRAD: the activation and whether `res` holds its input */
  float ag_y_mul, ag_scale;                  /* factor = act'(res * ag_y_mul) * ag_scale (0 is read as 1) */
} gwd_gemm_desc;

int gwd_conv_gemm(const gwd_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------
 * gwd_attention: fused multi-head softmax attention  O = softmax(Q K^T * scale + bias + mask) V
 * for one (item, head) at a time with K/V staged in shared memory (head_dim in {4,8,16,32}).
 * Q/K/V/O are bf16 views: element (item, row, head h, d) lives at  base + item*item_stride +
 * row*row_stride + h*head_dim + d.  bias is fp32 [heads, Lq, Lk] (relative position bias), mask is
 * fp32 [mask_windows, Lq, Lk] selected by item % mask_windows (shifted-window mask, value -100),
 * key_padding is uint8 [items, Lk] (1 = key ignored, scores -> -inf).
 * Replaces src/models/multi_head_attention.py:317-372 (bmm, masked_fill, softmax, bmm, transposes; the
 * head-averaged weights of :375-378 are dead and not produced) and the attention cores of
 * src/models/multiscale_transformerr.py:311-328 and :539-556.
 * Three implementations behind this one entry point (the call never falls back to the host):
 *   - head_dim 32, no bias / mask, Lk <= 480 (DETR self / cross attention): tcgen05 + TMEM (gwd_attn_tc.cu);
 *   - bias given, Lq == Lk == 49 (7x7 windows): persistent mma.sync kernel with the bias table in shared memory
 *     (gwd_attn_win.cu); the window mask must be two-valued (0 / one negative constant), as the reference builds it;
 *   - anything else: thread-per-query kernel with K/V in shared memory (gwd_attn.cu).
 * ------------------------------------------------------------------------------------------ */
typedef struct gwd_attn_desc {
  const void* q; const void* k; const void* v; void* o;
  int32_t items, heads, Lq, Lk, hd;
  int64_t q_item_stride, q_row_stride, k_item_stride, k_row_stride;
  int64_t v_item_stride, v_row_stride, o_item_stride, o_row_stride;
  const float* bias;
  const float* mask;
  int32_t mask_windows;
  const uint8_t* key_padding;
  float scale;
  /* train-mode dropout of the attention probabilities (src/models/multi_head_attention.py:368): P <- keep ? P / (1 - p) : 0
   * ahead of P V.  dropout_seed: DEVICE uint32 (NULL or dropout_p == 0: no dropout), dropout_site: id of the call site; the mask
   * of element (item, head, query, key) is a pure function of (*dropout_seed, dropout_site, indices) and is regenerated by
   * gwd_attention_bwd.  Supported on the tcgen05 path (head dim 32, no bias / window mask). */
  const uint32_t* dropout_seed;
  uint32_t dropout_site;
  float dropout_p;
} gwd_attn_desc;
int gwd_attention(const gwd_attn_desc* d, void* stream);

/* class-token CHANNEL attention of WindowClassAttention (multiscale_transformerr.py:561-578): for every window
 * (item) and head, A = softmax_c(scale * tq^T tk) [td x tc], out = (A tv^T)^T; depth and seg tokens share tk/tv. */
int gwd_token_attention(const void* depth_q, const void* seg_q, const void* tk, const void* tv, void* depth_out,
                        void* seg_out, int32_t items, int32_t N, int32_t heads, int32_t td, int32_t tc, int64_t q_rs,
                        int64_t k_rs, int64_t v_rs, int64_t o_rs, float scale, void* stream);

/* line end-point re-query of WindowAttention (multiscale_transformerr.py:281-310):
 *   gwd_ref_scores : a[b][h][w*N+n][r] = scale * q . ref_k                      (:295-298)
 *   gwd_ref_diffuse: a_out = a_in + gelu(layer_norm_[P,R](conv3x3_{h->h}(a_in)))  one of the 3 steps of :299-302
 *   gwd_ref_requery: q_new = scale * softmax_r(a) ref_v                          (:307-310), bf16 window layout */
int gwd_ref_scores(const void* q, int64_t q_rs, const float* ref_k, int64_t ref_rs, float* a, int32_t B, int32_t nW,
                   int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream);
int gwd_ref_diffuse(const float* a_in, float* a_out, const float* conv_w_host, const float* conv_b_host, float* raw_workspace,
                    double* stats_workspace, int32_t B, int32_t heads, int32_t P, int32_t R, void* stream);
/* conv_w_host / conv_b_host: HOST arrays [16*16*9] / [16] (they travel as kernel parameters);
 * raw_workspace: device fp32 [B*heads*P*R]; stats_workspace: device fp64 [B*heads*2] */
int gwd_ref_requery(const float* a, const float* ref_v, int64_t ref_rs, void* q_new, int64_t o_rs, int32_t B, int32_t nW,
                    int32_t N, int32_t heads, int32_t hd, int32_t R, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * bandwidth kernels (bf16 channels-last rows, strides in elements, C and strides multiples of 8)
 * ------------------------------------------------------------------------------------------ */
/* out = act(LayerNorm_n(x + res)); res / gamma may be NULL.  nn.LayerNorm + residual adds of
 * src/models/transformer.py:113-120,157-161 and multiscale_transformerr.py:659,664-665 */
int gwd_layernorm(const void* x, int64_t x_rs, const void* res, int64_t res_rs, const float* gamma, const float* beta,
                  float eps, int32_t act, void* out, int64_t out_rs, int64_t rows, int32_t C, int32_t n, void* stream);
/* out[r] = x[r] + addend[r % period]   (with_pos_embed, src/models/transformer.py:154,219,224-225) */
int gwd_add_rows(const void* x, int64_t x_rs, const void* addend, int64_t a_rs, int64_t period, void* out, int64_t out_rs,
                 int64_t rows, int32_t C, void* stream);
/* LayerNorm + zero pad + cyclic shift + window partition in one pass (multiscale_transformerr.py:659-707) */
int gwd_window_gather(const void* x, int64_t x_rs, const float* gamma, const float* beta, float eps, void* out,
                      int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C, int32_t n,
                      void* stream);
/* window reverse + un-shift + crop + residual add (+ LayerNorm of the sum into out_ln) (:731-755) */
int gwd_window_merge(const void* win, int64_t win_rs, const void* shortcut, int64_t sc_rs, void* out, int64_t out_rs,
                     const float* gamma, const float* beta, float eps, void* out_ln, int64_t ln_rs, int32_t B, int32_t H,
                     int32_t W, int32_t ws, int32_t shift, int32_t C, int32_t n, void* stream);
/* F.interpolate(mode='nearest') (+ optional add of a same-size map) */
int gwd_upsample_nearest(const void* x, int64_t x_rs, int32_t B, int32_t h, int32_t w, void* out, int64_t out_rs, int32_t H,
                         int32_t W, int32_t C, const void* add, int64_t add_rs, void* stream);
/* nn.AvgPool2d(k, stride=k)  (points_sample.py:61-75) */
int gwd_avgpool(const void* x, int64_t x_rs, int32_t B, int32_t H, int32_t W, int32_t k, void* out, int64_t out_rs,
                int32_t C, void* stream);
/* The four pooling branches of PyramidLayer (nn.AvgPool2d(k, k) for k = 2, 4, 8, 16, points_sample.py:61-75,115-121) in one pass
 * over x [B,H,W,C] (row stride x_rs): o2 / o4 / o8 / o16 = contiguous bf16 [B, H/k, W/k, C], floor mode; H, W >= 16. */
int gwd_avgpool_pyramid(const void* x, int64_t x_rs, int32_t B, int32_t H, int32_t W, void* o2, void* o4, void* o8, void* o16,
                        int32_t C, void* stream);
/* gwd_bilinear_up of four contiguous maps x_j [B, hw[2j], hw[2j+1], C] (hw: HOST array of 8) into the adjacent channel slices
 * [j*C, (j+1)*C) of out (row stride out_rs, already offset to the first slice): the PyramidLayer concat, points_sample.py:115-121 */
int gwd_bilinear_up4(const void* x0, const void* x1, const void* x2, const void* x3, const int32_t* hw, int32_t B, void* out,
                     int64_t out_rs, int32_t H, int32_t W, int32_t C, void* stream);
/* F.interpolate(mode='bilinear', align_corners=True)  (points_sample.py:115-121) */
int gwd_bilinear_up(const void* x, int64_t x_rs, int32_t B, int32_t h, int32_t w, void* out, int64_t out_rs, int32_t H,
                    int32_t W, int32_t C, void* stream);
/* F.grid_sample(bilinear, align_corners=False, zeros) of a bf16 map (+ fp32 [H,W,C] table) at K points -> fp32 [B,K,C]
 * (points_sample.py:264-267).  table_bstride: elements between the tables of consecutive images (0 = one table shared
 * by the batch; per-image tables are what a padded batch has, src/models/position_encoding.py:33-35). */
int gwd_sample_bilinear(const void* x, int64_t x_rs, int32_t x_coff, const float* table, int64_t table_bstride, int32_t B,
                        int32_t H, int32_t W, int32_t C, const float* coords, int32_t K, float* out, void* stream);
/* same for a 1-channel fp32 map -> fp32 [B,K]  (points_sample.py:268) */
int gwd_sample_scalar(const float* x, int32_t B, int32_t H, int32_t W, const float* coords, int32_t K, float* out,
                      void* stream);
/* ---- Training data path, pixel side: src/datasets/transforms_depth.py on decoded uint8 images, bit for bit
 * what Pillow / torchvision do to the PIL image in the reference's DataLoader workers. ---- */
/* HOST: Pillow ImagingResample coefficient tables of a BILINEAR (antialiased) resize of one axis (transforms_depth.py:343 F.resize):
 * ksize = table width; xmin / cnt int32 [out_size], kk int32 [out_size * ksize] (22-bit fixed point). */
int gwd_pil_bilinear_ksize(int32_t in_size, int32_t out_size);
int gwd_pil_bilinear_coeffs(int32_t in_size, int32_t out_size, int32_t* xmin, int32_t* cnt, int32_t* kk);
/* HOST: source index of every output index of a Pillow NEAREST resize (transforms_depth.py:368-370, the auxiliary maps) */
int gwd_pil_nearest_index(int32_t in_size, int32_t out_size, int32_t* idx);
/* One 8-bit pass of the resize along x (axis 1: src [H,W,C] rows src_rs bytes apart -> dst [H,out_size,C]) or y (axis 0: -> dst
 * [out_size,W,C]); tables on the DEVICE; flip != 0 mirrors the source index along that axis (hflip :206 / vflip :234 come first). */
int gwd_resample_u8(const void* src, int64_t src_rs, int32_t H, int32_t W, int32_t C, void* dst, int32_t out_size, int32_t axis,
                    const int32_t* xmin, const int32_t* cnt, const int32_t* kk, int32_t ksize, int32_t flip, void* stream);
/* dst[y,x] = src[sy, sx] for maps of 1 / 2 / 4 / 8-byte elements (rows src_rs ELEMENTS apart): sy = iy[y] (device int32 table, null =
 * y), mirrored when flip_v; same for x.  Nearest resize, flips and crops of the depth / segmentation maps in one pass. */
int gwd_gather2d(const void* src, int64_t src_rs, int32_t elem_bytes, int32_t H, int32_t W, void* dst, int32_t oh, int32_t ow,
                 const int32_t* iy, const int32_t* ix, int32_t flip_h, int32_t flip_v, void* stream);
/* ColorJitter (transforms_depth.py:551-604) in place on uint8 [npix,3]: up to 4 ops (HOST arrays; 0 brightness, 1 contrast,
 * 2 saturation, 3 hue) in the given order.  Contrast must be the first op of a launch and reads the L sum of the image from gray_in
 * (device uint64, written by the previous launch through gray_out; n_ops = 0 just measures). */
int gwd_jitter_u8(void* img, int64_t npix, int32_t n_ops, const int32_t* ops, const float* factors, const void* gray_in, void* gray_out,
                  void* stream);
/* Input builder (src/datasets/coco.py:77-78 ToTensor + Normalize, src/util/misc.py:291-313 padding + mask): table is a
 * DEVICE int64 [B][3] = {pointer to a uint8 HWC image on the device, its height, its width}; out fp32 [B,3,H,W] =
 * ((u8 / 255) - mean[c]) / std[c] inside an image and 0 in its padding (IEEE division: bit-identical to torchvision);
 * mask uint8 [B,H,W] = 1 on padding (optional).  mean3 / std3 are HOST arrays. */
int gwd_images_to_batch(const int64_t* table, int32_t B, int32_t H, int32_t W, const float* mean3, const float* std3, float* out,
                        uint8_t* mask, void* stream);
/* nearest sample of the windowed feature map + shifted position table at R line end points
 * (multiscale_transformerr.py:676-701) -> bf16 [B,R,C] */
int gwd_line_ref_gather(const void* win, int64_t win_rs, const float* pos, int64_t pos_bstride, const float* coords, int32_t R,
                        void* out, int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C, void* stream);
/* depth[p] = sum_k softmax_k(logits[p,:K]) anchor[b,k]  (points_sample.py:277-279) -> fp32 [B,HW] */
int gwd_anchor_mix(const void* logits, int64_t l_rs, const float* anchor, int32_t B, int64_t HW, int32_t K, float* out,
                   void* stream);
/* ResNet stem in one launch: 7x7/2 convolution (3 -> 64, padding 3) + folded FrozenBatchNorm shift + ReLU + 3x3/2
 * max-pool (padding 1); replaces conv1/bn1/relu/maxpool of the torchvision body wrapped by src/models/backbone.py:58-92
 * (FrozenBatchNorm2d :19-55).  images: fp32 [B,3,H,W]; w_packed: bf16 [64][160], column ky*22 + kx*3 + c holds the
 * BN-scaled weight w[n][c][ky][kx] (other columns zero); bias: fp32 [64]; out: bf16 [B, PH, PW, 64] channels-last with
 * PH = floor((floor((H-1)/2)+1 - 1)/2)+1 (the usual conv / pool output sizes). */
int gwd_stem_conv_pool(const float* images, const void* w_packed, const float* bias, void* out, int32_t B, int32_t H,
                       int32_t W, void* stream);
/* fp32 NCHW image -> bf16 NHWC with Cp >= C channels (zero padded) */
int gwd_nchw_to_nhwc(const float* x, int32_t B, int32_t C, int64_t HW, void* out, int32_t Cp, void* stream);

/* ------------------------------------------------------------------------------------------
 * selection / reduction kernels (fp32 inputs)
 * ------------------------------------------------------------------------------------------ */
/* CertainSample.forward (src/models/points/points_sample.py:291-364): K most uncertain pixels per image with the
 * reference's per-depth-bin quota / repeat / trim rules.  edges_host: HOST array of nbins+1 bin edges.
 * coords fp32 [B,K,2] = (x/W*2-1, y/H*2-1); index int32 [B,K] = y*W+x. */
int gwd_certain_sample(const float* pred_small, int32_t h, int32_t w, const float* pred_large, int32_t H, int32_t W,
                       int32_t B, int32_t K, const float* edges_host, int32_t nbins, float* coords, int32_t* index,
                       void* stream);
/* HungarianMatcher_Line cost, block diagonal only (src/models/matcher.py:52-71):
 * cost[tgt_offsets[b]*Q + q*T_b + t] = w_line*L1(lines[b,q], tgt[t]) - w_class*softmax(logits[b,q])[label[t]]
 * row_min (optional) fp32 [B,Q]. tgt_offsets int32 [B+1] (device). */
int gwd_match_cost(const float* logits, const float* lines, const float* tgt_lines, const int64_t* tgt_labels,
                   const int32_t* tgt_offsets, int32_t B, int32_t Q, int32_t num_classes, int32_t line_dim, float w_class,
                   float w_line, float* cost, float* row_min, void* stream);
/* HOST function (no GPU work): the assignments of HungarianMatcher_Line for a batch of problems
 * (src/models/matcher.py:73-74, scipy.optimize.linear_sum_assignment per image and decoder stage).  Problem p is the
 * row-major fp32 [Q, T[p]] block at cost + cost_offsets[p] (host memory, e.g. the D2H copy of gwd_match_cost's output).
 * Out: query_idx / target_idx int32 [n_problems][max(Q,1)], the first counts[p] = min(Q, T[p]) entries of row p are the
 * matched (query, target) pairs ordered by query, index-for-index what scipy returns (same algorithm and tie rules).
 * n_threads <= 0: one worker per host core (at most 16). */
int gwd_lsap_batch(const float* cost, const int64_t* cost_offsets, const int32_t* T, int32_t Q, int32_t n_problems,
                   int32_t* query_idx, int32_t* target_idx, int32_t* counts, int32_t n_threads);
/* compute_depth_errors per image (src/util/metrics.py:197-218 after the scrub of src/engine_glassrgbd.py:249-253):
 * metrics fp64 [B,9] = silog, abs_rel, log10, rms, sq_rel, log_rms, d1, d2, d3; workspace fp64 [B,10]. */
int gwd_depth_metrics(const float* pred, const float* gt, int32_t B, int64_t HW, float min_depth, float max_depth,
                      double* workspace, double* metrics, void* stream);
/* segmentation evaluation (src/engine_glassrgbd.py:232-240, src/util/metrics.py:43-78): confusion[gt * C + argmax_c logits]
 * += 1 over the pixels with gt != ignore_index (int64 [C*C], ACCUMULATED so a whole evaluation run can sum into it).
 * logits fp32 addressed as base + image*image_stride + pixel*pixel_stride + class*class_stride (NCHW: HW*C, 1, HW; the
 * channels-last map the forward produces: HW*C, C, 1); gt int64 [B, HW]. */
int gwd_seg_confusion(const float* logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride, const int64_t* gt,
                      int32_t B, int64_t HW, int32_t num_classes, int32_t ignore_index, int64_t* confusion, void* stream);
/* SilogLoss sums over valid pixels with gt nearest-resized to the prediction (src/models/glassrgbd.py:366-374,
 * src/engine_glassrgbd.py:74-80): sums3 fp64 = {count, sum d, sum d^2}. */
int gwd_silog_sums(const float* pred, int32_t B, int32_t h, int32_t w, const float* gt, int32_t H, int32_t W, float lo,
                   float hi, int32_t log_only, double* sums3, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training side of the line branch (DETR encoder / decoder + line heads): backward and optimizer kernels.
 * The reference gets all of this from torch.autograd over the modules of src/models/transformer.py:149-162,212-233,
 * src/models/multi_head_attention.py:317-373 and src/models/glassrgbd.py:87-90, and from torch.optim.AdamW +
 * clip_grad_norm_ (src/main_glassrgbd.py:59-67, src/engine_glassrgbd.py:155-159).
 * Data gradients dX = dY W and weight gradients dW = dY^T X are gwd_conv_gemm calls on transposed operands.
 * ------------------------------------------------------------------------------------------ */
/* y = post_act(LayerNorm_C(z) * gamma + beta):  dz[rows,C] (bf16) = backward of dy through the activation (evaluated from
 * the recomputed LN output; GWD_ACT_NONE: plain LayerNorm, beta may be NULL) and the LayerNorm (+ `add`, an optional bf16
 * gradient that joins at the same tensor, e.g. the residual branch); dgamma / dbeta fp32 [C] are ACCUMULATED (atomicAdd;
 * NULL = skip).  The statistics run over the n logical channels (n <= C; 0 = C), channels n..C are zero padding.
 * z is the pre-norm value (gwd_conv_gemm's y_raw).  Covers LayerNorm (transformer.py:149-233) and the
 * conv -> LayerNorm -> GELU blocks of src/models/points/points_sample.py:12-43. */
int gwd_layernorm_bwd(const void* dy, int64_t dy_rs, const void* z, int64_t z_rs, const float* gamma, const float* beta,
                      int32_t post_act, float eps, const void* add, int64_t add_rs, void* dz, int64_t dz_rs, float* dgamma,
                      float* dbeta, int64_t rows, int32_t C, int32_t n, void* stream);
/* out[r, c] (bf16, c < out_cols) = c < n ? dy[r,c] * act'(.) * scale : 0.  from_input == 0: act' is evaluated from the
 * activation's OUTPUT y * y_mul (GWD_ACT_RELU, GWD_ACT_SIGMOID, GWD_ACT_ELU; y_mul = 1 / max_depth and scale = max_depth
 * differentiate sigmoid * max_depth); from_input != 0: from its INPUT (also GWD_ACT_GELU, erf form).  GWD_ACT_NONE =
 * dtype conversion + padding (+ scale).  dy / y are fp32 or bf16. */
int gwd_act_bwd(const void* dy, int32_t dy_f32, int64_t dy_rs, const void* y, int32_t y_f32, int64_t y_rs, int32_t act,
                void* out, int64_t out_rs, int64_t rows, int32_t n, int32_t out_cols, float y_mul, float scale,
                int32_t from_input, void* stream);
/* out[c, r] = x[r, c] (bf16; r < rows, c < C), columns rows..rows_pad of out written as zeros; colsum (optional, fp32
 * [C]) is ACCUMULATED with the column sums of x = the bias gradient when x is dY. */
int gwd_transpose(const void* x, int64_t x_rs, void* out, int64_t out_rs, int64_t rows, int64_t rows_pad, int32_t C,
                  float* colsum, void* stream);
/* n_mats transposes in one launch (the W^T mirrors of every Linear of the branch after an optimizer step).  Device
 * tables: table int64 [n_mats][6] = {src ptr, dst ptr, rows, cols, src row stride, dst row stride} (bf16, even sizes,
 * 4-byte aligned), tile_prefix int32 [n_mats+1] = running count of 64x64 tiles (ceil(rows/64) * ceil(cols/64) each). */
int gwd_transpose_batch(const int64_t* table, const int32_t* tile_prefix, int32_t n_mats, int32_t total_tiles, void* stream);
/* nn.Linear weight gradient: dw[n, k] (fp32, row stride dw_rs) += sum_r dy[r, n] * x[r, k], db[n] += sum_r dy[r, n]
 * (db optional).  dy [rows, N] and x [rows, K] are bf16 with row strides; both are contracted over their slow axis, which
 * the kernel handles with ldmatrix.trans (no transposed copies); rows are split across CTAs and reduced with fp32
 * atomics, so dw / db must hold the running gradient (zeros at the start of a step). */
int gwd_linear_wgrad(const void* dy, int64_t dy_rs, const void* x, int64_t x_rs, int64_t rows, int32_t N, int32_t K, float* dw,
                     int64_t dw_rs, float* db, void* stream);
/* 3x3 convolution (stride 1, zero padding 1) weight gradient in the packed layout of gwd_conv_gemm:
 * dw[dx*3+dy][n][c] (fp32 [9][N][C]) += sum_{b,y,x} dy[b,y,x,n] * x[b,y+dy-1,x+dx-1,c]; db[n] += sum dy (optional).
 * dy bf16 [B,H,W,dy_cs], x bf16 [B,H,W,x_cs] channels-last; N, C multiples of 8.  (The data gradient of the same
 * convolution is gwd_conv_gemm of dy with the filter transposed in (n, c) and flipped in (dy, dx).) */
int gwd_conv3x3_wgrad(const void* dy, int64_t dy_cs, const void* x, int64_t x_cs, int32_t B, int32_t H, int32_t W, int32_t N,
                      int32_t C, float* dw, float* db, void* stream);
/* backward of O = softmax(scale Q K^T) V for head_dim 32, Lq, Lk <= 512 (no bias / mask): the soft-max is recomputed
 * from Q, K; dQ, dK, dV are bf16 views addressed like gwd_attn_desc.  With the forward output `o` given the kernel runs
 * on the tensor cores (mma.sync; D_i = dO_i . O_i); with o == NULL a CUDA-core kernel computes D itself (Lq, Lk <= 512, no mask).
 * Lq, Lk <= 1280. */
typedef struct gwd_attn_bwd_desc {
  const void* q; const void* k; const void* v; const void* d_o;
  void* dq; void* dk; void* dv;
  int32_t items, heads, Lq, Lk, hd;
  int64_t q_item_stride, q_row_stride, k_item_stride, k_row_stride, v_item_stride, v_row_stride;
  int64_t do_item_stride, do_row_stride, dq_item_stride, dq_row_stride, dk_item_stride, dk_row_stride;
  int64_t dv_item_stride, dv_row_stride;
  float scale;
  const void* o;          /* forward output (bf16), or NULL */
  int64_t o_item_stride, o_row_stride;
  float dq_mul, dk_mul;   /* dQ = dq_mul * dS K, dK = dk_mul * dS^T Q; 0 = `scale`.  For projections that were stored pre-scaled
                             (q' = a q, k' = b k, a b = softmax scale): scale = 1, dq_mul = a, dk_mul = b */
  const uint32_t* dropout_seed;   /* as in gwd_attn_desc: the forward's mask is regenerated (needs `o`, the forward output) */
  uint32_t dropout_site;
  float dropout_p;
  const uint8_t* key_padding;     /* [items, Lk], 1 = padded key (nn.MultiheadAttention key_padding_mask), or NULL (needs `o`) */
  float* stats_ws;                /* Lq or Lk in 513..1280 (two launches: per-query-block, per-key-block): fp32 scratch
                                     [items * heads * Lq * 2]; may be NULL for shorter sequences */
} gwd_attn_bwd_desc;
int gwd_attention_bwd(const gwd_attn_bwd_desc* d, void* stream);
/* SetCriterion forward + backward for all S decoder stages in one launch (src/models/glassrgbd.py:154-175,231-244,308-358):
 * logits fp32 [S,B,Q,C], lines fp32 [S,B,Q,D]; match int32 [4][M] = (stage, image, query, target row) of every matched
 * pair, grouped by stage with stage_off int32 [S+1]; class_w fp32 [C] (eos_coef on the last class); num_items fp32 [1]
 * on the device.  losses fp32 [S][2] = (loss_ce, loss_line); dlogits / dlines = gradients of
 * sum_s w_ce[s] loss_ce[s] + w_line[s] loss_line[s]. */
int gwd_set_loss(const float* logits, const float* lines, const float* tgt_lines, const int64_t* tgt_labels, const int32_t* match,
                 const int32_t* stage_off, const float* class_w, const float* w_ce, const float* w_line, const float* num_items,
                 int32_t S, int32_t B, int32_t Q, int32_t C, int32_t D, int32_t M, float* losses, float* dlogits, float* dlines,
                 void* stream);
/* *out_accum += sum g[i]^2 (fp64 accumulate; the caller zeroes it).  Input of the gradient clip below. */
int gwd_sumsq(const float* g, int64_t n, double* out_accum, void* stream);
/* torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.AdamW step `step` (1-based) over a flat fp32 segment:
 * g is first scaled by grad_scale (1 / world size after a sum all-reduce) and by the clip coefficient
 * min(1, max_norm / (sqrt(*sumsq) * grad_scale + 1e-6)) read on the DEVICE (no host sync; max_norm <= 0 or sumsq NULL =
 * no clipping); mirror_bf16 (optional) receives the bf16 copy of the new parameters that the forward kernels read. */
int gwd_adamw_step(float* p, const float* g, float* m, float* v, void* mirror_bf16, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int32_t step, float max_norm, float grad_scale,
                   const double* sumsq, void* stream);

/* Backward of the scale-invariant log loss of one prediction scale (src/models/glassrgbd.py:360-374 inside the loop of
 * src/engine_glassrgbd.py:65-82): loss = weight * 10 * sqrt(S2/n - variance_focus (S1/n)^2) with sums3 = (n, S1, S2) from
 * gwd_silog_sums on the same arguments (read on the DEVICE: no host sync).  sig_scale > 0: pred = sig_scale * sigmoid(z)
 * and the gradient is taken to z.  out_cols == 0: out is fp32 [B*h*w]; otherwise bf16 rows of out_cols columns (multiple of
 * 8) with the gradient in column 0 and zeros elsewhere = the dY operand of the last convolution's backward.  loss_out
 * (optional, fp32 [1]) receives the weighted loss value. */
int gwd_silog_bwd(const float* pred, int32_t B, int32_t h, int32_t w, const float* gt, int32_t H, int32_t W, float lo, float hi,
                  int32_t log_only, const double* sums3, float variance_focus, float weight, float sig_scale, void* out,
                  int32_t out_cols, float* loss_out, void* stream);
/* SegLoss = nn.CrossEntropyLoss (mean over pixels with gt != ignore_index; src/models/glassrgbd.py:376-383) times `weight`
 * (engine_glassrgbd.py:88-90), forward + backward: logits fp32 addressed as in gwd_seg_confusion, gt int64 [B*HW];
 * sums2 fp64 [2] = (valid pixels, sum of -log p[gt]) (zeroed here); dlogits (optional) bf16 rows of out_cols columns =
 * weight / n * (softmax - onehot), zeros beyond C; loss_out (optional, fp32 [1]) = weight * mean. */
int gwd_seg_ce(const float* logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride, const int64_t* gt,
               int32_t B, int64_t HW, int32_t C, int32_t ignore_index, float weight, double* sums2, void* dlogits,
               int32_t out_cols, float* loss_out, void* stream);
/* Backward of gwd_bilinear_up (F.interpolate(mode='bilinear', align_corners=True), the PyramidLayer branches of
 * src/models/points/points_sample.py:118-121): dy bf16 [B,H,W,dy_rs] (a channel slice of the concat gradient) ->
 * dx bf16 [B,h,w,dx_rs], C channels; a gather with the forward's own footprint arithmetic. */
int gwd_bilinear_up_bwd(const void* dy, int64_t dy_rs, int32_t B, int32_t H, int32_t W, void* dx, int64_t dx_rs, int32_t h,
                        int32_t w, int32_t C, void* stream);
/* Backward of gwd_avgpool (nn.AvgPool2d(k, k), floor mode) fused with the accumulation into the gradient of the pooled
 * map: out[b,Y,X,:] = add[b,Y,X,:] (optional; may alias out) + d[b,Y/k,X/k,:] * scale / k^2 inside the pooled region. */
int gwd_avgpool_bwd(const void* d, int64_t d_rs, int32_t k, float scale, const void* add, int64_t add_rs, void* out,
                    int64_t out_rs, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);

/* Backward of gwd_anchor_mix (PointBasedPred, src/models/points/points_sample.py:277-279): with a = softmax_k(logits[p,:K]) and
 * pred[p] = sum_k a_k anchor[b,k]:  dlogits[p,k] = a_k (anchor[b,k] - pred[p]) dpred[p] as bf16 rows of Kp columns (exact
 * zeros beyond K); danchor fp32 [B,K] += sum_p a_k dpred[p] (caller zeroes it).  K <= Kp <= 128. */
int gwd_anchor_mix_bwd(const void* logits, int64_t l_rs, const float* anchor, const float* dpred, int32_t B, int64_t HW, int32_t K,
                       int32_t Kp, void* dlogits, int64_t dl_rs, float* danchor, void* stream);
/* Backward of gwd_sample_bilinear w.r.t. the sampled bf16 map (F.grid_sample, align_corners=False, zero padding;
 * points_sample.py:264-267): d fp32 [B,K,C] -> dx bf16 [B,H,W,dx_rs] (C channels; every pixel written, zeros where no point
 * footprint lands).  A deterministic gather over the pixels, K <= 128. */
int gwd_sample_bilinear_bwd(const float* d, const float* coords, int32_t K, void* dx, int64_t dx_rs, int32_t B, int32_t H,
                            int32_t W, int32_t C, void* stream);
/* Backward of gwd_sample_scalar (anchor depths, points_sample.py:268): out fp32 [B,H,W] = add (optional, may alias out) +
 * the bilinear spread of d fp32 [B,K]. */
int gwd_sample_scalar_bwd(const float* d, const float* coords, int32_t K, const float* add, float* out, int32_t B, int32_t H,
                          int32_t W, void* stream);

/* Backward of the biased (shifted-)window self-attention core of the class-window Swin blocks (WindowClassAttention,
 * src/models/multiscale_transformerr.py:539-556 under torch.autograd): qkv bf16 [items*N, qkv_rs] = q | k | v (C = heads*hd
 * channels each, q un-scaled), S = scale q k^T + bias[head] + mask[item % mask_windows], P = softmax(S), O = P v.
 * d_o bf16 [items*N, do_rs] -> dqkv bf16 [items*N, dqkv_rs] = dq | dk | dv; dbias (optional) fp32 [heads, N, N] +=
 * sum over the windows of dS (the gradient of the gathered relative-position bias).  N <= 64, hd <= 32. */
int gwd_window_attention_bwd(const void* qkv, int64_t qkv_rs, const void* d_o, int64_t do_rs, void* dqkv, int64_t dqkv_rs,
                             const float* bias, const float* mask, int32_t mask_windows, float* dbias, int32_t items,
                             int32_t heads, int32_t N, int32_t hd, float scale, void* stream);
/* Backward of gwd_token_attention (class-token channel attention, multiscale_transformerr.py:561-578): inputs as the
 * forward (dq / sq: [rows, q_rs] token queries, td channels per head; tk / tv: [rows, k_rs / v_rs], tc channels per head)
 * plus the output gradients d_dout / d_sout [rows, o_rs]; writes the gradients of the queries (g_dq, g_sq: [rows, gq_rs])
 * and of the keys / values (g_tk, g_tv: [rows, gk_rs / gv_rs]; both token rows contribute).  N <= 64, 2 td <= 16, tc <= 32. */
int gwd_token_attention_bwd(const void* dq, const void* sq, const void* tk, const void* tv, const void* d_dout, const void* d_sout,
                            void* g_dq, void* g_sq, void* g_tk, void* g_tv, int32_t items, int32_t N, int32_t heads, int32_t td,
                            int32_t tc, int64_t q_rs, int64_t k_rs, int64_t v_rs, int64_t o_rs, int64_t gq_rs, int64_t gk_rs,
                            int64_t gv_rs, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training path of the 1/32 line-window stage ("glass-structure context", WindowAttention.forward,
 * src/models/multiscale_transformerr.py:267-332 under torch.autograd) -- gwd_train_line.cu
 * ------------------------------------------------------------------------------------------ */
/* One diffusion round (:299-302) with the filter read from DEVICE memory (the optimizer updates it there): filt_dev = fp32
 * [oc][ic][ky][kx] (16x16x3x3) followed by the 16 biases, as written by gwd_diffuse_filter_pack.  a_out = a_in +
 * gelu(instance_norm(conv(a_in))); raw (the convolution output) and stats (fp64 [B*heads][2] = sum, sum of squares per plane)
 * are kept by the caller for gwd_ref_diffuse_bwd.  Same kernels as gwd_ref_diffuse. */
int gwd_ref_diffuse_dev(const float* a_in, float* a_out, const float* filt_dev, float* raw, double* stats, int32_t B, int32_t heads,
                        int32_t P, int32_t R, void* stream);
/* out = add (optional) + conv3x3(a_in) with a device filter in the gwd_ref_diffuse_dev layout; stats_scratch fp64 [2*B*heads] */
int gwd_ref_diffuse_conv_dev(const float* a_in, const float* filt_dev, const float* add, float* out, double* stats_scratch,
                             int32_t B, int32_t heads, int32_t P, int32_t R, void* stream);
/* w_phys fp32 [9][16][16] (tap = kx*3+ky, the packed 3x3 layout of the flat parameter buffers) + bias [16] -> fwd / bwd filters
 * (2 320 floats each) in the gwd_ref_diffuse_dev layout; bwd = the adjoint filter (channels transposed, taps flipped, no bias) */
int gwd_diffuse_filter_pack(const float* w_phys, const float* bias, float* fwd, float* bwd, void* stream);
/* Backward of one diffusion round: g = d a_out fp32 [B,heads,P,R]; raw / stats / a_in from the forward; filt_bwd from
 * gwd_diffuse_filter_pack.  d_a_in = g + conv(d raw, adjoint filter); dw_phys ([9][16][16]) and db ([16]) are ACCUMULATED.
 * Workspaces: d_raw_ws fp32 [B,heads,P,R], stats2_ws fp64 [4*B*heads]. */
int gwd_ref_diffuse_bwd(const float* g, const float* raw, const double* stats, const float* a_in, const float* filt_bwd,
                        float* d_raw_ws, double* stats2_ws, float* d_a_in, float* dw_phys, float* db, int32_t B, int32_t heads,
                        int32_t P, int32_t R, void* stream);
/* ref_k = mu + exp(logsigma) * ref[:, :D]  (:281-288): ref fp32 [rows, ref_rs] -> out fp32 [rows, D] */
int gwd_ref_affine(const float* ref, int64_t ref_rs, const float* mu, const float* logsigma, float* out, int64_t rows, int32_t D,
                   void* stream);
/* d_kv fp32 [rows, 2D] = (d ref_k | d ref_v) -> d_ref bf16 [rows, 2D] = the output gradient of the ref_qk Linear;
 * dmu / dlogsigma fp32 [D] are accumulated */
int gwd_ref_affine_bwd(const float* d_kv, const float* ref, int64_t ref_rs, const float* logsigma, void* d_ref, float* dmu,
                       float* dlogsigma, int32_t rows, int32_t D, void* stream);
/* Backward of gwd_ref_requery: a = the diffused scores fp32 [B,heads,T,R], refv fp32 rows b*R+r (row stride ref_rs), d_qnew
 * bf16 rows b*T+t (row stride dq_rs) -> d_a fp32 [B,heads,T,R] (soft-max backward included), d_refv fp32 (row stride drv_rs) */
/* (d_refv / d_refk below are ACCUMULATED with atomics over the token tiles: zero them first) */
int gwd_ref_requery_bwd(const float* a, const float* refv, int64_t ref_rs, const void* d_qnew, int64_t dq_rs, float* d_a,
                        float* d_refv, int64_t drv_rs, int32_t B, int32_t T, int32_t heads, int32_t hd, int32_t R, float scale,
                        void* stream);
/* Backward of gwd_ref_scores: d_a fp32 [B,heads,T,R], refk fp32, q bf16 -> d_q bf16 (row stride dq_rs), d_refk fp32 */
int gwd_ref_scores_bwd(const float* d_a, const float* refk, int64_t ref_rs, const void* q, int64_t q_rs, void* d_q, int64_t dq_rs,
                       float* d_refk, int64_t drk_rs, int32_t B, int32_t T, int32_t heads, int32_t hd, int32_t R, float scale,
                       void* stream);
/* Adjoint of the feature part of gwd_line_ref_gather: d_win[row of point (b, r)] += d_ref[b, r]  (bf16, deterministic) */
int gwd_line_ref_scatter(const void* d_ref, int64_t dref_rs, const float* coords, int32_t R, void* d_win, int64_t win_rs, int32_t B,
                         int32_t H, int32_t W, int32_t ws, int32_t shift, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backbone backward helpers (stride-2 convolutions of torchvision's ResNet-50 v1.5 as wrapped by
 * src/models/backbone.py:58-92; FrozenBatchNorm2d :19-55 folded into the filters)
 * ------------------------------------------------------------------------------------------ */
/* y[b,i,j,:] = x[b,2i,2j,:]: bf16 [B,H,W,C] -> [B,ceil(H/2),ceil(W/2),C] */
int gwd_subsample2(const void* x, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* y[b,i,j,:] = add[b,i,j,:] (optional) + (i, j even ? s[b,i/2,j/2,:] : 0): the adjoint of gwd_subsample2 */
int gwd_zero_stuff2(const void* s, const void* add, void* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* g[i] *= scale[i] (gradient w.r.t. the folded filter -> gradient w.r.t. the parameter) */
int gwd_scale_rows(float* g, const float* scale, int64_t n, void* stream);
/* mirror[i] = bf16(p[i] * scale[i]) (the folded bf16 filter the forward kernels read) */
int gwd_fold_mirror(const float* p, const float* scale, void* mirror_bf16, int64_t n, void* stream);
/* Stride-2 3x3 convolution (padding 1) as im2col + Linear: x bf16 [B,H,W,C] -> col bf16 [B,ho,wo,9C] with
 * col[.., (ky*3+kx)*C + c] = x[2 oy + ky - 1, 2 ox + kx - 1, c] (0 outside), ho = (H-1)/2+1, wo = (W-1)/2+1.  Replaces the
 * cuDNN call behind conv2 of layer2.0 / layer3.0 / layer4.0 (torchvision ResNet-50 v1.5, src/models/backbone.py:58-92). */
int gwd_im2col3x3_s2(const void* x, void* col, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* adjoint of gwd_im2col3x3_s2: dx bf16 [B,H,W,C] = add (optional) + the taps of dcol bf16 [B,ho,wo,9C] that land on each pixel */
int gwd_col2im3x3_s2(const void* dcol, const void* add, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* Reference-line selection of the dense encoder (src/models/multiscale_transformerr.py:1165-1179: torch.topk over the RAW line
 * logit + gather + * 2 - 1): logits fp32 [B,Q,num_classes] (class 0 is ranked), lines fp32 [B,Q,line_dim] -> ids int64
 * [B,num_ref] (descending logit, ties by the lower index) and ref_xy fp32 [B, num_ref * points_per_line, 2] = the first
 * points_per_line (x, y) pairs of every selected line mapped to [-1, 1]. */
int gwd_select_lines(const float* logits, int32_t num_classes, const float* lines, int32_t line_dim, int32_t B, int32_t Q,
                     int32_t num_ref, int32_t points_per_line, float* ref_xy, int64_t* ids, void* stream);
/* element-wise train-mode dropout of a bf16 [n] buffer (n % 8 == 0): out = res (optional) + (keep ? x / (1 - p) : 0); seed: DEVICE
 * uint32, site: id of the call site.  The same call on a gradient is the backward (the mask is regenerated).  Replaces
 * nn.Dropout of src/models/transformer.py:149-162,212-233. */
int gwd_dropout(const void* x, const void* res, void* out, int64_t n, const uint32_t* seed, uint32_t site, float p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWD_B200_H_ */
