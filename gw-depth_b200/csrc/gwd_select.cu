// gwd_select.cu -- the discrete / reduction kernels of the path:
//   gwd_certain_sample  uncertainty ("reflection context") sampling     src/models/points/points_sample.py:291-364
//   gwd_match_cost      Hungarian-matcher cost matrix, block diagonal   src/models/matcher.py:52-71
//   gwd_depth_metrics   per-image depth error metrics                   src/util/metrics.py:197-218,
//                                                                       src/engine_glassrgbd.py:243-264
//   gwd_silog_sums      masked sums of the scale-invariant log loss      src/models/glassrgbd.py:366-374
// All inputs here are fp32 (selection inputs are kept in fp32 so that bf16 activations cannot flip a choice).
#include <algorithm>
#include "gwd_common.cuh"

namespace {

// -------------------------------------------------------------------------------------------------
// certain sample: one CTA per image
// -------------------------------------------------------------------------------------------------
constexpr int kCsThreads = 1024;
constexpr int kCsMaxK = 256;
constexpr int kCsMaxBins = 8;

struct CsParams {
  const float* small;  // [B, h, w]
  const float* large;  // [B, H, W]
  int h, w, H, W, K, nbins;
  float edges[kCsMaxBins + 1];
  float* coords;       // [B, K, 2]
  int32_t* index;      // [B, K]
};

__global__ void __launch_bounds__(kCsThreads) gwd_certain_sample_kernel(const CsParams p) {
  extern __shared__ float var[];  // [H*W]
  __shared__ int cnt[kCsMaxBins];
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int top[kCsMaxK];
  __shared__ int final_idx[kCsMaxK];
  const int b = blockIdx.x;
  const int HW = p.H * p.W;
  const float* sm = p.small + static_cast<int64_t>(b) * p.h * p.w;
  const float* lg = p.large + static_cast<int64_t>(b) * HW;
  if (threadIdx.x < kCsMaxBins) cnt[threadIdx.x] = 0;
  __syncthreads();
  // variance = (bilinear_align_corners(small) - large)^2, and the histogram of `large` over the depth bins
  const float ry = p.H > 1 ? static_cast<float>(p.h - 1) / (p.H - 1) : 0.f;
  const float rx = p.W > 1 ? static_cast<float>(p.w - 1) / (p.W - 1) : 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    int Y = i / p.W, X = i - Y * p.W;
    float fy = ry * Y, fx = rx * X;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    int y1 = min(y0 + 1, p.h - 1), x1 = min(x0 + 1, p.w - 1);
    float ly = fy - y0, lx = fx - x0;
    float up = (1.f - ly) * ((1.f - lx) * sm[y0 * p.w + x0] + lx * sm[y0 * p.w + x1]) +
               ly * ((1.f - lx) * sm[y1 * p.w + x0] + lx * sm[y1 * p.w + x1]);
    float d = lg[i];
    float dv = up - d;
    var[i] = dv * dv;
    for (int k = 0; k < p.nbins; ++k)
      if (d >= p.edges[k] && d < p.edges[k + 1]) atomicAdd(&cnt[k], 1);
  }
  __syncthreads();
  // per-bin quota n_k = min(floor(cnt_k / HW * K), cnt_k)   (fp32 arithmetic, :313-317)
  int n[kCsMaxBins];
  int nmax = 0, already = 0, nactive = 0;
  for (int k = 0; k < p.nbins; ++k) {
    float q = floorf((static_cast<float>(cnt[k]) / static_cast<float>(HW)) * static_cast<float>(p.K));
    n[k] = static_cast<int>(fminf(q, static_cast<float>(cnt[k])));
    nmax = max(nmax, n[k]);
    already += n[k];
    nactive += n[k] > 0;
  }
  const int need = nactive > 0 ? nmax : p.K;   // no populated bin: global top-K (:331-339)
  // iterative block arg-max: top[r] = index of the r-th largest variance (ties -> lower index)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < need; ++r) {
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float v = var[i];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < (blockDim.x >> 5) ? red_v[lane] : -1.f;
      bi = lane < (blockDim.x >> 5) ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { top[r] = bi; var[bi] = -2.f; }
    }
    __syncthreads();
  }
  // assemble the K indices (serial, tiny): every bin contributes the index-sorted prefix top[0..n_k)
  if (threadIdx.x == 0) {
    const int K = p.K;
    int len = 0;
    int seg_start[kCsMaxBins], seg_len[kCsMaxBins], nseg = 0;
    auto push_sorted_prefix = [&](int m) {
      int s = len;
      for (int i = 0; i < m && len < kCsMaxK; ++i) {
        int v = top[i], j = len++;
        while (j > s && final_idx[j - 1] > v) { final_idx[j] = final_idx[j - 1]; --j; }
        final_idx[j] = v;
      }
    };
    int remain;
    if (nactive > 0) {
      for (int k = 0; k < p.nbins; ++k)
        if (n[k] > 0) {
          seg_start[nseg] = len; seg_len[nseg] = n[k]; ++nseg;
          push_sorted_prefix(n[k]);
        }
      remain = K - already;
    } else {
      push_sorted_prefix(K);
      already = K;
      remain = 0;
    }
    if (remain > 0 && remain >= already) {            // repeat the whole list (:343-346)
      int times = remain / already + 1;
      for (int t = 1; t < times; ++t)
        for (int i = 0; i < already && len < kCsMaxK; ++i) final_idx[len++] = final_idx[i];
      remain = K - already * times;
    }
    if (remain > 0) {                                  // complement with the tail (:348-350)
      int start = len - remain;
      for (int i = 0; i < remain && len < kCsMaxK; ++i) final_idx[len++] = final_idx[start + i];
    }
    if (remain < 0 && nseg > 0) {                      // trim the largest bin (:351-355); unreachable in exact arithmetic
      int m = 0;
      for (int s = 1; s < nseg; ++s) if (seg_len[s] > seg_len[m]) m = s;
      int cut = -remain, dst = 0;
      for (int s = 0; s < nseg; ++s) {
        int keep = seg_len[s] - (s == m ? cut : 0);
        for (int i = 0; i < keep; ++i) final_idx[dst++] = final_idx[seg_start[s] + i];
      }
      len = dst;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    int idx = final_idx[k];
    int row = idx / p.W, col = idx - row * p.W;
    p.index[static_cast<int64_t>(b) * p.K + k] = idx;
    p.coords[(static_cast<int64_t>(b) * p.K + k) * 2 + 0] = (static_cast<float>(col) / p.W) * 2.f - 1.f;
    p.coords[(static_cast<int64_t>(b) * p.K + k) * 2 + 1] = (static_cast<float>(row) / p.H) * 2.f - 1.f;
  }
}

// -------------------------------------------------------------------------------------------------
// matcher cost: one warp per (image, query); lanes stride over that image's targets.
// cost[off_b*Q + q*T_b + t] = w_line * sum_d |line[b,q,d] - tgt[t,d]| - w_class * softmax(logits[b,q])[label[t]]
// The per-query minimum (a warp-shuffle reduction) is also returned: it is the row reduction step of the
// assignment solver that consumes the matrix.
// -------------------------------------------------------------------------------------------------
__global__ void gwd_match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ lines,
                                      const float* __restrict__ tgt_lines, const int64_t* __restrict__ tgt_labels,
                                      const int32_t* __restrict__ tgt_offsets, int B, int Q, int ncls, int D, float w_class,
                                      float w_line, float* __restrict__ cost, float* __restrict__ row_min) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wid >= B * Q) return;
  int b = wid / Q, q = wid - b * Q;
  int t0 = tgt_offsets[b], T = tgt_offsets[b + 1] - t0;
  const float* lg = logits + static_cast<int64_t>(wid) * ncls;
  float mx = lg[0];
  for (int c = 1; c < ncls; ++c) mx = fmaxf(mx, lg[c]);
  float den = 0.f;
  for (int c = 0; c < ncls; ++c) den += expf(lg[c] - mx);
  float ln[8];
  for (int d = 0; d < D; ++d) ln[d] = lines[static_cast<int64_t>(wid) * D + d];
  float* out = cost + static_cast<int64_t>(t0) * Q + static_cast<int64_t>(q) * T;
  float best = INFINITY;
  for (int t = lane; t < T; t += 32) {
    const float* tl = tgt_lines + static_cast<int64_t>(t0 + t) * D;
    float l1 = 0.f;
    for (int d = 0; d < D; ++d) l1 += fabsf(ln[d] - tl[d]);
    int label = tgt_labels ? static_cast<int>(tgt_labels[t0 + t]) : 0;
    float prob = expf(lg[label] - mx) / den;
    float c = w_line * l1 + w_class * (-prob);
    out[t] = c;
    best = fminf(best, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if (lane == 0 && row_min) row_min[wid] = best;
}

// -------------------------------------------------------------------------------------------------
// depth metrics: sums[b][0..9] (double) over valid pixels, then the 9 metrics
// -------------------------------------------------------------------------------------------------
__global__ void gwd_depth_metric_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int64_t HW,
                                             float min_d, float max_d, double* __restrict__ sums) {
  const int b = blockIdx.y;
  const float* pp = pred + static_cast<int64_t>(b) * HW;
  const float* gg = gt + static_cast<int64_t>(b) * HW;
  double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float g = gg[i], q = pp[i];
    // engine_glassrgbd.py:249-252: clamp, then Inf -> max, NaN -> min
    if (isnan(q)) q = min_d;
    else if (q < min_d) q = min_d;
    else if (q > max_d) q = max_d;
    if (!(g > min_d && g < max_d)) continue;
    float th = fmaxf(g / q, q / g);
    float lg = logf(g), lq = logf(q);
    float diff = g - q;
    acc[0] += 1.0;
    acc[1] += th < 1.25f;
    acc[2] += th < 1.25f * 1.25f;
    acc[3] += th < 1.25f * 1.25f * 1.25f;
    acc[4] += static_cast<double>(diff * diff);
    acc[5] += static_cast<double>((lg - lq) * (lg - lq));
    acc[6] += static_cast<double>(fabsf(diff) / g);
    acc[7] += static_cast<double>(diff * diff / g);
    acc[8] += static_cast<double>(lq - lg);
    acc[9] += static_cast<double>(fabsf(log10f(q) - log10f(g)));
  }
  __shared__ double red[10][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < 10; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double v = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    atomicAdd(&sums[b * 10 + threadIdx.x], v);
  }
}

__global__ void gwd_depth_metric_final_kernel(const double* __restrict__ sums, int B, double* __restrict__ out) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* s = sums + b * 10;
  double n = s[0] > 0 ? s[0] : 1.0;
  double me = s[8] / n;
  double* o = out + b * 9;
  // silog uses sum(err^2) = sum((log g - log q)^2)
  o[0] = sqrt(fmax(s[5] / n - me * me, 0.0)) * 100.0;  // silog
  o[1] = s[6] / n;                                     // abs_rel
  o[2] = s[9] / n;                                     // log10
  o[3] = sqrt(s[4] / n);                               // rms
  o[4] = s[7] / n;                                     // sq_rel
  o[5] = sqrt(s[5] / n);                               // log_rms
  o[6] = s[1] / n;                                     // d1
  o[7] = s[2] / n;                                     // d2
  o[8] = s[3] / n;                                     // d3
}

// masked silog sums per image-batch: sums[0]=count, [1]=sum d, [2]=sum d^2, d = f(pred) - f(gt) over valid gt,
// with gt / validity taken at the nearest-resized location (engine_glassrgbd.py:74-80)
__global__ void gwd_silog_sums_kernel(const float* __restrict__ pred, int B, int h, int w, const float* __restrict__ gt,
                                      int H, int W, float lo, float hi, int log_only, double* __restrict__ sums) {
  double a0 = 0, a1 = 0, a2 = 0;
  int64_t total = static_cast<int64_t>(B) * h * w;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int x = i % w;
    int y = (i / w) % h;
    int b = i / (static_cast<int64_t>(w) * h);
    int sy = min(static_cast<int>((static_cast<int64_t>(y) * H) / h), H - 1);
    int sx = min(static_cast<int>((static_cast<int64_t>(x) * W) / w), W - 1);
    float g = gt[(static_cast<int64_t>(b) * H + sy) * W + sx];
    if (!(g >= lo && g < hi)) continue;
    float q = pred[i];
    float d = log_only ? (logf(q) - logf(g)) : ((q + logf(q)) - (g + logf(g)));
    a0 += 1.0; a1 += d; a2 += static_cast<double>(d) * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], a0);
    atomicAdd(&sums[1], a1);
    atomicAdd(&sums[2], a2);
  }
}


// segmentation evaluation (src/engine_glassrgbd.py:232-240 + src/util/metrics.py:43-78): per-pixel argmax over the class
// logits (first maximum, as torch.argmax) and the [C, C] confusion counts conf[gt][pred] over pixels with gt != ignore,
// accumulated into an int64 matrix on the device (the reference copies both maps to the host and runs np.bincount).
__global__ void __launch_bounds__(256)
gwd_seg_confusion_kernel(const float* __restrict__ logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride,
                         const int64_t* __restrict__ gt, int64_t HW, int64_t total, int C, int ignore,
                         unsigned long long* __restrict__ conf) {
  __shared__ unsigned int hist[64];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t g = gt[i];
    if (g == ignore || g < 0 || g >= C) continue;
    const float* px = logits + (i / HW) * image_stride + (i % HW) * pixel_stride;
    int best = 0;
    float bv = px[0];
    for (int c = 1; c < C; ++c) {
      const float v = px[c * class_stride];
      if (v > bv) { bv = v; best = c; }
    }
    atomicAdd(&hist[static_cast<int>(g) * C + best], 1u);
  }
  __syncthreads();
  if (threadIdx.x < C * C && hist[threadIdx.x]) atomicAdd(&conf[threadIdx.x], static_cast<unsigned long long>(hist[threadIdx.x]));
}

// ------------------------------------------------------------------------------------------------
// reference-line selection of the dense encoder (multiscale_transformerr.py:1165-1179): the num_ref lines with the largest
// RAW line logit (class 0), in descending order (torch.topk, ties by the lower index), and their points mapped to [-1,1].
// One CTA per image, thread t ranks query t by counting the queries that beat it (Q = 100: 10 k comparisons).
// ------------------------------------------------------------------------------------------------
__global__ void gwd_select_lines_kernel(const float* __restrict__ logits, int ncls, const float* __restrict__ lines, int D, int Q,
                                        int num_ref, int npts, float* __restrict__ ref_xy, int64_t* __restrict__ ids) {
  extern __shared__ float sl[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < Q; t += blockDim.x) sl[t] = logits[(static_cast<int64_t>(b) * Q + t) * ncls];
  __syncthreads();
  for (int t = threadIdx.x; t < Q; t += blockDim.x) {
    const float v = sl[t];
    int rank = 0;
    for (int j = 0; j < Q; ++j) rank += (sl[j] > v) || (sl[j] == v && j < t);
    if (rank < num_ref) {
      ids[static_cast<int64_t>(b) * num_ref + rank] = t;
      const float* ln = lines + (static_cast<int64_t>(b) * Q + t) * D;
      float* o = ref_xy + (static_cast<int64_t>(b) * num_ref + rank) * npts * 2;
      for (int i = 0; i < npts * 2; ++i) o[i] = ln[i] * 2.f - 1.f;
    }
  }
}

}  // namespace

extern "C" int gwd_certain_sample(const float* pred_small, int32_t h, int32_t w, const float* pred_large, int32_t H,
                                  int32_t W, int32_t B, int32_t K, const float* edges_host, int32_t nbins, float* coords,
                                  int32_t* index, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(pred_small && pred_large && coords && index && edges_host, "gwd_certain_sample: null pointer");
  GWD_CHECK_ARG(K > 0 && K <= kCsMaxK && nbins > 0 && nbins <= kCsMaxBins, "gwd_certain_sample: K/nbins out of range");
  GWD_CHECK_ARG(static_cast<int64_t>(H) * W >= K, "gwd_certain_sample: fewer pixels than samples");
  CsParams p;
  p.small = pred_small; p.large = pred_large; p.h = h; p.w = w; p.H = H; p.W = W; p.K = K; p.nbins = nbins;
  for (int i = 0; i <= nbins; ++i) p.edges[i] = edges_host[i];
  p.coords = coords; p.index = index;
  size_t smem = static_cast<size_t>(H) * W * sizeof(float);
  GWD_CHECK_ARG(smem <= 200 * 1024, "gwd_certain_sample: map %dx%d too large for one CTA", H, W);
  if (smem > 40 * 1024) {
    static bool configured = false;
    if (!configured) {
      GWD_CUDA(cudaFuncSetAttribute(gwd_certain_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
  }
  gwd_certain_sample_kernel<<<B, kCsThreads, smem, stream>>>(p);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_select_lines(const float* logits, int32_t num_classes, const float* lines, int32_t line_dim, int32_t B, int32_t Q,
                                int32_t num_ref, int32_t points_per_line, float* ref_xy, int64_t* ids, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(logits && lines && ref_xy && ids && B > 0 && Q > 0, "gwd_select_lines: null pointer / empty");
  GWD_CHECK_ARG(num_ref > 0 && num_ref <= Q && points_per_line > 0 && 2 * points_per_line <= line_dim && Q <= 8192,
                "gwd_select_lines: needs num_ref <= Q <= 8192 and 2 * points_per_line <= line_dim");
  gwd_select_lines_kernel<<<static_cast<unsigned>(B), 128, sizeof(float) * Q, stream>>>(logits, num_classes, lines, line_dim, Q, num_ref,
                                                                                      points_per_line, ref_xy, ids);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_match_cost(const float* logits, const float* lines, const float* tgt_lines, const int64_t* tgt_labels,
                              const int32_t* tgt_offsets, int32_t B, int32_t Q, int32_t num_classes, int32_t line_dim,
                              float w_class, float w_line, float* cost, float* row_min, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(logits && lines && tgt_lines && tgt_offsets && cost, "gwd_match_cost: null pointer");
  GWD_CHECK_ARG(line_dim > 0 && line_dim <= 8 && num_classes >= 1, "gwd_match_cost: bad dims");
  int64_t warps = static_cast<int64_t>(B) * Q;
  gwd_match_cost_kernel<<<static_cast<unsigned>(gwd_ceil_div(warps * 32, 128)), 128, 0, stream>>>(
      logits, lines, tgt_lines, tgt_labels, tgt_offsets, B, Q, num_classes, line_dim, w_class, w_line, cost, row_min);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_depth_metrics(const float* pred, const float* gt, int32_t B, int64_t HW, float min_depth,
                                 float max_depth, double* workspace, double* metrics, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(pred && gt && workspace && metrics && B > 0 && HW > 0, "gwd_depth_metrics: bad argument");
  GWD_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * 10 * B, stream));
  int bx = static_cast<int>(std::min<int64_t>(gwd_ceil_div(HW, 256 * 4), 4 * gwd_num_sms()));
  gwd_depth_metric_sums_kernel<<<dim3(bx, B), 256, 0, stream>>>(pred, gt, HW, min_depth, max_depth, workspace);
  GWD_LAUNCHED();
  gwd_depth_metric_final_kernel<<<static_cast<unsigned>(gwd_ceil_div(B, 64)), 64, 0, stream>>>(workspace, B, metrics);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_silog_sums(const float* pred, int32_t B, int32_t h, int32_t w, const float* gt, int32_t H, int32_t W,
                              float lo, float hi, int32_t log_only, double* sums3, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(pred && gt && sums3, "gwd_silog_sums: null pointer");
  GWD_CUDA(cudaMemsetAsync(sums3, 0, sizeof(double) * 3, stream));
  int64_t total = static_cast<int64_t>(B) * h * w;
  int bx = static_cast<int>(std::min<int64_t>(gwd_ceil_div(total, 256 * 4), 4 * gwd_num_sms()));
  gwd_silog_sums_kernel<<<bx, 256, 0, stream>>>(pred, B, h, w, gt, H, W, lo, hi, log_only, sums3);
  GWD_LAUNCHED();
  return GWD_OK;
}

extern "C" int gwd_seg_confusion(const float* logits, int64_t pixel_stride, int64_t class_stride, int64_t image_stride,
                                 const int64_t* gt, int32_t B, int64_t HW, int32_t num_classes, int32_t ignore_index,
                                 int64_t* confusion, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GWD_CHECK_ARG(logits && gt && confusion && B > 0 && HW > 0, "gwd_seg_confusion: bad argument");
  GWD_CHECK_ARG(num_classes >= 2 && num_classes <= 8, "gwd_seg_confusion: 2..8 classes");
  const int64_t total = static_cast<int64_t>(B) * HW;
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>(gwd_ceil_div(total, 256 * 8), 8 * gwd_num_sms()));
  gwd_seg_confusion_kernel<<<grid, 256, 0, stream>>>(logits, pixel_stride, class_stride, image_stride, gt, HW, total, num_classes,
                                                     ignore_index, reinterpret_cast<unsigned long long*>(confusion));
  GWD_LAUNCHED();
  return GWD_OK;
}
