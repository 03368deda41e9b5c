// Losses / metrics adjacent to the view-synthesis path (include/vsl.h "Depth metrics"):
//   compute_depth_errors            layers.py:335-353
//   Trainer.compute_depth_losses    trainer.py:688-716 (up-sample to the ground-truth size, Garg/Eigen crop, median
//                                   scaling, clamp, then the seven error metrics)
//   SLlog                           layers.py:32-56 (scale-invariant log loss of the GAN prior, trainer.py:565-583)
// All reductions are fixed-order (per-block partials in fp64, summed by the last block), so results are
// reproducible; the medians are exact (radix select on the float bits: torch.median's lower median).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vsl.h"
#include "vsl_math.cuh"

namespace vsl {

extern thread_local int g_last_cuda_error;  // defined in vsl_fused.cu
#define VSL_M_OK(expr)                                                \
  do {                                                                \
    cudaError_t e__ = (expr);                                         \
    if (e__ != cudaSuccess) { g_last_cuda_error = (int)e__; return VSL_ERR_CUDA; } \
  } while (0)

constexpr int kMT = 256;          // threads per block
constexpr int kMaxBlocks = 592;   // 4 per SM: grid-stride loops
constexpr int kTerms = 8;

struct MetricsWs {                // layout of the caller's workspace
  double partial[kMaxBlocks][kTerms];
  unsigned hist[2][256];
  unsigned prefix[2], rank[2];
  unsigned counter, n_valid;
  float ratio, pad;
};

// Where the (gt, pred) pairs come from: flat arrays, or Trainer.compute_depth_losses' masked / up-sampled pixels
struct PairSource {
  const float* gt;
  const float* pred;
  size_t n;             // flat: element count; depth-losses: B * gt_h * gt_w
  int depth_losses;     // 0: flat arrays (every element valid)
  int h, w, gt_h, gt_w; // depth_pred [B,1,h,w], depth_gt [B,1,gt_h,gt_w]
  int cy0, cy1, cx0, cx1;  // crop (trainer.py:701-703), half-open
  float lo, hi;         // clamp (trainer.py:694, :711)
  float scale_h, scale_w;
};

// the pair at flat index i (gt, un-scaled pred); false if masked out
__device__ __forceinline__ bool fetch_pair(const PairSource& s, size_t i, float& gt, float& pred) {
  if (!s.depth_losses) {
    gt = s.gt[i]; pred = s.pred[i];
    return true;
  }
  const int plane = s.gt_h * s.gt_w;
  const int b = (int)(i / plane), r = (int)(i - (size_t)b * plane);
  const int y = r / s.gt_w, x = r - y * s.gt_w;
  if (y < s.cy0 || y >= s.cy1 || x < s.cx0 || x >= s.cx1) return false;
  gt = s.gt[i];
  if (!(gt > 0.f)) return false;
  // F.interpolate(depth_pred, [gt_h, gt_w], "bilinear", align_corners=False), the arithmetic of the disp up-sample
  const float v = upsample_disp(s.pred + (size_t)b * s.h * s.w, s.h, s.w, s.scale_h, s.scale_w, false, y, x, 0);
  pred = fminf(fmaxf(v, s.lo), s.hi);
  return true;
}

__device__ __forceinline__ double block_sum_d(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double r = 0.0;
#pragma unroll
  for (int i = 0; i < kMT / 32; ++i) r += scratch[i];
  return r;
}

// ---- exact median by radix select: one 8-bit digit per pass, most significant first -----------------------
// (values are positive floats, so their bit patterns order like the values)
__global__ void __launch_bounds__(kMT) k_median_hist(const PairSource s, MetricsWs* ws, int shift) {
  __shared__ unsigned h[2][256];
  h[0][threadIdx.x] = 0u; h[1][threadIdx.x] = 0u;
  __syncthreads();
  const unsigned p0 = ws->prefix[0], p1 = ws->prefix[1];
  const unsigned himask = shift >= 24 ? 0u : (0xffffffffu << (shift + 8));
  for (size_t i = (size_t)blockIdx.x * kMT + threadIdx.x; i < s.n; i += (size_t)gridDim.x * kMT) {
    float gt, pred;
    if (!fetch_pair(s, i, gt, pred)) continue;
    const unsigned a = __float_as_uint(gt), b = __float_as_uint(pred);
    if ((a & himask) == p0) atomicAdd(&h[0][(a >> shift) & 255u], 1u);
    if ((b & himask) == p1) atomicAdd(&h[1][(b >> shift) & 255u], 1u);
  }
  __syncthreads();
  if (h[0][threadIdx.x]) atomicAdd(&ws->hist[0][threadIdx.x], h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(&ws->hist[1][threadIdx.x], h[1][threadIdx.x]);
}
__global__ void __launch_bounds__(64) k_median_select(MetricsWs* ws, int shift, int first, int last) {
  // warp q (0: gt, 1: pred): lane l owns bins 8l .. 8l+7; exclusive scan over the lanes finds the digit whose
  // cumulative count passes the rank
  __shared__ unsigned pref[2];
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned c[8], mine = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = ws->hist[q][8 * lane + j]; mine += c[j]; }
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned k = first ? (total ? (total - 1) / 2 : 0) : ws->rank[q];  // torch.median: the lower of the two middle elements
  if (first && q == 0 && lane == 0) ws->n_valid = total;
  const unsigned before = incl - mine;
  const bool here = total == 0 ? lane == 0 : (k >= before && k < incl);
  if (here) {
    unsigned r = k - before, d = 0;
    for (; d < 7; ++d) {
      if (r < c[d]) break;
      r -= c[d];
    }
    ws->rank[q] = r;
    const unsigned p = ws->prefix[q] | ((unsigned)(8 * lane + d) << shift);
    ws->prefix[q] = p;
    pref[q] = p;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) ws->hist[q][8 * lane + j] = 0u;
  __syncthreads();
  if (last && threadIdx.x == 0) ws->ratio = __uint_as_float(pref[0]) / __uint_as_float(pref[1]);  // trainer.py:709
}

// ---- the seven metrics -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMT) k_depth_errors(const PairSource s, MetricsWs* ws, float* __restrict__ out) {
  __shared__ double scratch[kMT / 32];
  __shared__ bool is_last;
  double acc[kTerms] = {0, 0, 0, 0, 0, 0, 0, 0};  // count, abs_rel, sq_rel, se, sle, a1, a2, a3
  const float ratio = s.depth_losses ? ws->ratio : 1.0f;
  const float t1 = 1.25f, t2 = (float)(1.25 * 1.25), t3 = (float)(1.25 * 1.25 * 1.25);
  for (size_t i = (size_t)blockIdx.x * kMT + threadIdx.x; i < s.n; i += (size_t)gridDim.x * kMT) {
    float gt, pred;
    if (!fetch_pair(s, i, gt, pred)) continue;
    if (s.depth_losses) pred = fminf(fmaxf(pred * ratio, s.lo), s.hi);  // trainer.py:709-711
    const float thresh = fmaxf(gt / pred, pred / gt);
    const float diff = gt - pred, dl = logf(gt) - logf(pred);
    acc[0] += 1.0;
    acc[1] += (double)(fabsf(diff) / gt);
    acc[2] += (double)(diff * diff / gt);
    acc[3] += (double)(diff * diff);
    acc[4] += (double)(dl * dl);
    acc[5] += thresh < t1 ? 1.0 : 0.0;
    acc[6] += thresh < t2 ? 1.0 : 0.0;
    acc[7] += thresh < t3 ? 1.0 : 0.0;
  }
#pragma unroll
  for (int k = 0; k < kTerms; ++k) {
    const double v = block_sum_d(acc[k], scratch);
    if (threadIdx.x == 0) ws->partial[blockIdx.x][k] = v;
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // the blocks' partials: thread t adds blocks t, t + 256, ... (independent loads), then a fixed-order block sum
  __shared__ double total[kTerms];
#pragma unroll
  for (int k = 0; k < kTerms; ++k) {
    double v = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kMT) v += __ldcg(&ws->partial[b][k]);
    v = block_sum_d(v, scratch);
    if (threadIdx.x == 0) total[k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double* scratch = total;
    const double n = scratch[0];
    out[0] = (float)(scratch[1] / n);        // abs_rel
    out[1] = (float)(scratch[2] / n);        // sq_rel
    out[2] = (float)sqrt(scratch[3] / n);    // rmse
    out[3] = (float)sqrt(scratch[4] / n);    // rmse_log
    out[4] = (float)(scratch[5] / n);        // a1
    out[5] = (float)(scratch[6] / n);        // a2
    out[6] = (float)(scratch[7] / n);        // a3
    ws->counter = 0u;
  }
}

// ---- SLlog (layers.py:32-56) --------------------------------------------------------------------------------
__device__ __forceinline__ bool sllog_term(float fake, float real, float& d) {
  if (real <= 0.f || fake <= 0.f) { d = 0.f; return false; }  // both set to 1: log 1 - log 1
  d = logf(real) - logf(fake);
  return true;
}
__global__ void __launch_bounds__(kMT) k_sllog_fwd(size_t n, const float* __restrict__ fake, const float* __restrict__ real,
                                                   MetricsWs* ws, float* __restrict__ loss, float* __restrict__ stats) {
  __shared__ double scratch[kMT / 32];
  __shared__ bool is_last;
  double cnt = 0.0, sd = 0.0, sdd = 0.0;
  for (size_t i = (size_t)blockIdx.x * kMT + threadIdx.x; i < n; i += (size_t)gridDim.x * kMT) {
    const float r = real[i];
    float d;
    sllog_term(fake[i], r, d);
    cnt += r > 0.f ? 1.0 : 0.0;  // N counts real > 0 only (layers.py:44)
    sd += (double)d;
    sdd += (double)(d * d);
  }
  const double a = block_sum_d(cnt, scratch), b = block_sum_d(sd, scratch), c = block_sum_d(sdd, scratch);
  if (threadIdx.x == 0) { ws->partial[blockIdx.x][0] = a; ws->partial[blockIdx.x][1] = b; ws->partial[blockIdx.x][2] = c; }
  __threadfence();
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t0 = 0.0, t1 = 0.0, t2 = 0.0;
  for (unsigned k = threadIdx.x; k < gridDim.x; k += kMT) {
    t0 += __ldcg(&ws->partial[k][0]); t1 += __ldcg(&ws->partial[k][1]); t2 += __ldcg(&ws->partial[k][2]);
  }
  const double N = block_sum_d(t0, scratch), S1 = block_sum_d(t1, scratch), S2 = block_sum_d(t2, scratch);
  if (threadIdx.x != 0) return;
  const double mean = S1 / N, l = sqrt(S2 / N - mean * mean);
  *loss = (float)l;
  stats[0] = (float)N; stats[1] = (float)mean; stats[2] = (float)l;
  ws->counter = 0u;
}
__global__ void __launch_bounds__(kMT) k_sllog_bwd(size_t n, const float* __restrict__ fake, const float* __restrict__ real,
                                                   const float* __restrict__ stats, const float* __restrict__ gl,
                                                   float* __restrict__ gfake, float* __restrict__ greal) {
  // loss = sqrt(S2/N - (S1/N)^2), d_i = log real_i - log fake_i  ->  d loss / d d_i = (d_i - mean) / (N loss)
  const float c = *gl / (stats[0] * stats[2]), mean = stats[1];
  for (size_t i = (size_t)blockIdx.x * kMT + threadIdx.x; i < n; i += (size_t)gridDim.x * kMT) {
    const float f = fake[i], r = real[i];
    float d;
    const bool live = sllog_term(f, r, d);   // masked entries were overwritten by the constant 1: no gradient
    const float g = live ? c * (d - mean) : 0.f;
    if (gfake) gfake[i] = live ? -g / f : 0.f;
    if (greal) greal[i] = live ? g / r : 0.f;
  }
}

static unsigned blocks_for_n(size_t n) {
  size_t b = (n + kMT - 1) / kMT;
  return (unsigned)(b < 1 ? 1 : (b > kMaxBlocks ? kMaxBlocks : b));
}

}  // namespace vsl

using namespace vsl;

extern "C" {

size_t vsl_metrics_workspace_bytes(void) { return sizeof(MetricsWs); }

int vsl_depth_errors(size_t n, const float* gt, const float* pred, float* out7, void* workspace, size_t workspace_bytes,
                     void* stream) {
  if (n < 1) return VSL_ERR_BAD_DESC;
  if (!gt || !pred || !out7 || !workspace) return VSL_ERR_NULL_POINTER;
  if (workspace_bytes < sizeof(MetricsWs)) return VSL_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 7u) != 0) return VSL_ERR_MISALIGNED;
  MetricsWs* ws = (MetricsWs*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  VSL_M_OK(cudaMemsetAsync(&ws->counter, 0, sizeof(unsigned), st));
  PairSource s = {};
  s.gt = gt; s.pred = pred; s.n = n;
  k_depth_errors<<<blocks_for_n(n), kMT, 0, st>>>(s, ws, out7);
  VSL_M_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_depth_losses(int batch, int height, int width, int gt_height, int gt_width, const int crop[4], float clamp_min,
                     float clamp_max, const float* depth_pred, const float* depth_gt, float* out7, void* workspace,
                     size_t workspace_bytes, void* stream) {
  if (batch < 1 || height < 1 || width < 1 || gt_height < 1 || gt_width < 1 || !crop) return VSL_ERR_BAD_DESC;
  if (!depth_pred || !depth_gt || !out7 || !workspace) return VSL_ERR_NULL_POINTER;
  if (workspace_bytes < sizeof(MetricsWs)) return VSL_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 7u) != 0) return VSL_ERR_MISALIGNED;
  MetricsWs* ws = (MetricsWs*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // hist, prefix, rank, counter, n_valid, ratio
  VSL_M_OK(cudaMemsetAsync(&ws->hist[0][0], 0, sizeof(MetricsWs) - offsetof(MetricsWs, hist), st));
  PairSource s = {};
  s.gt = depth_gt; s.pred = depth_pred; s.n = (size_t)batch * gt_height * gt_width; s.depth_losses = 1;
  s.h = height; s.w = width; s.gt_h = gt_height; s.gt_w = gt_width;
  s.cy0 = crop[0] < 0 ? 0 : crop[0]; s.cy1 = crop[1] > gt_height ? gt_height : crop[1];
  s.cx0 = crop[2] < 0 ? 0 : crop[2]; s.cx1 = crop[3] > gt_width ? gt_width : crop[3];
  s.lo = clamp_min; s.hi = clamp_max;
  s.scale_h = (float)height / (float)gt_height; s.scale_w = (float)width / (float)gt_width;
  const unsigned nb = blocks_for_n(s.n);
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    k_median_hist<<<nb, kMT, 0, st>>>(s, ws, shift);
    k_median_select<<<1, 64, 0, st>>>(ws, shift, pass == 0, pass == 3);
  }
  k_depth_errors<<<nb, kMT, 0, st>>>(s, ws, out7);
  VSL_M_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_sllog_forward(size_t n, const float* fake, const float* real, float* loss, float* stats3, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (n < 1) return VSL_ERR_BAD_DESC;
  if (!fake || !real || !loss || !stats3 || !workspace) return VSL_ERR_NULL_POINTER;
  if (workspace_bytes < sizeof(MetricsWs)) return VSL_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 7u) != 0) return VSL_ERR_MISALIGNED;
  MetricsWs* ws = (MetricsWs*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  VSL_M_OK(cudaMemsetAsync(&ws->counter, 0, sizeof(unsigned), st));
  k_sllog_fwd<<<blocks_for_n(n), kMT, 0, st>>>(n, fake, real, ws, loss, stats3);
  VSL_M_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_sllog_backward(size_t n, const float* fake, const float* real, const float* stats3, const float* grad_loss,
                       float* grad_fake, float* grad_real, void* stream) {
  if (n < 1) return VSL_ERR_BAD_DESC;
  if (!fake || !real || !stats3 || !grad_loss) return VSL_ERR_NULL_POINTER;
  k_sllog_bwd<<<blocks_for_n(n), kMT, 0, (cudaStream_t)stream>>>(n, fake, real, stats3, grad_loss, grad_fake, grad_real);
  VSL_M_OK(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
