// Gradient with respect to the source images (include/vsl.h, "Gradient with respect to the SOURCE IMAGES").
//
// The reference's autograd produces d loss / d inputs[("color", f, 0)] whenever that tensor requires grad: through
// F.grid_sample's backward for the warped candidates (trainer.py:534-537) and through the identity reprojection
// losses (trainer.py:620-633).  The fused kernel does not carry it (the images are inputs, not leaves, in training);
// this file holds the two kernels the host layer composes it from, next to vsl_reprojection_loss_backward:
//   k_source_grad_upstream   which candidate won the per-pixel minimum (trainer.py:663-666) -> what each
//                            candidate's reprojection loss receives from the loss dict (trainer.py:672-685)
//   k_grid_sample_bwd_source the bilinear scatter (ATen grid_sampler_2d_backward, border padding,
//                            align_corners=True): tile-local accumulation in shared memory, then one
//                            red.global.add.f32 per touched source pixel; far taps go to global memory directly.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsl.h"
#include "vsl_math.cuh"

namespace vsl {

extern thread_local int g_last_cuda_error;  // defined in vsl_fused.cu
#define VSL_S_OK(expr)                                                \
  do {                                                                \
    cudaError_t e__ = (expr);                                         \
    if (e__ != cudaSuccess) { g_last_cuda_error = (int)e__; return VSL_ERR_CUDA; } \
  } while (0)

struct UpstreamParams {
  const float* up;                              // [2S+1]
  const unsigned char* winner[VSL_MAX_SCALES];  // [B,H,W]
  float* up_identity;                           // [F][B,H,W] or null
  float* up_warped[VSL_MAX_SCALES];             // [F][B,H,W] each
  int S, F, avg;
  size_t n;                                     // B*H*W
  float wpix;
};

__global__ void __launch_bounds__(256) k_source_grad_upstream(const UpstreamParams p) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.n) return;
  float idacc[VSL_MAX_SRC] = {0.f, 0.f, 0.f, 0.f};
  const float tot = p.up[2 * p.S] / (float)p.S;
  for (int s = 0; s < p.S; ++s) {
    const float a = (p.up[s] + p.up[p.S + s] + tot) * p.wpix;
    const int w = p.winner[s][i];
    for (int f = 0; f < p.F; ++f) {
      float vw, vi;
      if (p.avg) {  // trainer.py:629-630, 649-650: the candidates are means over the frames
        vw = w == 1 ? a / (float)p.F : 0.f;
        vi = w == 0 ? a / (float)p.F : 0.f;
      } else {
        vw = w == p.F + f ? a : 0.f;
        vi = w == f ? a : 0.f;
      }
      p.up_warped[s][(size_t)f * p.n + i] = vw;
      idacc[f] += vi;
    }
  }
  if (p.up_identity)
    for (int f = 0; f < p.F; ++f) p.up_identity[(size_t)f * p.n + i] = idacc[f];
}

// One CTA: a 32 x 8 tile of TARGET pixels, one per thread.  Shared window of the SOURCE image around the tile
// (+kPad on every side, 3 channels); under the near-identity poses of this loss almost every tap lands in it.
constexpr int kSTW = 32, kSTH = 8, kPad = 8;
constexpr int kWinW = kSTW + 2 * kPad, kWinH = kSTH + 2 * kPad;

__device__ __forceinline__ void scatter_tap(float* __restrict__ win, float* __restrict__ gsrc, int HW, int W, int wx0, int wy0,
                                            int x, int y, const float g[3], float wgt) {
  const int lx = x - wx0, ly = y - wy0;
  if (lx >= 0 && lx < kWinW && ly >= 0 && ly < kWinH) {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(win + (c * kWinH + ly) * kWinW + lx, g[c] * wgt);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(gsrc + (size_t)c * HW + (size_t)y * W + x, g[c] * wgt);
  }
}

__global__ void __launch_bounds__(kSTW * kSTH) k_grid_sample_bwd_source(int H, int W, const float* __restrict__ grid,
                                                                       const float* __restrict__ gpred,
                                                                       float* __restrict__ gsrc) {
  __shared__ float win[3 * kWinH * kWinW];
  const int tid = threadIdx.y * kSTW + threadIdx.x;
  for (int k = tid; k < 3 * kWinH * kWinW; k += kSTW * kSTH) win[k] = 0.f;
  __syncthreads();
  const int b = blockIdx.z, HW = H * W;
  const int x = blockIdx.x * kSTW + threadIdx.x, y = blockIdx.y * kSTH + threadIdx.y;
  const int wx0 = blockIdx.x * kSTW - kPad, wy0 = blockIdx.y * kSTH - kPad;
  float* gs = gsrc + (size_t)b * 3 * HW;
  if (x < W && y < H) {
    const size_t o = (size_t)b * HW + (size_t)y * W + x;
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = gpred[(size_t)b * 3 * HW + (size_t)c * HW + (size_t)y * W + x];
    if (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f) {
      // un-normalise + border clip exactly like the forward (GridSampler.cuh:23-31, 55-57; project_pixel)
      const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
      float ix = mul_rn(mul_rn(add_rn(grid[2 * o], 1.0f), 0.5f), wm1);
      float iy = mul_rn(mul_rn(add_rn(grid[2 * o + 1], 1.0f), 0.5f), hm1);
      ix = fminf(wm1, fmaxf(ix, 0.f));
      iy = fminf(hm1, fmaxf(iy, 0.f));
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
      const float wx1 = (float)(x0 + 1) - ix, wx0f = ix - (float)x0, wy1 = (float)(y0 + 1) - iy, wy0f = iy - (float)y0;
      const bool x1ok = x0 + 1 < W, y1ok = y0 + 1 < H;
      scatter_tap(win, gs, HW, W, wx0, wy0, x0, y0, g, wx1 * wy1);
      if (x1ok) scatter_tap(win, gs, HW, W, wx0, wy0, x0 + 1, y0, g, wx0f * wy1);
      if (y1ok) scatter_tap(win, gs, HW, W, wx0, wy0, x0, y0 + 1, g, wx1 * wy0f);
      if (x1ok && y1ok) scatter_tap(win, gs, HW, W, wx0, wy0, x0 + 1, y0 + 1, g, wx0f * wy0f);
    }
  }
  __syncthreads();
  for (int k = tid; k < 3 * kWinH * kWinW; k += kSTW * kSTH) {
    const float v = win[k];
    if (v == 0.f) continue;
    const int c = k / (kWinH * kWinW), r = k - c * (kWinH * kWinW);
    const int ly = r / kWinW, lx = r - ly * kWinW;
    const int sy = wy0 + ly, sx = wx0 + lx;
    if (sx >= 0 && sx < W && sy >= 0 && sy < H) atomicAdd(gs + (size_t)c * HW + (size_t)sy * W + sx, v);
  }
}

}  // namespace vsl

using namespace vsl;

extern "C" {

int vsl_source_grad_upstream(const VslDesc* d, const float* upstream, const uint8_t* const winner[VSL_MAX_SCALES],
                             float* up_identity, float* const up_warped[VSL_MAX_SCALES], void* stream) {
  if (!d || d->abi_version != VSL_ABI_VERSION) return VSL_ERR_BAD_DESC;
  if (d->batch < 1 || d->height < 2 || d->width < 2 || d->num_scales < 1 || d->num_scales > VSL_MAX_SCALES ||
      d->num_src < 1 || d->num_src > VSL_MAX_SRC)
    return VSL_ERR_BAD_DESC;
  if (!upstream || !winner || !up_warped) return VSL_ERR_NULL_POINTER;
  if ((d->flags & VSL_FLAG_AUTOMASK) && !up_identity) return VSL_ERR_NULL_POINTER;
  UpstreamParams p = {};
  p.up = upstream;
  p.S = d->num_scales; p.F = d->num_src;
  p.avg = ((d->flags & VSL_FLAG_AVG_REPROJECTION) && d->num_src > 1) ? 1 : 0;
  p.n = (size_t)d->batch * d->height * d->width;
  p.wpix = 1.0f / ((float)d->batch * d->height * d->width);
  p.up_identity = (d->flags & VSL_FLAG_AUTOMASK) ? up_identity : nullptr;
  for (int s = 0; s < p.S; ++s) {
    if (!winner[s] || !up_warped[s]) return VSL_ERR_NULL_POINTER;
    p.winner[s] = winner[s];
    p.up_warped[s] = up_warped[s];
  }
  k_source_grad_upstream<<<(unsigned)((p.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  VSL_S_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_grid_sample_backward_source(int batch, int height, int width, const float* grid, const float* grad_pred,
                                    float* grad_source, void* stream) {
  if (batch < 1 || height < 2 || width < 2) return VSL_ERR_BAD_DESC;
  if (!grid || !grad_pred || !grad_source) return VSL_ERR_NULL_POINTER;
  dim3 grd((width + kSTW - 1) / kSTW, (height + kSTH - 1) / kSTH, batch), blk(kSTW, kSTH);
  k_grid_sample_bwd_source<<<grd, blk, 0, (cudaStream_t)stream>>>(height, width, grid, grad_pred, grad_source);
  VSL_S_OK(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
