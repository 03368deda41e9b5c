// Stand-alone layers behind the reference's layers.py call surface (include/vsl.h, second half):
// BackprojectDepth, Project3D, SSIM, compute_reprojection_loss, get_smooth_loss, forward and
// backward.  These serve callers that use the layers one at a time; the training hot path is the
// fused kernel in vsl_fused.cu.  Forward arithmetic follows eager PyTorch-CUDA's rounding
// (vsl_math.cuh); backward arithmetic is plain fp32.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsl.h"
#include "vsl_math.cuh"

namespace vsl {

extern thread_local int g_last_cuda_error;  // defined in vsl_fused.cu
#define VSL_L_OK(expr)                                                \
  do {                                                                \
    cudaError_t e__ = (expr);                                         \
    if (e__ != cudaSuccess) { g_last_cuda_error = (int)e__; return VSL_ERR_CUDA; } \
  } while (0)

constexpr int kNT = 256;

__device__ __forceinline__ float warp_sum_l(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum_l(float v, float* scratch) {
  v = warp_sum_l(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kNT / 32; ++i) r += scratch[i];
  return r;
}

// ---- BackprojectDepth (layers.py:234-239) ---------------------------------------------------------
__global__ void __launch_bounds__(kNT) k_backproject_fwd(int B, int H, int W, int arith, const float* __restrict__ depth,
                                                         const float* __restrict__ invK, float* __restrict__ cam) {
  int HW = H * W;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)B * HW) return;
  int b = (int)(gid / HW), i = (int)(gid - (size_t)b * HW);
  float fu = (float)(i % W), fv = (float)(i / W);
  const float* k = invK + b * 16;
  float z = depth[gid];
  float* o = cam + (size_t)b * 4 * HW + i;
  o[0] = mul_rn(z, dot3(k[0], fu, k[1], fv, k[2], 1.0f, arith));
  o[HW] = mul_rn(z, dot3(k[4], fu, k[5], fv, k[6], 1.0f, arith));
  o[2 * HW] = mul_rn(z, dot3(k[8], fu, k[9], fv, k[10], 1.0f, arith));
  o[3 * HW] = 1.0f;
}
__global__ void __launch_bounds__(kNT) k_backproject_bwd(int B, int H, int W, const float* __restrict__ gcam,
                                                         const float* __restrict__ invK, float* __restrict__ gdepth) {
  int HW = H * W;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)B * HW) return;
  int b = (int)(gid / HW), i = (int)(gid - (size_t)b * HW);
  float fu = (float)(i % W), fv = (float)(i / W);
  const float* k = invK + b * 16;
  const float* g = gcam + (size_t)b * 4 * HW + i;
  gdepth[gid] = g[0] * (k[0] * fu + k[1] * fv + k[2]) + g[HW] * (k[4] * fu + k[5] * fv + k[6]) +
                g[2 * HW] * (k[8] * fu + k[9] * fv + k[10]);
}

// ---- Project3D (layers.py:253-264) ------------------------------------------------------------------
__global__ void __launch_bounds__(kNT) k_project_fwd(int B, int H, int W, float eps, int arith,
                                                     const float* __restrict__ pts, const float* __restrict__ P,
                                                     float* __restrict__ pix) {
  int HW = H * W;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)B * HW) return;
  int b = (int)(gid / HW), i = (int)(gid - (size_t)b * HW);
  const float* q = pts + (size_t)b * 4 * HW + i;
  const float* p = P + b * 12;
  float X = q[0], Y = q[HW], Z = q[2 * HW], Wc = q[3 * HW];
  float c0 = dot4(p[0], X, p[1], Y, p[2], Z, p[3], Wc, arith);
  float c1 = dot4(p[4], X, p[5], Y, p[6], Z, p[7], Wc, arith);
  float c2 = dot4(p[8], X, p[9], Y, p[10], Z, p[11], Wc, arith);
  float zeta = add_rn(c2, eps);
  float px = div_rn(c0, zeta), py = div_rn(c1, zeta);
  float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float nx = (arith & kTrueDiv) ? div_rn(px, wm1) : mul_rn(px, 1.0f / wm1);
  float ny = (arith & kTrueDiv) ? div_rn(py, hm1) : mul_rn(py, 1.0f / hm1);
  pix[gid * 2] = mul_rn(sub_rn(nx, 0.5f), 2.0f);
  pix[gid * 2 + 1] = mul_rn(sub_rn(ny, 0.5f), 2.0f);
}
// grad_points per pixel + per-block partial sums of grad_P (12 floats), reduced by k_project_bwd_reduce
__global__ void __launch_bounds__(kNT) k_project_bwd(int B, int H, int W, float eps, const float* __restrict__ pts,
                                                     const float* __restrict__ P, const float* __restrict__ gpix,
                                                     float* __restrict__ gpts, float* __restrict__ part) {
  __shared__ float scratch[kNT / 32];
  int HW = H * W, b = blockIdx.y;
  int i = blockIdx.x * kNT + threadIdx.x;
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  if (i < HW) {
    const float* q = pts + (size_t)b * 4 * HW + i;
    const float* p = P + b * 12;
    float X[4] = {q[0], q[HW], q[2 * HW], q[3 * HW]};
    float c0 = p[0] * X[0] + p[1] * X[1] + p[2] * X[2] + p[3] * X[3];
    float c1 = p[4] * X[0] + p[5] * X[1] + p[6] * X[2] + p[7] * X[3];
    float c2 = p[8] * X[0] + p[9] * X[1] + p[10] * X[2] + p[11] * X[3];
    float iz = 1.0f / (c2 + eps);
    size_t gid = (size_t)b * HW + i;
    float gpx = gpix[gid * 2] * 2.0f / (float)(W - 1), gpy = gpix[gid * 2 + 1] * 2.0f / (float)(H - 1);
    float g0 = gpx * iz, g1 = gpy * iz, g2 = -(g0 * c0 + g1 * c1) * iz;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[k] = g0 * X[k]; acc[4 + k] = g1 * X[k]; acc[8 + k] = g2 * X[k];
      gpts[(size_t)b * 4 * HW + (size_t)k * HW + i] = g0 * p[k] + g1 * p[4 + k] + g2 * p[8 + k];
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    float r = block_sum_l(acc[k], scratch);
    if (threadIdx.x == 0) part[((size_t)b * gridDim.x + blockIdx.x) * 12 + k] = r;
  }
}
__global__ void k_project_bwd_reduce(int nblk, const float* __restrict__ part, float* __restrict__ gP) {
  int b = blockIdx.x, k = threadIdx.x;
  if (k >= 12) return;
  double acc = 0.0;
  for (int j = 0; j < nblk; ++j) acc += (double)part[((size_t)b * nblk + j) * 12 + k];
  gP[b * 12 + k] = (float)acc;
}

// ---- SSIM (layers.py:318-332) and the reprojection loss (trainer.py:543-555) --------------------------
struct Win {  // reflect-padded 3x3 neighbourhood offsets of a window centre
  int o[9];
};
__device__ __forceinline__ Win window_offsets(int y, int x, int H, int W) {
  Win w;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) w.o[(dy + 1) * 3 + dx + 1] = reflect1(y + dy, H) * W + reflect1(x + dx, W);
  return w;
}
__device__ __forceinline__ SsimOut ssim_at(const float* __restrict__ x, const float* __restrict__ y, const Win& w) {
  float sx = 0.f, sxx = 0.f, sxy = 0.f, sy = 0.f, syy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float xv = x[w.o[k]], yv = y[w.o[k]];
    sx = add_rn(sx, xv); sxx = add_rn(sxx, mul_rn(xv, xv)); sxy = add_rn(sxy, mul_rn(xv, yv));
    sy = add_rn(sy, yv); syy = add_rn(syy, mul_rn(yv, yv));
  }
  float mu_y = div9(sy);
  float sig_y = sub_rn(div9(syy), mul_rn(mu_y, mu_y));
  return ssim_from_sums(sx, sxx, sxy, mu_y, sig_y);
}

__global__ void __launch_bounds__(kNT) k_ssim_fwd(int planes, int H, int W, const float* __restrict__ x,
                                                  const float* __restrict__ y, float* __restrict__ out) {
  int HW = H * W;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)planes * HW) return;
  size_t pl = gid / HW;
  int i = (int)(gid - pl * HW);
  Win w = window_offsets(i / W, i % W, H, W);
  out[gid] = ssim_at(x + pl * HW, y + pl * HW, w).val;
}

// Backward of SSIM in two passes.  Pass 1, per window p: the four adjoint coefficients
//   Ax = k dr/dmu_x, Ay = k dr/dmu_y, Bq = k dr/dE[x^2] (= dr/dE[y^2]), Cq = k dr/dE[xy],  k = -g_p/18 * live
// (SSIM is symmetric in its arguments, so x and y share Bq and Cq).  Pass 2, per pixel q: gather the <= 9
// windows that contain q, reflected border rows/columns counting twice:
//   dL/dx_q = sum_p cnt (Ax_p + 2 x_q Bq_p + y_q Cq_p),   dL/dy_q = sum_p cnt (Ay_p + 2 y_q Bq_p + x_q Cq_p).
// coef: [4][planes*HW] workspace.  `gstride`: 1 = one upstream value per plane element (SSIM layer),
// 0-like broadcast over the 3 channels is expressed by passing gdiv = 3 (reprojection loss, go is [B,1,H,W]).
__global__ void __launch_bounds__(kNT) k_ssim_coef(int planes, int H, int W, int gdiv, float gscale,
                                                   const float* __restrict__ x, const float* __restrict__ y,
                                                   const float* __restrict__ go, float* __restrict__ coef) {
  int HW = H * W;
  size_t n = (size_t)planes * HW;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= n) return;
  size_t pl = gid / HW;
  int i = (int)(gid - pl * HW);
  float g = go[(pl / gdiv) * HW + i] * gscale;
  float ax = 0.f, ay = 0.f, bq = 0.f, cq = 0.f;
  if (g != 0.f) {
    const float* xp = x + pl * HW;
    const float* yp = y + pl * HW;
    Win w = window_offsets(i / W, i % W, H, W);
    float sy = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) sy = add_rn(sy, yp[w.o[k]]);
    float mu_y = div9(sy);
    SsimOut so = ssim_at(xp, yp, w);
    if (so.live) {
      float dmu, dexx, dexy;
      ssim_r_grads(so, mu_y, dmu, dexx, dexy);
      // d r / d mu_y: the same expression with the roles of x and y exchanged
      float inv_d = fast_rcp(so.d1 * so.d2);
      float dmuy = inv_d * (2.0f * so.mu_x * (so.n2 - so.n1) - so.r * 2.0f * mu_y * (so.d2 - so.d1));
      float k = g * (-0.5f / 9.0f);
      ax = k * dmu; ay = k * dmuy; bq = k * dexx; cq = k * dexy;
    }
  }
  coef[gid] = ax; coef[n + gid] = ay; coef[2 * n + gid] = bq; coef[3 * n + gid] = cq;
}

__device__ __forceinline__ void ssim_gather(const float* __restrict__ coef, size_t n, size_t plane_off, int qy, int qx,
                                            int H, int W, float& sa_x, float& sa_y, float& sb, float& sc) {
  sa_x = sa_y = sb = sc = 0.f;
  for (int dy = -1; dy <= 1; ++dy) {
    int py = qy + dy;
    if (py < 0 || py >= H) continue;
    float cy = ((dy == -1 && qy == 1) || (dy == 1 && qy == H - 2)) ? 2.f : 1.f;
    for (int dx = -1; dx <= 1; ++dx) {
      int px = qx + dx;
      if (px < 0 || px >= W) continue;
      float cnt = cy * (((dx == -1 && qx == 1) || (dx == 1 && qx == W - 2)) ? 2.f : 1.f);
      size_t o = plane_off + (size_t)py * W + px;
      sa_x += cnt * coef[o]; sa_y += cnt * coef[n + o]; sb += cnt * coef[2 * n + o]; sc += cnt * coef[3 * n + o];
    }
  }
}

__global__ void __launch_bounds__(kNT) k_ssim_bwd(int planes, int H, int W, const float* __restrict__ x,
                                                  const float* __restrict__ y, const float* __restrict__ coef,
                                                  float* __restrict__ gx, float* __restrict__ gy) {
  int HW = H * W;
  size_t n = (size_t)planes * HW;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= n) return;
  size_t pl = gid / HW;
  int i = (int)(gid - pl * HW);
  float sax, say, sb, sc;
  ssim_gather(coef, n, pl * HW, i / W, i % W, H, W, sax, say, sb, sc);
  float xq = x[gid], yq = y[gid];
  if (gx) gx[gid] = sax + 2.f * xq * sb + yq * sc;
  if (gy) gy[gid] = say + 2.f * yq * sb + xq * sc;
}

__global__ void __launch_bounds__(kNT) k_reproj_fwd(int B, int H, int W, int no_ssim, int arith,
                                                    const float* __restrict__ pred, const float* __restrict__ tgt,
                                                    float* __restrict__ out) {
  int HW = H * W;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)B * HW) return;
  int b = (int)(gid / HW), i = (int)(gid - (size_t)b * HW);
  const float* x = pred + (size_t)b * 3 * HW;
  const float* y = tgt + (size_t)b * 3 * HW;
  float l1[3], ss[3];
  Win w = window_offsets(i / W, i % W, H, W);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    l1[c] = fabsf(sub_rn(y[c * HW + i], x[c * HW + i]));
    ss[c] = no_ssim ? 0.f : ssim_at(x + c * HW, y + c * HW, w).val;
  }
  float ml = mean3(l1[0], l1[1], l1[2], arith);
  out[gid] = no_ssim ? ml : add_rn(mul_rn(0.85f, mean3(ss[0], ss[1], ss[2], arith)), mul_rn(0.15f, ml));
}

__global__ void __launch_bounds__(kNT) k_reproj_bwd(int B, int H, int W, int no_ssim, const float* __restrict__ pred,
                                                    const float* __restrict__ tgt, const float* __restrict__ go,
                                                    const float* __restrict__ coef, float* __restrict__ gpred,
                                                    float* __restrict__ gtgt) {
  int HW = H * W;
  size_t n = (size_t)B * 3 * HW;
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= n) return;
  size_t pl = gid / HW;  // b * 3 + c
  int i = (int)(gid - pl * HW);
  float xq = pred[gid], yq = tgt[gid];
  float d = xq - yq;
  float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  float l1g = go[(pl / 3) * HW + i] * (no_ssim ? (1.0f / 3.0f) : (0.15f / 3.0f)) * sgn;
  float sax = 0.f, say = 0.f, sb = 0.f, sc = 0.f;
  if (!no_ssim) ssim_gather(coef, n, pl * HW, i / W, i % W, H, W, sax, say, sb, sc);
  if (gpred) gpred[gid] = l1g + sax + 2.f * xq * sb + yq * sc;
  if (gtgt) gtgt[gid] = -l1g + say + 2.f * yq * sb + xq * sc;
}

// ---- get_smooth_loss (layers.py:286-299) --------------------------------------------------------------
__device__ __forceinline__ float edge_w(const float* img, int n, int ia, int ib) {
  float g = fabsf(img[ia] - img[ib]) + fabsf(img[n + ia] - img[n + ib]) + fabsf(img[2 * n + ia] - img[2 * n + ib]);
  return expf(-g * (1.0f / 3.0f));
}
__global__ void __launch_bounds__(kNT) k_smooth_fwd(int H, int W, const float* __restrict__ disp,
                                                    const float* __restrict__ img, float* __restrict__ part) {
  __shared__ float scratch[kNT / 32];
  int n = H * W, b = blockIdx.y;
  int i = blockIdx.x * kNT + threadIdx.x;
  const float* d = disp + (size_t)b * n;
  const float* im = img + (size_t)b * 3 * n;
  float sx = 0.f, sy = 0.f;
  if (i < n) {
    int v = i / W, u = i - v * W;
    if (u + 1 < W) sx = fabsf(d[i] - d[i + 1]) * edge_w(im, n, i, i + 1);
    if (v + 1 < H) sy = fabsf(d[i] - d[i + W]) * edge_w(im, n, i, i + W);
  }
  sx = block_sum_l(sx, scratch);
  sy = block_sum_l(sy, scratch);
  if (threadIdx.x == 0) {
    part[((size_t)b * gridDim.x + blockIdx.x) * 2] = sx;
    part[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = sy;
  }
}
__global__ void k_smooth_fwd_reduce(int nblk, int B, int H, int W, const float* __restrict__ part,
                                    float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sx = 0.0, sy = 0.0;
  for (size_t j = 0; j < (size_t)nblk * B; ++j) { sx += (double)part[j * 2]; sy += (double)part[j * 2 + 1]; }
  *loss = (float)(sx / ((double)B * H * (W - 1)) + sy / ((double)B * (H - 1) * W));
}
__global__ void __launch_bounds__(kNT) k_smooth_bwd(int B, int H, int W, const float* __restrict__ disp,
                                                    const float* __restrict__ img, const float* __restrict__ gl,
                                                    float* __restrict__ gdisp) {
  int n = H * W, b = blockIdx.y;
  int i = blockIdx.x * kNT + threadIdx.x;
  if (i >= n) return;
  const float* d = disp + (size_t)b * n;
  const float* im = img + (size_t)b * 3 * n;
  float cx = *gl / ((float)B * H * (W - 1)), cy = *gl / ((float)B * (H - 1) * W);
  int v = i / W, u = i - v * W;
  float di = d[i], g = 0.f;
  auto sg = [](float t) { return t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f); };
  if (u + 1 < W) g += sg(di - d[i + 1]) * edge_w(im, n, i, i + 1) * cx;
  if (u > 0) g -= sg(d[i - 1] - di) * edge_w(im, n, i - 1, i) * cx;
  if (v + 1 < H) g += sg(di - d[i + W]) * edge_w(im, n, i, i + W) * cy;
  if (v > 0) g -= sg(d[i - W] - di) * edge_w(im, n, i - W, i) * cy;
  gdisp[(size_t)b * n + i] = g;
}

__global__ void __launch_bounds__(kNT) k_probe_bmm(int B, int K, int N, int arith, const float* __restrict__ A,
                                                   const float* __restrict__ X, float* __restrict__ out) {
  size_t gid = (size_t)blockIdx.x * kNT + threadIdx.x;
  if (gid >= (size_t)B * N) return;
  int b = (int)(gid / N), n = (int)(gid - (size_t)b * N);
  const float* a = A + (size_t)b * 3 * K;
  const float* x = X + (size_t)b * K * N + n;
  for (int i = 0; i < 3; ++i) {
    float r = (K == 3) ? dot3(a[i * 3], x[0], a[i * 3 + 1], x[N], a[i * 3 + 2], x[2 * (size_t)N], arith)
                       : dot4(a[i * 4], x[0], a[i * 4 + 1], x[N], a[i * 4 + 2], x[2 * (size_t)N], a[i * 4 + 3],
                              x[3 * (size_t)N], arith);
    out[(size_t)b * 3 * N + (size_t)i * N + n] = r;
  }
}

// ---- transformation_from_parameters (layers.py:97-172) ---------------------------------------------------
__global__ void __launch_bounds__(kNT) k_pose_fwd(int B, int invert, int arith, const float* __restrict__ aa,
                                                  const float* __restrict__ tr, float* __restrict__ T) {
  int b = blockIdx.x * kNT + threadIdx.x;
  if (b >= B) return;
  float v[3] = {aa[3 * b], aa[3 * b + 1], aa[3 * b + 2]}, t[3] = {tr[3 * b], tr[3 * b + 1], tr[3 * b + 2]}, M[16];
  pose_matrix(v, t, invert != 0, arith, M);
#pragma unroll
  for (int k = 0; k < 16; ++k) T[16 * b + k] = M[k];
}
// Rodrigues backward: R = ca I + sa [a]x + (1 - ca) a a^T, a = v / (|v| + 1e-7)
__device__ __forceinline__ void pose_backward(const float v[3], const float t[3], bool invert, const float* __restrict__ g,
                                              float gaa[3], float gtr[3]) {
  const float th = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), den = th + 1e-7f;
  const float a[3] = {v[0] / den, v[1] / den, v[2] / den};
  const float ca = cosf(th), sa = sinf(th), C = 1.0f - ca;
  float R[9] = {a[0] * a[0] * C + ca, a[0] * a[1] * C - a[2] * sa, a[2] * a[0] * C + a[1] * sa,
                a[0] * a[1] * C + a[2] * sa, a[1] * a[1] * C + ca, a[1] * a[2] * C - a[0] * sa,
                a[2] * a[0] * C - a[1] * sa, a[1] * a[2] * C + a[0] * sa, a[2] * a[2] * C + ca};
  float gR[9], gt[3];
  if (!invert) {  // M = [R | t]
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) gR[i * 3 + j] = g[i * 4 + j];
      gt[i] = g[i * 4 + 3];
    }
  } else {        // M = [R^T | -R^T t]
    const float gc[3] = {g[3], g[7], g[11]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gt[k] = -(R[k * 3] * gc[0] + R[k * 3 + 1] * gc[1] + R[k * 3 + 2] * gc[2]);
#pragma unroll
      for (int i = 0; i < 3; ++i) gR[k * 3 + i] = g[i * 4 + k] - t[k] * gc[i];
    }
  }
  // R -> (ca, sa, a)
  float gC = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) gC += gR[i * 3 + j] * a[i] * a[j];
  const float gca = gR[0] + gR[4] + gR[8] - gC;
  const float gsa = a[0] * (gR[7] - gR[5]) + a[1] * (gR[2] - gR[6]) + a[2] * (gR[3] - gR[1]);
  float ga[3] = {sa * (gR[7] - gR[5]), sa * (gR[2] - gR[6]), sa * (gR[3] - gR[1])};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) ga[i] += C * (gR[i * 3 + j] + gR[j * 3 + i]) * a[j];
  float gth = gsa * ca - gca * sa;
  // a = v / (th + eps)
  gth -= (ga[0] * v[0] + ga[1] * v[1] + ga[2] * v[2]) / (den * den);
  const float k = th > 0.f ? gth / th : 0.f;  // d|v|/dv = v/|v| (torch gives 0 at the origin)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    gaa[i] = ga[i] / den + k * v[i];
    gtr[i] = gt[i];
  }
}
__global__ void __launch_bounds__(kNT) k_pose_bwd(int B, int invert, const float* __restrict__ aa,
                                                  const float* __restrict__ tr, const float* __restrict__ gT,
                                                  float* __restrict__ gaa, float* __restrict__ gtr) {
  int b = blockIdx.x * kNT + threadIdx.x;
  if (b >= B) return;
  const float v[3] = {aa[3 * b], aa[3 * b + 1], aa[3 * b + 2]}, t[3] = {tr[3 * b], tr[3 * b + 1], tr[3 * b + 2]};
  float ga[3], gt[3];
  pose_backward(v, t, invert != 0, gT + 16 * b, ga, gt);
#pragma unroll
  for (int i = 0; i < 3; ++i) { gaa[3 * b + i] = ga[i]; gtr[3 * b + i] = gt[i]; }
}

// ---- posecnn pose tail (trainer.py:516-525) ------------------------------------------------------------------
// T_{s,f} = transformation_from_parameters(axisangle_f, translation_f * mean_inv_depth_s, f < 0) with
// mean_inv_depth_s = mean over the pixels of 1 / depth_s (depth_s from the up-sampled disp_s): one reduction
// launch for every scale and image, one launch for every pose.  The per-pixel values are the reference's bits
// (up-sample, disp_to_depth, two reciprocals); the MEAN is accumulated in fp64 in a fixed order, whereas the
// reference takes two fp32 torch means in a row, so T agrees to ~1e-7 relative, not bit for bit (opt-in, see
// trainer.vsl_posecnn_tail).
constexpr int kPcChunk = 4096;  // pixels per block of the reduction
struct PoseCnnParams {
  const float* disp[VSL_MAX_SCALES];
  int hs[VSL_MAX_SCALES], ws[VSL_MAX_SCALES], identity[VSL_MAX_SCALES];
  float scale_h[VSL_MAX_SCALES], scale_w[VSL_MAX_SCALES];
  const float* aa[VSL_MAX_SRC];
  const float* tr[VSL_MAX_SRC];
  int invert[VSL_MAX_SRC];
  int B, H, W, S, F, arith, nchunk;
  GeoConst g;
  double* partial;   // [S][B][nchunk]
  float* mean_inv;   // [S][B]
  float* T;          // [S][F][B][16]
};
__global__ void __launch_bounds__(kNT) k_posecnn_partial(const PoseCnnParams p) {
  __shared__ double scratch[kNT / 32];
  const int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x, HW = p.H * p.W;
  const float* d = p.disp[s] + (size_t)b * p.hs[s] * p.ws[s];
  double acc = 0.0;
  for (int i = chunk * kPcChunk + threadIdx.x; i < min(HW, (chunk + 1) * kPcChunk); i += kNT) {
    const int v = i / p.W, u = i - v * p.W;
    const float D = upsample_disp(d, p.hs[s], p.ws[s], p.scale_h[s], p.scale_w[s], p.identity[s] != 0, v, u, p.arith);
    acc += (double)rcp_rn(disp_to_z(D, p.g));  // inv_depth = 1 / depth, depth = 1 / scaled_disp (layers.py:90-93)
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < kNT / 32; ++i) r += scratch[i];
    p.partial[((size_t)s * p.B + b) * p.nchunk + chunk] = r;
  }
}
__global__ void __launch_bounds__(kNT) k_posecnn_pose(const PoseCnnParams p) {
  // grid-stride over (s, f, b); every thread re-adds its (s, b) partials in chunk order (a few dozen values)
  for (int idx = blockIdx.x * kNT + threadIdx.x; idx < p.S * p.F * p.B; idx += gridDim.x * kNT) {
    const int b = idx % p.B, f = (idx / p.B) % p.F, s = idx / (p.B * p.F);
    double sum = 0.0;
    for (int c = 0; c < p.nchunk; ++c) sum += p.partial[((size_t)s * p.B + b) * p.nchunk + c];
    const float mean = (float)(sum / (double)(p.H * p.W));
    if (f == 0) p.mean_inv[s * p.B + b] = mean;
    const float v[3] = {p.aa[f][3 * b], p.aa[f][3 * b + 1], p.aa[f][3 * b + 2]};
    const float t[3] = {mul_rn(p.tr[f][3 * b], mean), mul_rn(p.tr[f][3 * b + 1], mean), mul_rn(p.tr[f][3 * b + 2], mean)};
    float M[16];
    pose_matrix(v, t, p.invert[f] != 0, p.arith, M);
#pragma unroll
    for (int k = 0; k < 16; ++k) p.T[(size_t)idx * 16 + k] = M[k];
  }
}
struct PoseCnnBwdParams {
  const float* aa[VSL_MAX_SRC];
  const float* tr[VSL_MAX_SRC];
  int invert[VSL_MAX_SRC];
  float* gaa[VSL_MAX_SRC];
  float* gtr[VSL_MAX_SRC];
  const float* mean_inv;   // [S][B]
  const float* gT;         // [S][F][B][16]
  float* gdisp_const;      // [S][B]: d L / d disp_s[b, any pixel] contributed through the mean
  float per_pixel[VSL_MAX_SCALES];  // disp_range / (hs * ws): d mean_inv_depth_s / d disp_s[j] (the up-sample preserves the mean)
  int B, S, F;
};
__global__ void __launch_bounds__(kNT) k_posecnn_bwd(const PoseCnnBwdParams p) {
  const int b = blockIdx.x * kNT + threadIdx.x;
  if (b >= p.B) return;
  float gmean[VSL_MAX_SCALES] = {0.f, 0.f, 0.f, 0.f};
  for (int f = 0; f < p.F; ++f) {
    const float v[3] = {p.aa[f][3 * b], p.aa[f][3 * b + 1], p.aa[f][3 * b + 2]};
    const float tr[3] = {p.tr[f][3 * b], p.tr[f][3 * b + 1], p.tr[f][3 * b + 2]};
    float ga[3] = {0.f, 0.f, 0.f}, gt[3] = {0.f, 0.f, 0.f};
    for (int s = 0; s < p.S; ++s) {
      const float m = p.mean_inv[s * p.B + b];
      const float t[3] = {tr[0] * m, tr[1] * m, tr[2] * m};
      float ga_s[3], gt_s[3];
      pose_backward(v, t, p.invert[f] != 0, p.gT + ((size_t)(s * p.F + f) * p.B + b) * 16, ga_s, gt_s);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ga[i] += ga_s[i];
        gt[i] += gt_s[i] * m;
        gmean[s] += gt_s[i] * tr[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { p.gaa[f][3 * b + i] = ga[i]; p.gtr[f][3 * b + i] = gt[i]; }
  }
  for (int s = 0; s < p.S; ++s) p.gdisp_const[s * p.B + b] = gmean[s] * p.per_pixel[s];
}

static unsigned blocks_for(size_t n) { return (unsigned)((n + kNT - 1) / kNT); }

}  // namespace vsl

using namespace vsl;

extern "C" {

int vsl_pose_forward(int B, int invert, int arith, const float* axisangle, const float* translation, float* T,
                     void* stream) {
  if (B < 1) return VSL_ERR_BAD_DESC;
  if (!axisangle || !translation || !T) return VSL_ERR_NULL_POINTER;
  k_pose_fwd<<<blocks_for((size_t)B), kNT, 0, (cudaStream_t)stream>>>(B, invert, arith, axisangle, translation, T);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_pose_backward(int B, int invert, const float* axisangle, const float* translation, const float* grad_T,
                      float* grad_axisangle, float* grad_translation, void* stream) {
  if (B < 1) return VSL_ERR_BAD_DESC;
  if (!axisangle || !translation || !grad_T || !grad_axisangle || !grad_translation) return VSL_ERR_NULL_POINTER;
  k_pose_bwd<<<blocks_for((size_t)B), kNT, 0, (cudaStream_t)stream>>>(B, invert, axisangle, translation, grad_T,
                                                                      grad_axisangle, grad_translation);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_probe_bmm(int B, int K, int N, int arith, const float* A, const float* X, float* out, void* stream) {
  if (B < 1 || N < 1 || (K != 3 && K != 4)) return VSL_ERR_BAD_DESC;
  if (!A || !X || !out) return VSL_ERR_NULL_POINTER;
  k_probe_bmm<<<blocks_for((size_t)B * N), kNT, 0, (cudaStream_t)stream>>>(B, K, N, arith, A, X, out);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_backproject_forward(int B, int H, int W, int arith, const float* depth, const float* inv_K, float* cam,
                            void* stream) {
  if (B < 1 || H < 1 || W < 1) return VSL_ERR_BAD_DESC;
  if (!depth || !inv_K || !cam) return VSL_ERR_NULL_POINTER;
  k_backproject_fwd<<<blocks_for((size_t)B * H * W), kNT, 0, (cudaStream_t)stream>>>(B, H, W, arith, depth, inv_K, cam);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_backproject_backward(int B, int H, int W, const float* gcam, const float* inv_K, float* gdepth, void* stream) {
  if (B < 1 || H < 1 || W < 1) return VSL_ERR_BAD_DESC;
  if (!gcam || !inv_K || !gdepth) return VSL_ERR_NULL_POINTER;
  k_backproject_bwd<<<blocks_for((size_t)B * H * W), kNT, 0, (cudaStream_t)stream>>>(B, H, W, gcam, inv_K, gdepth);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_project_forward(int B, int H, int W, float eps, int arith, const float* points, const float* P, float* pix,
                        void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!points || !P || !pix) return VSL_ERR_NULL_POINTER;
  k_project_fwd<<<blocks_for((size_t)B * H * W), kNT, 0, (cudaStream_t)stream>>>(B, H, W, eps, arith, points, P, pix);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
size_t vsl_project_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  return (size_t)B * blocks_for((size_t)H * W) * 12 * sizeof(float);
}
int vsl_project_backward(int B, int H, int W, float eps, const float* points, const float* P, const float* gpix,
                         float* gpoints, float* gP, void* ws, size_t ws_bytes, void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!points || !P || !gpix || !gpoints || !gP || !ws) return VSL_ERR_NULL_POINTER;
  if (ws_bytes < vsl_project_workspace_bytes(B, H, W)) return VSL_ERR_WORKSPACE;
  unsigned nblk = blocks_for((size_t)H * W);
  k_project_bwd<<<dim3(nblk, B), kNT, 0, (cudaStream_t)stream>>>(B, H, W, eps, points, P, gpix, gpoints, (float*)ws);
  VSL_L_OK(cudaGetLastError());
  k_project_bwd_reduce<<<B, 32, 0, (cudaStream_t)stream>>>((int)nblk, (const float*)ws, gP);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_ssim_forward(int B, int C, int H, int W, const float* x, const float* y, float* out, void* stream) {
  if (B < 1 || C < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!x || !y || !out) return VSL_ERR_NULL_POINTER;
  k_ssim_fwd<<<blocks_for((size_t)B * C * H * W), kNT, 0, (cudaStream_t)stream>>>(B * C, H, W, x, y, out);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
size_t vsl_ssim_workspace_bytes(int B, int C, int H, int W) {
  if (B < 1 || C < 1 || H < 2 || W < 2) return 0;
  return (size_t)4 * B * C * H * W * sizeof(float);
}
int vsl_ssim_backward(int B, int C, int H, int W, const float* x, const float* y, const float* go, float* gx, float* gy,
                      void* ws, size_t ws_bytes, void* stream) {
  if (B < 1 || C < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!x || !y || !go || !ws) return VSL_ERR_NULL_POINTER;
  if (ws_bytes < vsl_ssim_workspace_bytes(B, C, H, W)) return VSL_ERR_WORKSPACE;
  unsigned nb = blocks_for((size_t)B * C * H * W);
  k_ssim_coef<<<nb, kNT, 0, (cudaStream_t)stream>>>(B * C, H, W, 1, 1.0f, x, y, go, (float*)ws);
  VSL_L_OK(cudaGetLastError());
  k_ssim_bwd<<<nb, kNT, 0, (cudaStream_t)stream>>>(B * C, H, W, x, y, (const float*)ws, gx, gy);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_reprojection_loss_forward(int B, int H, int W, int no_ssim, int arith, const float* pred, const float* target,
                                  float* out, void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!pred || !target || !out) return VSL_ERR_NULL_POINTER;
  k_reproj_fwd<<<blocks_for((size_t)B * H * W), kNT, 0, (cudaStream_t)stream>>>(B, H, W, no_ssim, arith, pred, target, out);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_reprojection_loss_backward(int B, int H, int W, int no_ssim, const float* pred, const float* target,
                                   const float* go, float* gpred, float* gtarget, void* ws, size_t ws_bytes,
                                   void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!pred || !target || !go || (!no_ssim && !ws)) return VSL_ERR_NULL_POINTER;
  if (!no_ssim && ws_bytes < vsl_ssim_workspace_bytes(B, 3, H, W)) return VSL_ERR_WORKSPACE;
  unsigned nb = blocks_for((size_t)B * 3 * H * W);
  if (!no_ssim) {  // 0.85 * mean_c ssim: the upstream value of pixel p reaches each channel's window scaled by 0.85/3
    k_ssim_coef<<<nb, kNT, 0, (cudaStream_t)stream>>>(B * 3, H, W, 3, 0.85f / 3.0f, pred, target, go, (float*)ws);
    VSL_L_OK(cudaGetLastError());
  }
  k_reproj_bwd<<<nb, kNT, 0, (cudaStream_t)stream>>>(B, H, W, no_ssim, pred, target, go, (const float*)ws, gpred, gtarget);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
size_t vsl_smooth_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  return (size_t)B * blocks_for((size_t)H * W) * 2 * sizeof(float);
}
int vsl_smooth_loss_forward(int B, int H, int W, const float* disp, const float* img, float* loss, void* ws,
                            size_t ws_bytes, void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!disp || !img || !loss || !ws) return VSL_ERR_NULL_POINTER;
  if (ws_bytes < vsl_smooth_workspace_bytes(B, H, W)) return VSL_ERR_WORKSPACE;
  unsigned nblk = blocks_for((size_t)H * W);
  k_smooth_fwd<<<dim3(nblk, B), kNT, 0, (cudaStream_t)stream>>>(H, W, disp, img, (float*)ws);
  VSL_L_OK(cudaGetLastError());
  k_smooth_fwd_reduce<<<1, 32, 0, (cudaStream_t)stream>>>((int)nblk, B, H, W, (const float*)ws, loss);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_smooth_loss_backward(int B, int H, int W, const float* disp, const float* img, const float* gl, float* gdisp,
                             void* stream) {
  if (B < 1 || H < 2 || W < 2) return VSL_ERR_BAD_DESC;
  if (!disp || !img || !gl || !gdisp) return VSL_ERR_NULL_POINTER;
  k_smooth_bwd<<<dim3(blocks_for((size_t)H * W), B), kNT, 0, (cudaStream_t)stream>>>(B, H, W, disp, img, gl, gdisp);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}


static bool posecnn_desc_ok(const VslDesc* d, int num_frames) {
  if (!d || d->abi_version != VSL_ABI_VERSION) return false;
  if (d->batch < 1 || d->height < 2 || d->width < 2 || d->num_scales < 1 || d->num_scales > VSL_MAX_SCALES) return false;
  if (num_frames < 1 || num_frames > VSL_MAX_SRC) return false;
  for (int s = 0; s < d->num_scales; ++s) {
    const int e = d->scale_ids[s];
    if (e < 0 || e > 3 || ((d->height >> e) << e) != d->height || ((d->width >> e) << e) != d->width) return false;
  }
  return true;
}
size_t vsl_posecnn_workspace_bytes(const VslDesc* d) {
  if (!posecnn_desc_ok(d, 1)) return 0;
  const size_t nchunk = ((size_t)d->height * d->width + kPcChunk - 1) / kPcChunk;
  return (size_t)d->num_scales * d->batch * nchunk * sizeof(double);
}
int vsl_posecnn_forward(const VslDesc* d, const float* const disp[VSL_MAX_SCALES], int num_frames,
                        const float* const axisangle[VSL_MAX_SRC], const float* const translation[VSL_MAX_SRC],
                        const int32_t* invert, int pose_arith, float* T, float* mean_inv, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!posecnn_desc_ok(d, num_frames)) return VSL_ERR_BAD_DESC;
  if (!disp || !axisangle || !translation || !invert || !T || !mean_inv || !workspace) return VSL_ERR_NULL_POINTER;
  if (workspace_bytes < vsl_posecnn_workspace_bytes(d)) return VSL_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 7u) != 0) return VSL_ERR_MISALIGNED;
  PoseCnnParams p = {};
  p.B = d->batch; p.H = d->height; p.W = d->width; p.S = d->num_scales; p.F = num_frames;
  p.arith = d->arith | pose_arith;
  p.nchunk = (d->height * d->width + kPcChunk - 1) / kPcChunk;
  p.g.min_disp = d->min_disp; p.g.disp_range = d->disp_range; p.g.eps = d->eps; p.g.one = 1.0f;
  p.g.W = d->width; p.g.H = d->height; p.g.arith = d->arith;
  p.partial = (double*)workspace; p.mean_inv = mean_inv; p.T = T;
  for (int s = 0; s < p.S; ++s) {
    if (!disp[s]) return VSL_ERR_NULL_POINTER;
    const int e = d->scale_ids[s];
    p.disp[s] = disp[s]; p.hs[s] = d->height >> e; p.ws[s] = d->width >> e; p.identity[s] = e == 0;
    p.scale_h[s] = (float)p.hs[s] / (float)d->height; p.scale_w[s] = (float)p.ws[s] / (float)d->width;
  }
  for (int f = 0; f < p.F; ++f) {
    if (!axisangle[f] || !translation[f]) return VSL_ERR_NULL_POINTER;
    p.aa[f] = axisangle[f]; p.tr[f] = translation[f]; p.invert[f] = invert[f];
  }
  cudaStream_t st = (cudaStream_t)stream;
  k_posecnn_partial<<<dim3(p.nchunk, p.B, p.S), kNT, 0, st>>>(p);
  VSL_L_OK(cudaGetLastError());
  k_posecnn_pose<<<blocks_for((size_t)p.S * p.F * p.B), kNT, 0, st>>>(p);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}
int vsl_posecnn_backward(const VslDesc* d, int num_frames, const float* const axisangle[VSL_MAX_SRC],
                         const float* const translation[VSL_MAX_SRC], const int32_t* invert, const float* mean_inv,
                         const float* grad_T, float* const grad_axisangle[VSL_MAX_SRC],
                         float* const grad_translation[VSL_MAX_SRC], float* grad_disp_const, void* stream) {
  if (!posecnn_desc_ok(d, num_frames)) return VSL_ERR_BAD_DESC;
  if (!axisangle || !translation || !invert || !mean_inv || !grad_T || !grad_axisangle || !grad_translation || !grad_disp_const)
    return VSL_ERR_NULL_POINTER;
  PoseCnnBwdParams p = {};
  p.B = d->batch; p.S = d->num_scales; p.F = num_frames;
  p.mean_inv = mean_inv; p.gT = grad_T; p.gdisp_const = grad_disp_const;
  for (int s = 0; s < p.S; ++s) {
    const int e = d->scale_ids[s];
    p.per_pixel[s] = d->disp_range / ((float)(d->height >> e) * (float)(d->width >> e));
  }
  for (int f = 0; f < p.F; ++f) {
    if (!axisangle[f] || !translation[f] || !grad_axisangle[f] || !grad_translation[f]) return VSL_ERR_NULL_POINTER;
    p.aa[f] = axisangle[f]; p.tr[f] = translation[f]; p.invert[f] = invert[f];
    p.gaa[f] = grad_axisangle[f]; p.gtr[f] = grad_translation[f];
  }
  k_posecnn_bwd<<<blocks_for((size_t)p.B), kNT, 0, (cudaStream_t)stream>>>(p);
  VSL_L_OK(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
