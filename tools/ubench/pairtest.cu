// Developer check: packed-pair SSIM helpers vs the scalar chain, bit for bit, on random plausible inputs.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../unsupervised_pose_estimation_b200/csrc/vsl_math.cuh"
using namespace vsl;
__device__ unsigned h32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ float u01(unsigned x) { return (h32(x) >> 8) * (1.0f / 16777216.0f); }
__global__ void k(unsigned long long* bad, unsigned n, float onef) {
  const F2 one = splat(onef);
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // 9 taps of two x images and one y image
  float x0[9], x1[9], y[9];
  for (int t = 0; t < 9; ++t) { x0[t] = u01(i * 31u + t); x1[t] = u01(i * 31u + 9 + t); y[t] = u01(i * 31u + 18 + t); }
  if ((i & 7) == 0) for (int t = 0; t < 9; ++t) x1[t] = 0.f;           // black window
  if ((i & 7) == 1) for (int t = 0; t < 9; ++t) x0[t] = y[t];           // identical images
  float sy = 0, syy = 0, sx0 = 0, sxx0 = 0, sxy0 = 0, sx1 = 0, sxx1 = 0, sxy1 = 0;
  F2 sx = splat(0.f), sxx = splat(0.f), sxy = splat(0.f);
  for (int t = 0; t < 9; ++t) {
    sy = add_rn(sy, y[t]); syy = add_rn(syy, mul_rn(y[t], y[t]));
    sx0 = add_rn(sx0, x0[t]); sxx0 = add_rn(sxx0, mul_rn(x0[t], x0[t])); sxy0 = add_rn(sxy0, mul_rn(x0[t], y[t]));
    sx1 = add_rn(sx1, x1[t]); sxx1 = add_rn(sxx1, mul_rn(x1[t], x1[t])); sxy1 = add_rn(sxy1, mul_rn(x1[t], y[t]));
    F2 xv = f2(x0[t], x1[t]);
    sx = add2(sx, xv); sxx = addp(sxx, xv, xv, one); sxy = addp(sxy, xv, splat(y[t]), one);
  }
  float mu_y = div9(sy), sig_y = sub_rn(div9(syy), mul_rn(mu_y, mu_y));
  unsigned long long nb[6] = {0, 0, 0, 0, 0, 0};
  if (__float_as_uint(sx.x) != __float_as_uint(sx0) || __float_as_uint(sx.y) != __float_as_uint(sx1)) nb[0]++;
  if (__float_as_uint(sxx.x) != __float_as_uint(sxx0) || __float_as_uint(sxx.y) != __float_as_uint(sxx1)) nb[1]++;
  if (__float_as_uint(sxy.x) != __float_as_uint(sxy0) || __float_as_uint(sxy.y) != __float_as_uint(sxy1)) nb[2]++;
  F2 q = div9_2(sxx);
  if (__float_as_uint(q.x) != __float_as_uint(div9(sxx0)) || __float_as_uint(q.y) != __float_as_uint(div9(sxx1))) nb[3]++;
  F2 v = ssim_val2(sx, sxx, sxy, mu_y, sig_y, one);
  float v0 = ssim_from_sums(sx0, sxx0, sxy0, mu_y, sig_y).val, v1 = ssim_from_sums(sx1, sxx1, sxy1, mu_y, sig_y).val;
  if (__float_as_uint(v.x) != __float_as_uint(v0) || __float_as_uint(v.y) != __float_as_uint(v1)) nb[4]++;
  F2 m = mean3_2(v, f2(x0[0], x1[0]), f2(x0[1], x1[1]), 0);
  if (__float_as_uint(m.x) != __float_as_uint(mean3(v0, x0[0], x0[1], 0)) || __float_as_uint(m.y) != __float_as_uint(mean3(v1, x1[0], x1[1], 0))) nb[5]++;
  for (int j = 0; j < 6; ++j) if (nb[j]) atomicAdd(bad + j, nb[j]);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 48); cudaMemset(d, 0, 48);
  unsigned n = 1u << 26;
  k<<<n / 256, 256>>>(d, n, 1.0f);
  unsigned long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("mismatches of %u: sx %llu sxx %llu sxy %llu div9 %llu ssim %llu mean %llu (%s)\n", n, h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
}
