// Developer micro-benchmarks / exhaustive checks run on the B200 before committing to a design:
//  1. do packed fp32x2 ops (FADD2/FMUL2/FFMA2, sm_100) issue at the scalar rate?
//  2. is the 3-instruction division by 9 (Markstein correction with RN(1/9)) correctly rounded for
//     every float, and the branch-free 8-instruction general division equal to __fdiv_rn?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float div9_fast(float a) {
  const float y = 1.0f / 9.0f;  // RN(1/9)
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-9.0f, q, a);
  return __fmaf_rn(r, y, q);
}
__device__ __forceinline__ float div_fast(float n, float d) {
  float y = rcp_approx(d);
  float e = __fmaf_rn(-d, y, 1.0f);
  y = __fmaf_rn(y, e, y);
  float q = __fmul_rn(n, y);
  float r = __fmaf_rn(-d, q, n);
  q = __fmaf_rn(y, r, q);
  r = __fmaf_rn(-d, q, n);
  return __fmaf_rn(y, r, q);
}

__global__ void k_div9_exhaustive(unsigned long long* bad, unsigned long long* tested) {
  unsigned long long nb = 0, nt = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32);
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float a = __uint_as_float((unsigned)i);
    float m = fabsf(a);
    if (!(m >= 1e-30f && m <= 1e30f)) continue;
    ++nt;
    if (__float_as_uint(div9_fast(a)) != __float_as_uint(__fdiv_rn(a, 9.0f))) ++nb;
  }
  atomicAdd(bad, nb); atomicAdd(tested, nt);
}
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// random pairs with exponents in [2^-40, 2^40]
__global__ void k_div_random(unsigned long long n_pairs, unsigned long long* bad, unsigned seed) {
  unsigned long long nb = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned h1 = hash32((unsigned)i * 2u + seed), h2 = hash32((unsigned)i * 2u + 1u + seed * 77u);
    unsigned e1 = 87 + (h1 >> 23) % 80, e2 = 87 + (h2 >> 23) % 80;
    float n = __uint_as_float((h1 & 0x807fffffu) | (e1 << 23));
    float d = __uint_as_float((h2 & 0x807fffffu) | (e2 << 23));
    if (__float_as_uint(div_fast(n, d)) != __float_as_uint(__fdiv_rn(n, d))) ++nb;
  }
  atomicAdd(bad, nb);
}
// all-ones and other special divisor mantissas x random numerators
__global__ void k_div_special(unsigned long long* bad) {
  unsigned long long nb = 0;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  for (unsigned man = 0x7ffff0u; man <= 0x7fffffu; ++man)
    for (unsigned rep = 0; rep < 64; ++rep) {
      unsigned h = hash32(i * 64u + rep);
      float n = __uint_as_float((h & 0x007fffffu) | ((100u + (h >> 26)) << 23));
      float d = __uint_as_float(man | ((110u + (rep & 31u)) << 23));
      if (__float_as_uint(div_fast(n, d)) != __float_as_uint(__fdiv_rn(n, d))) ++nb;
    }
  atomicAdd(bad, nb);
}

template <int MODE>
__global__ void k_tput(float* out, int iters) {
  float2 a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = make_float2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f - k);
  float2 c = make_float2(1.0000001f, 0.9999999f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) { a[k].x = __fadd_rn(a[k].x, c.x); a[k].y = __fadd_rn(a[k].y, c.y); }
      if (MODE == 1) a[k] = __fadd2_rn(a[k], c);
      if (MODE == 2) { a[k].x = __fmaf_rn(a[k].x, c.x, c.y); a[k].y = __fmaf_rn(a[k].y, c.y, c.x); }
      if (MODE == 3) a[k] = __ffma2_rn(a[k], c, c);
      if (MODE == 4) { a[k].x = __fmul_rn(a[k].x, c.x); a[k].y = __fmul_rn(a[k].y, c.y); }
      if (MODE == 5) a[k] = __fmul2_rn(a[k], c);
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k].x + a[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float time_mode(float* out, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tput<MODE><<<148 * 8, 256>>>(out, 16);
  cudaEventRecord(e0);
  k_tput<MODE><<<148 * 8, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  unsigned long long *d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  const int iters = 4096;
  const char* names[6] = {"2x FADD scalar", "FADD2 packed", "2x FFMA scalar", "FFMA2 packed", "2x FMUL scalar", "FMUL2 packed"};
  float ms[6] = {time_mode<0>(out, iters), time_mode<1>(out, iters), time_mode<2>(out, iters), time_mode<3>(out, iters),
                 time_mode<4>(out, iters), time_mode<5>(out, iters)};
  double lane_ops = 148.0 * 8 * 256 * 8 * 2 * (double)iters;
  for (int m = 0; m < 6; ++m) printf("%-16s %8.3f ms  %7.2f T fp32-ops/s\n", names[m], ms[m], lane_ops / ms[m] / 1e9);
  k_div9_exhaustive<<<148 * 16, 256>>>(d, d + 1);
  k_div_random<<<148 * 16, 256>>>(1ull << 33, d + 2, 12345u);
  k_div_special<<<64, 256>>>(d + 3);
  unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("div9_fast vs __fdiv_rn(a,9): %llu mismatches of %llu floats\n", h[0], h[1]);
  printf("div_fast vs __fdiv_rn: %llu mismatches of %llu random pairs, %llu on near-all-ones divisors\n", h[2], 1ull << 33, h[3]);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
