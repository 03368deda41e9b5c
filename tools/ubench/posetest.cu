// Developer probe: can a CUDA kernel reproduce torch's axis-angle -> rotation entries bit for bit?
// Compiled to a .so; tools/probe_pose.py feeds it torch's inputs and compares with torch's outputs.
#include <cuda_runtime.h>
extern "C" __global__ void k_rot(int n, int variant, const float* __restrict__ v, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x0 = v[3 * i], x1 = v[3 * i + 1], x2 = v[3 * i + 2];
  float ss;
  if (variant == 0) ss = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2));
  else if (variant == 1) ss = __fmaf_rn(x2, x2, __fmaf_rn(x1, x1, __fmul_rn(x0, x0)));
  else if (variant == 2) ss = __fadd_rn(__fmul_rn(x0, x0), __fadd_rn(__fmul_rn(x1, x1), __fmul_rn(x2, x2)));
  else ss = __fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x2, x2)), __fmul_rn(x1, x1));  // 4-lane shuffle tree
  float angle = __fsqrt_rn(ss);
  float den = __fadd_rn(angle, 1e-7f);
  float ax = __fdiv_rn(x0, den), ay = __fdiv_rn(x1, den), az = __fdiv_rn(x2, den);
  float ca = cosf(angle), sa = sinf(angle);
  float C = __fsub_rn(1.0f, ca);
  float xs = __fmul_rn(ax, sa), ys = __fmul_rn(ay, sa), zs = __fmul_rn(az, sa);
  float xC = __fmul_rn(ax, C), yC = __fmul_rn(ay, C), zC = __fmul_rn(az, C);
  float xyC = __fmul_rn(ax, yC), yzC = __fmul_rn(ay, zC), zxC = __fmul_rn(az, xC);
  float* o = out + 12 * i;
  o[0] = angle; o[1] = ca; o[2] = sa;
  o[3] = __fadd_rn(__fmul_rn(ax, xC), ca); o[4] = __fsub_rn(xyC, zs); o[5] = __fadd_rn(zxC, ys);
  o[6] = __fadd_rn(xyC, zs); o[7] = __fadd_rn(__fmul_rn(ay, yC), ca); o[8] = __fsub_rn(yzC, xs);
  o[9] = __fsub_rn(zxC, ys); o[10] = __fadd_rn(yzC, xs); o[11] = __fadd_rn(__fmul_rn(az, zC), ca);
}
extern "C" int run_rot(int n, int variant, const float* v, float* out, void* stream) {
  k_rot<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, variant, v, out);
  return (int)cudaGetLastError();
}
