// Per-pixel arithmetic of the view-synthesis loss path, written once for the device kernels.
//
// The forward chain is rounded step by step exactly like eager PyTorch-CUDA runs the reference
// (one rounding per Python-level op, FMA contraction only where a single ATen kernel contracts),
// because bilinear tap indices and the auto-mask arg-min are compared bit for bit.  Everything
// that matters for that is spelt with *_rn intrinsics so nvcc cannot re-associate or contract it.
// Reference formulas: layers.py:85-94 (disp_to_depth), :234-239 (BackprojectDepth), :253-264
// (Project3D), :318-332 (SSIM); trainer.py:500-501 (up-sample), :534-537 (grid_sample), :543-555.
//
// The header also compiles as plain C++ (tests/emul builds the tile logic for the host to check
// indexing and the analytic gradients without a GPU); the host build is test-only.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VSL_HD __host__ __device__ __forceinline__
#else
#define VSL_HD inline
#endif

namespace vsl {

#if defined(__CUDA_ARCH__)
VSL_HD float mul_rn(float a, float b) { return __fmul_rn(a, b); }
VSL_HD float add_rn(float a, float b) { return __fadd_rn(a, b); }
VSL_HD float sub_rn(float a, float b) { return __fsub_rn(a, b); }
VSL_HD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VSL_HD float div_rn(float a, float b) { return __fdiv_rn(a, b); }
VSL_HD float rcp_rn(float a) { return __frcp_rn(a); }
// gradients only (never on the bit-exact forward chain): MUFU.RCP without the IEEE fix-up and its slow-path
// branch; <= 1 ulp, far inside the gradient tolerance
VSL_HD float fast_rcp(float a) {
#if defined(VSL_EXACT_RCP)  // build knob for accuracy comparisons (tools/stress_parity.py)
  return __frcp_rn(a);
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
#endif
}
#else  // host build (tests/emul): compiled with -ffp-contract=off
VSL_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
VSL_HD float add_rn(float a, float b) { volatile float r = a + b; return r; }
VSL_HD float sub_rn(float a, float b) { volatile float r = a - b; return r; }
VSL_HD float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
VSL_HD float div_rn(float a, float b) { volatile float r = a / b; return r; }
VSL_HD float rcp_rn(float a) { volatile float r = 1.0f / a; return r; }
VSL_HD float fast_rcp(float a) { return 1.0f / a; }
#endif

// ---- image storage: fp32 or bf16 (arithmetic is always fp32; bf16 -> fp32 is exact) ---------------
struct bf16_t { uint16_t bits; };
VSL_HD float ldimg(const float* __restrict__ p, size_t i) { return p[i]; }
VSL_HD float ldimg(const bf16_t* __restrict__ p, size_t i) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float((uint32_t)p[i].bits << 16);
#else
  union { uint32_t u; float f; } c;
  c.u = (uint32_t)p[i].bits << 16;
  return c.f;
#endif
}

// ---- packed pairs: two independent fp32 values per operation -------------------------------------
// sm_100 issues add/mul/fma on fp32 PAIRS as one instruction (FADD2 / FMUL2 / FFMA2), each half rounded
// to nearest like the scalar op.  The kernel is issue-bound, so evaluating two source frames of one SSIM
// window in the two halves nearly halves the instruction count of the hot loop at identical bits.
#if defined(__CUDA_ARCH__)
typedef float2 F2;
VSL_HD F2 f2(float x, float y) { return make_float2(x, y); }
VSL_HD F2 add2(F2 a, F2 b) { return __fadd2_rn(a, b); }
VSL_HD F2 fma2(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }
// NOTE: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 although both carry .rn — even
// when the product is written fma(a, b, -0) — which it never does for the scalar forms.  Products that
// feed an add or subtract are therefore formed by two scalar __fmul_rn (mul2); only sums, explicit FMAs
// and products that are not added to anything (mul2_packed) use the packed instructions.  Checked bit
// for bit against the scalar chain on the B200 (tools/ubench/pairtest.cu).
VSL_HD F2 mul2(F2 a, F2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
VSL_HD F2 mul2_packed(F2 a, F2 b) { return __fmul2_rn(a, b); }
#else
struct F2 { float x, y; };
VSL_HD F2 f2(float x, float y) { F2 r; r.x = x; r.y = y; return r; }
VSL_HD F2 add2(F2 a, F2 b) { return f2(add_rn(a.x, b.x), add_rn(a.y, b.y)); }
VSL_HD F2 mul2(F2 a, F2 b) { return f2(mul_rn(a.x, b.x), mul_rn(a.y, b.y)); }
VSL_HD F2 fma2(F2 a, F2 b, F2 c) { return f2(fma_rn(a.x, b.x, c.x), fma_rn(a.y, b.y, c.y)); }
VSL_HD F2 mul2_packed(F2 a, F2 b) { return mul2(a, b); }
#endif
VSL_HD F2 splat(float v) { return f2(v, v); }
// s + a*b with the product rounded first (two roundings, like add_rn(s, mul_rn(a, b))), all packed:
// the product feeds only the MULTIPLICAND of an FMA whose other factor is a run-time 1.0 (GeoConst::one),
// so neither nvcc nor ptxas can contract or simplify it.  p*1 is exact, so the FMA rounds p + s once.
VSL_HD F2 addp(F2 s, F2 a, F2 b, F2 one) { return fma2(mul2_packed(a, b), one, s); }
VSL_HD F2 sub2(F2 a, F2 b) { return add2(a, f2(-b.x, -b.y)); }  // a + (-b) rounds exactly like a - b

// arithmetic-order selectors; mirror VSL_ARITH_* in include/vsl.h
enum : int {
  kTrueDiv = 1 << 0, kDotNoFma = 1 << 1, kDotReverse = 1 << 2, kUpsRight = 1 << 3,
  kUpsNoFma = 1 << 4, kTapNoFma = 1 << 5, kMeanDiv = 1 << 6, kDot3NoFma = 1 << 7, kDot3Reverse = 1 << 8,
  kDotKTNoFma = 1 << 9, kDotKTReverse = 1 << 10, kNormSeq = 1 << 11
};

// ---- bmm dot products ----------------------------------------------------------------------------
// cuBLAS SGEMM with K = 3 (rays, layers.py:235) / K = 4 (projection, layers.py:256).  For batch >= 2
// it accumulates one FMA chain, k ascending (the default here).  For batch 1 at some sizes cuBLAS
// picks a kernel that adds un-fused products; the host layer calibrates against torch.bmm once per
// shape and selects the matching variant (functional.calibrate_arith).
VSL_HD float dot3(float a0, float b0, float a1, float b1, float a2, float b2, int arith) {
  if (arith & kDot3NoFma) return add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2));
  if (arith & kDot3Reverse) return fma_rn(a0, b0, fma_rn(a1, b1, mul_rn(a2, b2)));
  return fma_rn(a2, b2, fma_rn(a1, b1, mul_rn(a0, b0)));
}
VSL_HD float dot4(float a0, float b0, float a1, float b1, float a2, float b2, float a3, float b3, int arith) {
  if (arith & kDotNoFma)
    return add_rn(add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2)), mul_rn(a3, b3));
  if (arith & kDotReverse) return fma_rn(a0, b0, fma_rn(a1, b1, fma_rn(a2, b2, mul_rn(a3, b3))));
  return fma_rn(a3, b3, fma_rn(a2, b2, fma_rn(a1, b1, mul_rn(a0, b0))));
}

// P = (K @ T)[:3,:] (layers.py:254): a [B,4,4] x [B,4,4] bmm, calibrated separately (its own cuBLAS kernel)
VSL_HD float dot4kt(float a0, float b0, float a1, float b1, float a2, float b2, float a3, float b3, int arith) {
  return dot4(a0, b0, a1, b1, a2, b2, a3, b3,
              ((arith & kDotKTNoFma) ? kDotNoFma : 0) | ((arith & kDotKTReverse) ? kDotReverse : 0));
}

// ---- F.interpolate(disp, [H,W], "bilinear", align_corners=False)  (trainer.py:500-501) --------
// Source index / weights as ATen's upsample_bilinear2d CUDA kernel computes them
// (torch/include/ATen/native/cuda/UpSample.cuh:115-130).
struct UpsTap { int i0, i1; float l0, l1; };
VSL_HD UpsTap ups_tap(int dst, int src_size, float scale) {
  float s = fma_rn(scale, (float)dst + 0.5f, -0.5f);
  if (s < 0.f) s = 0.f;
  UpsTap t;
  t.i0 = (int)s;
  t.i1 = t.i0 + ((t.i0 < src_size - 1) ? 1 : 0);
  t.l1 = sub_rn(s, (float)t.i0);
  t.l0 = sub_rn(1.0f, t.l1);
  return t;
}
VSL_HD float ups_combine(float l0, float a, float l1, float b, int arith) {
  if (arith & kUpsNoFma) return add_rn(mul_rn(l0, a), mul_rn(l1, b));
  if (arith & kUpsRight) return fma_rn(l1, b, mul_rn(l0, a));
  return fma_rn(l0, a, mul_rn(l1, b));
}
// disp: one image plane [hs, ws]; returns the up-sampled value at full-res pixel (v, u)
VSL_HD float upsample_disp(const float* __restrict__ disp, int hs, int ws, float scale_h, float scale_w,
                           bool identity, int v, int u, int arith) {
  if (identity) return disp[v * ws + u];
  UpsTap ty = ups_tap(v, hs, scale_h), tx = ups_tap(u, ws, scale_w);
  const float* r0 = disp + ty.i0 * ws;
  const float* r1 = disp + ty.i1 * ws;
  float top = ups_combine(tx.l0, r0[tx.i0], tx.l1, r0[tx.i1], arith);
  float bot = ups_combine(tx.l0, r1[tx.i0], tx.l1, r1[tx.i1], arith);
  return ups_combine(ty.l0, top, ty.l1, bot, arith);
}

// The same up-sample reading a staged window of disp_s (row stride `stride`, origin (cy0, cx0) in level
// coordinates; the staging clamps to the level, so every tap index ups_tap() can produce is inside it).
VSL_HD float upsample_disp_staged(const float* __restrict__ st, int stride, int cy0, int cx0, int hs, int ws,
                                  float scale_h, float scale_w, bool identity, int v, int u, int arith) {
  if (identity) return st[(v - cy0) * stride + (u - cx0)];
  UpsTap ty = ups_tap(v, hs, scale_h), tx = ups_tap(u, ws, scale_w);
  const float* r0 = st + (ty.i0 - cy0) * stride - cx0;
  const float* r1 = st + (ty.i1 - cy0) * stride - cx0;
  float top = ups_combine(tx.l0, r0[tx.i0], tx.l1, r0[tx.i1], arith);
  float bot = ups_combine(tx.l0, r1[tx.i0], tx.l1, r1[tx.i1], arith);
  return ups_combine(ty.l0, top, ty.l1, bot, arith);
}

// ---- asynchronous global -> shared copies (LDGSTS): issued early, awaited right before the data is read,
// so the global-memory latency overlaps the arithmetic of the phases in between.  Host build: plain copy.
VSL_HD void stage4(float* __restrict__ dst_shared, const float* __restrict__ src_global) {
#if defined(__CUDA_ARCH__)
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src_global) : "memory");
#else
  *dst_shared = *src_global;
#endif
}
VSL_HD void stage_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
VSL_HD void stage_wait() {  // all but the N most recently committed groups of this thread have landed
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---- geometry ---------------------------------------------------------------------------------
struct GeoConst {  // per call
  float min_disp, disp_range, eps;
  float one;               // 1.0f, deliberately a run-time value (see addp)
  float wm1, hm1;          // (float)(W-1), (float)(H-1)
  float inv_wm1, inv_hm1;  // 1.0f/(W-1), 1.0f/(H-1) rounded (PyTorch-CUDA divides by a CPU scalar this way)
  int W, H, arith;
};

struct Cam {  // camera-space point of a target pixel: X~ = (z*ray, 1)
  float z, rx, ry, rz, X, Y, Z;
};

// disp_to_depth (layers.py:90-93) + BackprojectDepth (layers.py:235-236); invK: row-major 4x4
VSL_HD float disp_to_z(float D, const GeoConst& g) { return rcp_rn(add_rn(mul_rn(g.disp_range, D), g.min_disp)); }
// camera point from a known depth z (invK: the first three rows of inv_K, row-major with stride 4)
VSL_HD Cam backproject_z(float z, const float* __restrict__ invK, int u, int v, const GeoConst& g) {
  Cam c;
  c.z = z;
  float fu = (float)u, fv = (float)v;
  c.rx = dot3(invK[0], fu, invK[1], fv, invK[2], 1.0f, g.arith);
  c.ry = dot3(invK[4], fu, invK[5], fv, invK[6], 1.0f, g.arith);
  c.rz = dot3(invK[8], fu, invK[9], fv, invK[10], 1.0f, g.arith);
  c.X = mul_rn(c.z, c.rx);
  c.Y = mul_rn(c.z, c.ry);
  c.Z = mul_rn(c.z, c.rz);
  return c;
}
VSL_HD Cam backproject_pixel(float D, const float* __restrict__ invK, int u, int v, const GeoConst& g) {
  return backproject_z(disp_to_z(D, g), invK, u, v, g);
}

struct Proj {       // Project3D + grid_sample coordinate chain for one source frame
  float gx, gy;     // outputs[("sample",f,s)] (normalised grid)
  float ix, iy;     // clipped sample position in pixels
  int x0, y0;       // north-west tap (the "projection indices")
  bool inx, iny;    // gradient passes through the clip (strictly inside)
};

// P: row-major 3x4 = (K@T)[:3,:]  (layers.py:254-263) then grid_sample's un-normalise + border clip
// (torch/include/ATen/native/cuda/GridSampler.cuh:23-31, 55-57), align_corners=True.
VSL_HD Proj project_pixel(const Cam& c, const float* __restrict__ P, const GeoConst& g) {
  float c0 = dot4(P[0], c.X, P[1], c.Y, P[2], c.Z, P[3], 1.0f, g.arith);
  float c1 = dot4(P[4], c.X, P[5], c.Y, P[6], c.Z, P[7], 1.0f, g.arith);
  float c2 = dot4(P[8], c.X, P[9], c.Y, P[10], c.Z, P[11], 1.0f, g.arith);
  float zeta = add_rn(c2, g.eps);
  float px = div_rn(c0, zeta), py = div_rn(c1, zeta);
  float nx = (g.arith & kTrueDiv) ? div_rn(px, g.wm1) : mul_rn(px, g.inv_wm1);
  float ny = (g.arith & kTrueDiv) ? div_rn(py, g.hm1) : mul_rn(py, g.inv_hm1);
  Proj r;
  r.gx = mul_rn(sub_rn(nx, 0.5f), 2.0f);
  r.gy = mul_rn(sub_rn(ny, 0.5f), 2.0f);
  float ix = mul_rn(mul_rn(add_rn(r.gx, 1.0f), 0.5f), g.wm1);
  float iy = mul_rn(mul_rn(add_rn(r.gy, 1.0f), 0.5f), g.hm1);
  r.inx = (ix > 0.f) && (ix < g.wm1);
  r.iny = (iy > 0.f) && (iy < g.hm1);
  r.ix = fminf(g.wm1, fmaxf(ix, 0.f));
  r.iy = fminf(g.hm1, fmaxf(iy, 0.f));
  r.x0 = (int)floorf(r.ix);
  r.y0 = (int)floorf(r.iy);
  return r;
}

struct Taps {  // bilinear weights of grid_sampler_2d (nw, ne, sw, se)
  float nw, ne, sw, se, wx0, wx1, wy0, wy1;
  bool x1ok, y1ok;
};
VSL_HD Taps bilinear_taps(const Proj& r, int W, int H) {
  Taps t;
  t.wx1 = sub_rn((float)(r.x0 + 1), r.ix);  // weight of the west column
  t.wx0 = sub_rn(r.ix, (float)r.x0);        // weight of the east column
  t.wy1 = sub_rn((float)(r.y0 + 1), r.iy);
  t.wy0 = sub_rn(r.iy, (float)r.y0);
  t.nw = mul_rn(t.wx1, t.wy1);
  t.ne = mul_rn(t.wx0, t.wy1);
  t.sw = mul_rn(t.wx1, t.wy0);
  t.se = mul_rn(t.wx0, t.wy0);
  t.x1ok = r.x0 + 1 < W;
  t.y1ok = r.y0 + 1 < H;
  return t;
}
VSL_HD float bilinear_value(const Taps& t, float vnw, float vne, float vsw, float vse, int arith) {
  if (arith & kTapNoFma) {
    float acc = mul_rn(vnw, t.nw);
    if (t.x1ok) acc = add_rn(acc, mul_rn(vne, t.ne));
    if (t.y1ok) acc = add_rn(acc, mul_rn(vsw, t.sw));
    if (t.x1ok && t.y1ok) acc = add_rn(acc, mul_rn(vse, t.se));
    return acc;
  }
  float acc = mul_rn(vnw, t.nw);
  if (t.x1ok) acc = fma_rn(vne, t.ne, acc);
  if (t.y1ok) acc = fma_rn(vsw, t.sw, acc);
  if (t.x1ok && t.y1ok) acc = fma_rn(vse, t.se, acc);
  return acc;
}

// ---- SSIM + L1 (layers.py:318-332, trainer.py:546-553) ------------------------------------------
VSL_HD float c1f() { return (float)(0.01 * 0.01); }
VSL_HD float c2f() { return (float)(0.03 * 0.03); }

// avg_pool2d's `sum / 9` (ATen AveragePool2d.cu divides the sequential window sum by the pool size).
// Correctly rounded a/9 in 3 FP instructions: q = a*RN(1/9), one Markstein correction with the exact
// residual.  Verified against __fdiv_rn(a, 9) for EVERY float with 1e-30 <= |a| <= 1e30 on the B200
// (tools/ubench/ubench.cu: 0 mismatches of 3,343,868,118); outside that range (and for 0) the IEEE
// division is used.
VSL_HD float div9(float a) {
#if defined(__CUDA_ARCH__)
  const float y = 1.0f / 9.0f;
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-9.0f, q, a);
  q = __fmaf_rn(r, y, q);
  // fast result kept iff the biased exponent is in [28, 225], i.e. 2^-99 <= |a| < 2^99 (inside the verified range)
  if (((__float_as_uint(a) & 0x7fffffffu) - 0x0e000000u) >= 0x63000000u) q = __fdiv_rn(a, 9.0f);
  return q;
#else
  return div_rn(a, 9.0f);
#endif
}

// The unguarded 3-instruction quotient and the guard as a key: the fast result is the correctly rounded
// a/9 iff div9_key(a) < kDiv9KeyLimit.  Callers that divide several sums take the maximum key and branch
// once for all of them (div9_all) instead of once per quotient.
constexpr unsigned kDiv9KeyLimit = 0x63000000u;
VSL_HD float div9_fast(float a) {
#if defined(__CUDA_ARCH__)
  const float y = 1.0f / 9.0f;
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-9.0f, q, a);
  return __fmaf_rn(r, y, q);
#else
  return div_rn(a, 9.0f);
#endif
}
VSL_HD unsigned div9_key(float a) {
#if defined(__CUDA_ARCH__)
  return (__float_as_uint(a) & 0x7fffffffu) - 0x0e000000u;
#else
  (void)a;
  return 0u;
#endif
}
VSL_HD unsigned umax2(unsigned a, unsigned b) { return a > b ? a : b; }
// q[i] = s[i] / 9 for N sums with one guard branch
template <int N>
VSL_HD void div9_all(const float (&s)[N], float (&q)[N]) {
  unsigned key = 0u;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    q[i] = div9_fast(s[i]);
    key = umax2(key, div9_key(s[i]));
  }
  if (key >= kDiv9KeyLimit) {  // some sum is zero, denormal-ish or huge: IEEE division for all (same bits where the fast path is valid)
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = div_rn(s[i], 9.0f);
  }
}

VSL_HD bool div9_fast_ok(float a) {
#if defined(__CUDA_ARCH__)
  return ((__float_as_uint(a) & 0x7fffffffu) - 0x0e000000u) < 0x63000000u;
#else
  (void)a;
  return false;
#endif
}
// both halves of a pair divided by 9 (same sequence and guard as div9)
VSL_HD F2 div9_2(F2 a) {
#if defined(__CUDA_ARCH__)
  const F2 y = splat(1.0f / 9.0f);
  F2 q = mul2_packed(a, y);  // q is only ever an FMA operand below, never the input of a plain add
  F2 r = fma2(splat(-9.0f), q, a);
  q = fma2(r, y, q);
  if (!div9_fast_ok(a.x)) q.x = __fdiv_rn(a.x, 9.0f);
  if (!div9_fast_ok(a.y)) q.y = __fdiv_rn(a.y, 9.0f);
  return q;
#else
  return f2(div_rn(a.x, 9.0f), div_rn(a.y, 9.0f));
#endif
}

// torch.mean over the 3 channels: sequential sum times float(1/3)
VSL_HD float mean3(float a, float b, float c, int arith) {
  float s = add_rn(add_rn(a, b), c);
  if (arith & kMeanDiv) return div_rn(s, 3.0f);
  return mul_rn(s, 1.0f / 3.0f);
}

struct SsimOut {
  float val;                 // clamp((1 - n/d)/2, 0, 1)
  float mu_x, n1, n2, d1, d2, r;
  bool live;                 // clamp passes gradient
};
// window sums are the row-major sequential sums ATen's avg_pool2d forms
// from the pooled means mu_x = sum x / 9, E[x^2], E[xy]
VSL_HD SsimOut ssim_from_means(float mu_x, float exx, float exy, float mu_y, float sig_y);
VSL_HD SsimOut ssim_from_sums(float sx, float sxx, float sxy, float mu_y, float sig_y) {
  return ssim_from_means(div9(sx), div9(sxx), div9(sxy), mu_y, sig_y);
}
VSL_HD SsimOut ssim_from_means(float mu_x, float exx, float exy, float mu_y, float sig_y) {
  SsimOut o;
  o.mu_x = mu_x;
  float mu_x2 = mul_rn(o.mu_x, o.mu_x);
  float sig_x = sub_rn(exx, mu_x2);
  float sig_xy = sub_rn(exy, mul_rn(o.mu_x, mu_y));
  o.n1 = add_rn(mul_rn(mul_rn(2.0f, o.mu_x), mu_y), c1f());
  o.n2 = add_rn(mul_rn(2.0f, sig_xy), c2f());
  o.d1 = add_rn(add_rn(mu_x2, mul_rn(mu_y, mu_y)), c1f());
  o.d2 = add_rn(add_rn(sig_x, sig_y), c2f());
  o.r = div_rn(mul_rn(o.n1, o.n2), mul_rn(o.d1, o.d2));
  float t = mul_rn(sub_rn(1.0f, o.r), 0.5f);
  o.live = (t >= 0.f) && (t <= 1.f);
  o.val = fminf(fmaxf(t, 0.f), 1.f);
  return o;
}

// The same SSIM value for two x-images against one y (the two halves are two source frames); op for op
// the chain of ssim_from_sums, so each half is bit-identical to the scalar evaluation.
// N pair quotients with one guard branch (both halves of every pair share it); same bits as div9_2
template <int N>
VSL_HD void div9_2_all(const F2 (&s)[N], F2 (&q)[N]) {
#if defined(__CUDA_ARCH__)
  const F2 y = splat(1.0f / 9.0f);
  unsigned key = 0u;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    F2 t = mul2_packed(s[i], y);  // only ever an FMA operand below
    F2 r = fma2(splat(-9.0f), t, s[i]);
    q[i] = fma2(r, y, t);
    key = umax2(key, umax2(div9_key(s[i].x), div9_key(s[i].y)));
  }
  if (key >= kDiv9KeyLimit) {
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = f2(__fdiv_rn(s[i].x, 9.0f), __fdiv_rn(s[i].y, 9.0f));
  }
#else
#pragma unroll
  for (int i = 0; i < N; ++i) q[i] = f2(div_rn(s[i].x, 9.0f), div_rn(s[i].y, 9.0f));
#endif
}
VSL_HD F2 ssim_val2_means(F2 mu_x, F2 exx, F2 exy, float mu_y, float sig_y, F2 one);
VSL_HD F2 ssim_val2(F2 sx, F2 sxx, F2 sxy, float mu_y, float sig_y, F2 one) {
  return ssim_val2_means(div9_2(sx), div9_2(sxx), div9_2(sxy), mu_y, sig_y, one);
}
VSL_HD F2 ssim_val2_means(F2 mu_x, F2 exx, F2 exy, float mu_y, float sig_y, F2 one) {
  const F2 muy = splat(mu_y), c1 = splat(c1f()), c2 = splat(c2f());
  const F2 mone = f2(-one.x, -one.y);
  F2 mu_x2 = mul2_packed(mu_x, mu_x);                       // consumed through FMAs by `one` only
  F2 sig_x = fma2(mu_x2, mone, exx);                        // E[x^2] - mu_x^2
  F2 sig_xy = fma2(mul2_packed(mu_x, muy), mone, exy);
  F2 n1 = addp(c1, add2(mu_x, mu_x), muy, one);             // (2 mu_x) mu_y + C1  (2 mu_x is exact)
  F2 n2 = add2(add2(sig_xy, sig_xy), c2);                   // 2 sigma_xy + C2
  F2 d1 = add2(fma2(mu_x2, one, splat(mul_rn(mu_y, mu_y))), c1);
  F2 d2 = add2(add2(sig_x, splat(sig_y)), c2);
  F2 n = mul2_packed(n1, n2), d = mul2_packed(d1, d2);
  F2 r = f2(div_rn(n.x, d.x), div_rn(n.y, d.y));
  F2 t = mul2_packed(sub2(splat(1.0f), r), splat(0.5f));
  return f2(fminf(fmaxf(t.x, 0.f), 1.f), fminf(fmaxf(t.y, 0.f), 1.f));
}
VSL_HD F2 mean3_2(F2 a, F2 b, F2 c, int arith) {
  F2 s = add2(add2(a, b), c);
  if (arith & kMeanDiv) return f2(div_rn(s.x, 3.0f), div_rn(s.y, 3.0f));
  return mul2_packed(s, splat(1.0f / 3.0f));  // callers combine it through addp, never a plain add
}

// d r / d(mu_x, E[x^2], E[xy]) of r = n1 n2 / (d1 d2)  
VSL_HD void ssim_r_grads(const SsimOut& o, float mu_y, float& dmu, float& dexx, float& dexy) {
  float inv_d = fast_rcp(o.d1 * o.d2);
  dexy = 2.0f * o.n1 * inv_d;
  dexx = -o.r * o.d1 * inv_d;  // -r/d2
  dmu = inv_d * (2.0f * mu_y * (o.n2 - o.n1) - o.r * 2.0f * o.mu_x * (o.d2 - o.d1));
}

// ---- pose network output -> 4x4 (layers.py:97-172), rounded like the reference's torch ops on CUDA ------
// torch.norm over the 3 contiguous components reduces across 4 lanes with a shuffle tree:
// (x0^2 + x2^2) + x1^2 (probed on the B200, tools/probe_pose.py; kNormSeq selects the sequential order).
// M = T(t) R, or inverted R^T T(-t): every product with the 0/1 entries is exact, only the inverted
// translation column is a real 3-term dot product (accumulated in the calibrated 4x4x4 bmm order).
VSL_HD void pose_matrix(const float v[3], const float t[3], bool invert, int arith, float M[16]) {
  const float q0 = mul_rn(v[0], v[0]), q1 = mul_rn(v[1], v[1]), q2 = mul_rn(v[2], v[2]);
  const float ss = (arith & kNormSeq) ? add_rn(add_rn(q0, q1), q2) : add_rn(add_rn(q0, q2), q1);
#if defined(__CUDA_ARCH__)
  const float angle = __fsqrt_rn(ss);
#else
  const float angle = sqrtf(ss);
#endif
  const float den = add_rn(angle, 1e-7f);
  const float x = div_rn(v[0], den), y = div_rn(v[1], den), z = div_rn(v[2], den);
  const float ca = cosf(angle), sa = sinf(angle);
  const float C = sub_rn(1.0f, ca);
  const float xs = mul_rn(x, sa), ys = mul_rn(y, sa), zs = mul_rn(z, sa);
  const float xC = mul_rn(x, C), yC = mul_rn(y, C), zC = mul_rn(z, C);
  const float xyC = mul_rn(x, yC), yzC = mul_rn(y, zC), zxC = mul_rn(z, xC);
  float R[9];
  R[0] = add_rn(mul_rn(x, xC), ca); R[1] = sub_rn(xyC, zs);           R[2] = add_rn(zxC, ys);
  R[3] = add_rn(xyC, zs);           R[4] = add_rn(mul_rn(y, yC), ca); R[5] = sub_rn(yzC, xs);
  R[6] = sub_rn(zxC, ys);           R[7] = add_rn(yzC, xs);           R[8] = add_rn(mul_rn(z, zC), ca);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i * 4 + j] = invert ? R[j * 3 + i] : R[i * 3 + j];
    M[i * 4 + 3] = invert ? dot4kt(R[i], -t[0], R[3 + i], -t[1], R[6 + i], -t[2], 0.0f, 1.0f, arith) : t[i];
    M[12 + i] = 0.0f;
  }
  M[15] = 1.0f;
}

// reflect-pad-1 index map (nn.ReflectionPad2d(1), layers.py:313): -1 -> 1, n -> n-2
VSL_HD int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

}  // namespace vsl
