// On-GPU colour augmentation of the ``color_aug`` inputs (the dataset-side remainder).
//
// The reference augments on the CPU, per 8-bit PIL level (datasets/mono_dataset2.py:92-97, :124):
//   transforms.Compose([ColorJitter(brightness, contrast, saturation, hue), RandomHorizontalFlip(0.5),
//                       RandomAutocontrast()])
// then transforms.ToTensor().  The arithmetic is torchvision's _functional_pil.py on top of Pillow's C code
// (Image.blend, convert("L" | "HSV" | "RGB"), ImageOps.autocontrast).  The kernels below reproduce those byte
// for byte -- C-float blends without contraction, Pillow's mixed float/double HSV round trip evaluated with
// explicitly rounded fp64 intrinsics, Python-double LUT for autocontrast -- given the random draws, which stay
// on the host (they come from torch's global generator; input_pipeline.draw_color_aug_params restates the order).
//
// Per call (one batch of equally sized 8-bit images, each with its own draws): four small launches,
//   k_aug_init     per-image statistics := neutral
//   k_aug_stats    images whose chain contains `contrast`: sum of the grey levels of the image as it is when the
//                  contrast step is reached (integer atomics: order-independent, reproducible)
//   k_aug_apply    the whole jitter chain per pixel -> 8-bit image; per-channel min / max for autocontrast
//   k_aug_finish   flip + autocontrast LUT + ToTensor
// Byte work on a few MB: latency-bound launches, HBM-bound bodies; no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsl.h"
#include "vsl_math.cuh"

namespace vsl {

extern thread_local int g_last_cuda_error;

#define VSL_CUDA_OK_AUG(expr)                     \
  do {                                            \
    cudaError_t e__ = (expr);                     \
    if (e__ != cudaSuccess) {                     \
      g_last_cuda_error = (int)e__;               \
      return VSL_ERR_CUDA;                        \
    }                                             \
  } while (0)

struct AugStats {          // one per image, in the workspace
  unsigned long long gray_sum;
  unsigned lo[3], hi[3];
};
static_assert(sizeof(AugStats) == 32, "AugStats layout");

struct Rgb { int r, g, b; };

// Image.blend(im1, im2, alpha) for one sample (libImaging/Blend.c): C float arithmetic, truncating cast;
// outside [0, 1] the float is clipped first
__device__ __forceinline__ int blend1(int in1, int in2, float alpha) {
  const float t = __fadd_rn((float)in1, __fmul_rn(alpha, (float)(in2 - in1)));
  if (alpha >= 0.0f && alpha <= 1.0f) return (int)t & 255;
  return t <= 0.0f ? 0 : (t >= 255.0f ? 255 : (int)t);
}
__device__ __forceinline__ Rgb blend3(Rgb a, Rgb b, float alpha) {
  Rgb o;
  o.r = blend1(a.r, b.r, alpha); o.g = blend1(a.g, b.g, alpha); o.b = blend1(a.b, b.b, alpha);
  return o;
}
// convert("L"), ITU-R 601-2 with 16 fractional bits (libImaging/Convert.c)
__device__ __forceinline__ int gray_of(Rgb v) {
  return (int)(((unsigned)v.r * 19595u + (unsigned)v.g * 38470u + (unsigned)v.b * 7471u + 0x8000u) >> 16);
}
__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// Per-CTA tables of the two quotients of hsv2rgb that depend on one byte only (built once per CTA with the exact
// double divisions, read per pixel): i = floor(h * 6.0 / 255.0), f = float(h * 6.0 / 255.0 - i), fs = float(s / 255.0)
struct HueLut {
  double f[256], fs[256];
  uint8_t i[256];
};
__device__ __forceinline__ void build_hue_lut(HueLut& lut, int tid, int nthreads) {
  for (int k = tid; k < 256; k += nthreads) {
    const double hf = __ddiv_rn(__dmul_rn((double)k, 6.0), 255.0);
    const int i = (int)floor(hf);
    lut.i[k] = (uint8_t)i;
    lut.f[k] = (double)__double2float_rn(__dsub_rn(hf, (double)i));
    lut.fs[k] = (double)__double2float_rn(__ddiv_rn((double)k, 255.0));
  }
}

// convert("HSV") (Convert.c rgb2hsv_row): float ratios, double hue wrap
__device__ __forceinline__ Rgb rgb_to_hsv(Rgb v) {
  const int maxc = max(v.r, max(v.g, v.b)), minc = min(v.r, min(v.g, v.b));
  Rgb o;
  o.b = maxc;
  if (minc == maxc) { o.r = 0; o.g = 0; return o; }
  const float cr = (float)(maxc - minc);
  const float s = __fdiv_rn(cr, (float)maxc);
  const float rc = __fdiv_rn((float)(maxc - v.r), cr);
  const float gc = __fdiv_rn((float)(maxc - v.g), cr);
  const float bc = __fdiv_rn((float)(maxc - v.b), cr);
  float h;
  if (v.r == maxc) h = __fsub_rn(bc, gc);
  else if (v.g == maxc) h = __double2float_rn(__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc));
  else h = __double2float_rn(__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc));
  // h / 6.0 as q = h * RN(1/6) with one fused correction of the exact residual: the correctly rounded quotient for
  // every h this function can produce (tests/test_gpu_color_aug.py compares all 2^24 RGB triples with the oracle)
  const double x = (double)h, c6 = 1.0 / 6.0;
  double q = __dmul_rn(x, c6);
  q = __fma_rn(__fma_rn(-6.0, q, x), c6, q);
  double w = __dadd_rn(q, 1.0);   // in (0.8, 2): fmod(w, 1.0) = w - floor(w), exact
  w = w - floor(w);
  h = __double2float_rn(w);
  o.r = clip8((int)__dmul_rn((double)h, 255.0));
  o.g = clip8((int)__dmul_rn((double)s, 255.0));
  return o;
}
// HSV -> RGB (Convert.c hsv2rgb, "following colorsys.py"); (h, s, v) in (r, g, b)
__device__ __forceinline__ Rgb hsv_to_rgb(Rgb hsv, const HueLut& lut) {
  const int h = hsv.r, s = hsv.g, v = hsv.b;
  Rgb o;
  if (s == 0) { o.r = o.g = o.b = v; return o; }
  const int i = lut.i[h];
  const double f = lut.f[h], fs = lut.fs[s];
  const double vf = (double)v;
  const int p = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, fs)), 0.5)));
  const int q = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, f))), 0.5)));
  const int t = clip8((int)floor(__dadd_rn(__dmul_rn(vf, __dsub_rn(1.0, __dmul_rn(fs, __dsub_rn(1.0, f)))), 0.5)));
  switch (i == 6 ? 0 : i) {   // i % 6 for i in 0..6
    case 0: o.r = v; o.g = t; o.b = p; break;
    case 1: o.r = q; o.g = v; o.b = p; break;
    case 2: o.r = p; o.g = v; o.b = t; break;
    case 3: o.r = p; o.g = q; o.b = v; break;
    case 4: o.r = t; o.g = p; o.b = v; break;
    default: o.r = v; o.g = p; o.b = q; break;
  }
  return o;
}

// position of the contrast step in the image's chain, or -1
__device__ __forceinline__ int contrast_pos(const VslAugParams& prm) {
  if (!prm.enabled) return -1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (prm.order[k] == 1) return k;
  return -1;
}
// the first `upto` steps of the jitter chain (ColorJitter.forward's loop over fn_idx) on N pixels at once: the step is
// chosen once (it is uniform over the image) and applied to the N independent pixels back to back, so their long
// fp64 / division chains overlap
template <int N>
__device__ __forceinline__ void apply_chain(Rgb (&v)[N], const VslAugParams& prm, int upto, int mean, const HueLut& lut) {
  for (int k = 0; k < upto; ++k) {
    const int fn = prm.order[k];
    if (fn == 0) {  // adjust_brightness: blend(black, img, factor)
      Rgb z; z.r = z.g = z.b = 0;
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = blend3(z, v[j], prm.factor[0]);
    } else if (fn == 1) {  // adjust_contrast: blend(mean grey, img, factor)
      Rgb z; z.r = z.g = z.b = mean;
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = blend3(z, v[j], prm.factor[1]);
    } else if (fn == 2) {  // adjust_saturation: blend(grey image, img, factor)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        Rgb z; z.r = z.g = z.b = gray_of(v[j]);
        v[j] = blend3(z, v[j], prm.factor[2]);
      }
    } else if (fn == 3) {  // adjust_hue: H += shift (8-bit wrap) in HSV
      Rgb hsv[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        hsv[j] = rgb_to_hsv(v[j]);
        hsv[j].r = (hsv[j].r + prm.hue_shift) & 255;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = hsv_to_rgb(hsv[j], lut);
    }
  }
}
// int(ImageStat.Stat(grey).mean[0] + 0.5): Python float division of the two integer sums
__device__ __forceinline__ int contrast_mean(unsigned long long gray_sum, int hw) {
  return (int)__dadd_rn(__ddiv_rn((double)gray_sum, (double)hw), 0.5);
}

constexpr int kAugNT = 256;
constexpr int kAugPer = 4;   // pixels per thread

__global__ void k_aug_init(AugStats* st, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  st[b].gray_sum = 0ull;
#pragma unroll
  for (int c = 0; c < 3; ++c) { st[b].lo[c] = 255u; st[b].hi[c] = 0u; }
}

__global__ void __launch_bounds__(kAugNT) k_aug_stats(const uint8_t* __restrict__ in, const VslAugParams* __restrict__ params,
                                                      AugStats* __restrict__ st, int hw) {
  const int b = blockIdx.y;
  const VslAugParams prm = params[b];
  const int cpos = contrast_pos(prm);
  if (cpos < 0) return;  // block-uniform
  __shared__ HueLut lut;
  build_hue_lut(lut, threadIdx.x, kAugNT);
  __syncthreads();
  const uint8_t* img = in + (size_t)b * hw * 3;
  unsigned sum = 0;
  const int base = (blockIdx.x * kAugNT + threadIdx.x) * kAugPer;
  Rgb v[kAugPer];
#pragma unroll
  for (int k = 0; k < kAugPer; ++k) {
    const int i = min(base + k, hw - 1);   // past the end: a duplicate of the last pixel, not counted below
    v[k].r = img[3 * i]; v[k].g = img[3 * i + 1]; v[k].b = img[3 * i + 2];
  }
  apply_chain<kAugPer>(v, prm, cpos, 0, lut);
#pragma unroll
  for (int k = 0; k < kAugPer; ++k)
    if (base + k < hw) sum += (unsigned)gray_of(v[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __shared__ unsigned wsum[kAugNT / 32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
#pragma unroll
    for (int i = 0; i < kAugNT / 32; ++i) tot += wsum[i];
    atomicAdd(&st[b].gray_sum, tot);
  }
}

__global__ void __launch_bounds__(kAugNT) k_aug_apply(const uint8_t* __restrict__ in, const VslAugParams* __restrict__ params,
                                                      AugStats* __restrict__ st, uint8_t* __restrict__ mid, int hw) {
  const int b = blockIdx.y;
  const VslAugParams prm = params[b];
  const uint8_t* img = in + (size_t)b * hw * 3;
  uint8_t* dst = mid + (size_t)b * hw * 3;
  const int mean = contrast_pos(prm) >= 0 ? contrast_mean(st[b].gray_sum, hw) : 0;
  __shared__ HueLut lut;
  if (prm.enabled) build_hue_lut(lut, threadIdx.x, kAugNT);   // block-uniform
  __syncthreads();
  unsigned lo[3] = {255u, 255u, 255u}, hi[3] = {0u, 0u, 0u};
  const int base = (blockIdx.x * kAugNT + threadIdx.x) * kAugPer;
  Rgb v[kAugPer];
#pragma unroll
  for (int k = 0; k < kAugPer; ++k) {
    const int i = min(base + k, hw - 1);   // past the end: a duplicate of the last pixel, neither stored nor counted
    v[k].r = img[3 * i]; v[k].g = img[3 * i + 1]; v[k].b = img[3 * i + 2];
  }
  if (prm.enabled) apply_chain<kAugPer>(v, prm, 4, mean, lut);
#pragma unroll
  for (int k = 0; k < kAugPer; ++k) {
    const int i = base + k;
    if (i < hw) {
      dst[3 * i] = (uint8_t)v[k].r; dst[3 * i + 1] = (uint8_t)v[k].g; dst[3 * i + 2] = (uint8_t)v[k].b;
      lo[0] = min(lo[0], (unsigned)v[k].r); hi[0] = max(hi[0], (unsigned)v[k].r);
      lo[1] = min(lo[1], (unsigned)v[k].g); hi[1] = max(hi[1], (unsigned)v[k].g);
      lo[2] = min(lo[2], (unsigned)v[k].b); hi[2] = max(hi[2], (unsigned)v[k].b);
    }
  }
  if (!(prm.enabled && prm.autocontrast)) return;  // block-uniform
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[c] = min(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
      hi[c] = max(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { atomicMin(&st[b].lo[c], lo[c]); atomicMax(&st[b].hi[c], hi[c]); }
  }
}

template <class Out> __device__ __forceinline__ void aug_store(Out* p, size_t i, int v);
// transforms.ToTensor(): byte.div(255), correctly rounded in three instructions (see div255 in vsl_input.cu)
__device__ __forceinline__ float aug_div255(int v) {
  const float a = (float)v, y = 1.0f / 255.0f;
  const float q = __fmul_rn(a, y);
  return __fmaf_rn(__fmaf_rn(-255.0f, q, a), y, q);
}
template <> __device__ __forceinline__ void aug_store<float>(float* p, size_t i, int v) {
  p[i] = aug_div255(v);
}
template <> __device__ __forceinline__ void aug_store<bf16_t>(bf16_t* p, size_t i, int v) {
  const uint32_t u = __float_as_uint(aug_div255(v));  // round-to-nearest-even, like Tensor.bfloat16()
  p[i].bits = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

// ImageOps.autocontrast's lookup table entry (cutoff 0), Python double arithmetic
__device__ __forceinline__ int autocontrast1(int v, int lo, int hi) {
  if (hi <= lo) return v;
  const double scale = __ddiv_rn(255.0, (double)(hi - lo));
  const double offset = __dmul_rn(-(double)lo, scale);
  return clip8((int)__dadd_rn(__dmul_rn((double)v, scale), offset));
}

template <class Out>
__global__ void __launch_bounds__(kAugNT) k_aug_finish(const uint8_t* __restrict__ mid, const VslAugParams* __restrict__ params,
                                                       const AugStats* __restrict__ st, Out* __restrict__ out,
                                                       uint8_t* __restrict__ out_u8, int hw, int W) {
  const int b = blockIdx.y;
  const VslAugParams prm = params[b];
  const bool on = prm.enabled != 0, fl = on && prm.flip, ac = on && prm.autocontrast;
  __shared__ uint8_t aclut[3][256];   // ImageOps.autocontrast's lookup tables of this image, one entry per thread and band
  if (ac) {  // block-uniform
#pragma unroll
    for (int c = 0; c < 3; ++c)
      for (int k = threadIdx.x; k < 256; k += kAugNT) aclut[c][k] = (uint8_t)autocontrast1(k, (int)st[b].lo[c], (int)st[b].hi[c]);
  }
  __syncthreads();
  const uint8_t* img = mid + (size_t)b * hw * 3;
  Out* dst = out ? out + (size_t)b * 3 * hw : nullptr;
  const int base = (blockIdx.x * kAugNT + threadIdx.x) * kAugPer;
#pragma unroll
  for (int k = 0; k < kAugPer; ++k) {
    const int i = base + k;
    if (i >= hw) continue;
    int src = i;
    if (fl) {
      const int y = i / W, x = i - y * W;
      src = y * W + (W - 1 - x);
    }
    int v[3] = {img[3 * src], img[3 * src + 1], img[3 * src + 2]};
    if (ac) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = aclut[c][v[c]];
    }
    if (dst) {
#pragma unroll
      for (int c = 0; c < 3; ++c) aug_store<Out>(dst, (size_t)c * hw + i, v[c]);
    }
    if (out_u8) {
      uint8_t* q = out_u8 + ((size_t)b * hw + i) * 3;
      q[0] = (uint8_t)v[0]; q[1] = (uint8_t)v[1]; q[2] = (uint8_t)v[2];
    }
  }
}

static size_t aug_stats_bytes(int batch) { return ((size_t)batch * sizeof(AugStats) + 255) / 256 * 256; }

}  // namespace vsl

using namespace vsl;

extern "C" {

size_t vsl_color_aug_workspace_bytes(int batch, int height, int width) {
  if (batch < 1 || height < 1 || width < 1) return 0;
  return aug_stats_bytes(batch) + (size_t)batch * height * width * 3;
}

int vsl_color_aug_forward(int batch, int height, int width, int out_dtype, const uint8_t* frames_hwc,
                          const VslAugParams* params, void* out, uint8_t* out_u8, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (batch < 1 || height < 1 || width < 1) return VSL_ERR_BAD_DESC;
  if ((size_t)height * width > (size_t)1 << 28) return VSL_ERR_BAD_DESC;  // the grey sum of an image stays far below 2^53
  if (out_dtype != VSL_DTYPE_F32 && out_dtype != VSL_DTYPE_BF16) return VSL_ERR_UNSUPPORTED;
  if (!frames_hwc || !params || !workspace || (!out && !out_u8)) return VSL_ERR_NULL_POINTER;
  if (((uintptr_t)workspace & 255u) != 0 || ((uintptr_t)params & 3u) != 0) return VSL_ERR_MISALIGNED;
  if (workspace_bytes < vsl_color_aug_workspace_bytes(batch, height, width)) return VSL_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  AugStats* stats = (AugStats*)workspace;
  uint8_t* mid = (uint8_t*)workspace + aug_stats_bytes(batch);
  const int hw = height * width;
  const dim3 grid((hw + kAugNT * kAugPer - 1) / (kAugNT * kAugPer), batch);
  k_aug_init<<<(batch + 63) / 64, 64, 0, st>>>(stats, batch);
  k_aug_stats<<<grid, kAugNT, 0, st>>>(frames_hwc, params, stats, hw);
  k_aug_apply<<<grid, kAugNT, 0, st>>>(frames_hwc, params, stats, mid, hw);
  if (out_dtype == VSL_DTYPE_BF16)
    k_aug_finish<bf16_t><<<grid, kAugNT, 0, st>>>(mid, params, stats, (bf16_t*)out, out_u8, hw, width);
  else
    k_aug_finish<float><<<grid, kAugNT, 0, st>>>(mid, params, stats, (float*)out, out_u8, hw, width);
  VSL_CUDA_OK_AUG(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
