// On-GPU input pipeline for the loss path: 8-bit frames -> the ("color", f, s) pyramid.
//
// The reference builds the pyramid on the CPU in its dataset (datasets/mono_dataset2.py:85-89, 103-124):
// level s = transforms.Resize((H // 2^s, W // 2^s), Image.ANTIALIAS)(level s-1) on PIL images, then
// transforms.ToTensor() (HWC uint8 -> CHW float32 / 255), and ships fp32 tensors to the GPU
// (trainer.py:373-374).  Here the host ships the 8-bit level-0 frames (a quarter of the bytes, and no
// pyramid) and the GPU reproduces Pillow's 8-bit LANCZOS resampling bit for bit:
//   * coefficient tables: Pillow's precompute_coeffs + normalize_coeffs_8bpc, evaluated on the host in
//     double precision exactly like Pillow's C code (vsl_pyramid_plan, once per shape);
//   * k_lanczos_half: horizontal pass (rounded to 8 bits) then vertical pass (rounded to 8 bits) of one
//     output tile, integer arithmetic with 22 fractional bits, like ImagingResampleHorizontal/Vertical_8bpc;
//   * ToTensor: v / 255 as an IEEE division.
// Byte/integer work, HBM-bound and tiny next to the loss kernels; no tensor cores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/vsl.h"
#include "vsl_math.cuh"

namespace vsl {

extern thread_local int g_last_cuda_error;

#define VSL_CUDA_OK_IN(expr)                      \
  do {                                            \
    cudaError_t e__ = (expr);                     \
    if (e__ != cudaSuccess) {                     \
      g_last_cuda_error = (int)e__;               \
      return VSL_ERR_CUDA;                        \
    }                                             \
  } while (0)

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS
constexpr int kKsize = 13;                  // ceil(3 * 2) * 2 + 1 taps reserved per output for a 2:1 LANCZOS

// One axis of one level: for output index o, input taps [lo, lo + n) with integer weights w[o][0..n).
struct AxisTable {
  const int32_t* bounds;  // [out][2] = (lo, n)
  const int32_t* coefs;   // [out][kKsize]
};

struct PyramidPlan {  // offsets (bytes) into the caller's workspace; identical in workspace_bytes() and the launchers
  size_t off_xb[VSL_MAX_SCALES], off_xc[VSL_MAX_SCALES], off_yb[VSL_MAX_SCALES], off_yc[VSL_MAX_SCALES];
  size_t off_u8[VSL_MAX_SCALES];  // 8-bit level s >= 1, HWC
  size_t total;
};

static bool pyr_desc_ok(const VslPyramidDesc* d) {
  if (!d || d->abi_version != VSL_ABI_VERSION) return false;
  if (d->batch < 1 || d->num_levels < 1 || d->num_levels > VSL_MAX_SCALES) return false;
  if (d->height < 2 || d->width < 2) return false;
  const int e = d->num_levels - 1;
  if (((d->height >> e) << e) != d->height || ((d->width >> e) << e) != d->width) return false;
  if ((d->height >> e) < 1 || (d->width >> e) < 1) return false;
  if (d->out_dtype != VSL_DTYPE_F32 && d->out_dtype != VSL_DTYPE_BF16) return false;
  return true;
}

static PyramidPlan make_pyr_plan(const VslPyramidDesc* d) {
  PyramidPlan pl = {};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  for (int s = 1; s < d->num_levels; ++s) {
    const int hs = d->height >> s, ws = d->width >> s;
    pl.off_xb[s] = take((size_t)ws * 2 * 4);
    pl.off_xc[s] = take((size_t)ws * kKsize * 4);
    pl.off_yb[s] = take((size_t)hs * 2 * 4);
    pl.off_yc[s] = take((size_t)hs * kKsize * 4);
    pl.off_u8[s] = take((size_t)d->batch * hs * ws * 3);
  }
  pl.total = off;
  return pl;
}

// Pillow's lanczos_filter / sinc_filter (Resample.c), double precision
static double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}
// Pillow's precompute_coeffs (box = the whole axis) followed by normalize_coeffs_8bpc.  `stride`: words reserved per
// output in `coefs` (>= axis_ksize; the tail stays zero)
static int axis_ksize(int in_size, int out_size) {
  const double scale = (double)in_size / (double)out_size;
  return (int)ceil(3.0 * (scale < 1.0 ? 1.0 : scale)) * 2 + 1;
}
static void axis_coeffs_stride(int in_size, int out_size, int stride, std::vector<int32_t>& bounds, std::vector<int32_t>& coefs) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 3.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  coefs.assign((size_t)out_size * stride, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      coefs[(size_t)xx * stride + x] = k[x] < 0 ? (int32_t)(-0.5 + k[x] * (1 << kPrecisionBits))
                                                : (int32_t)(0.5 + k[x] * (1 << kPrecisionBits));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}
static bool axis_coeffs(int in_size, int out_size, std::vector<int32_t>& bounds, std::vector<int32_t>& coefs) {
  if (axis_ksize(in_size, out_size) > kKsize) return false;
  axis_coeffs_stride(in_size, out_size, kKsize, bounds, coefs);
  return true;
}

__device__ __forceinline__ uint8_t clip8(int acc) {  // Pillow: clip8_lookups[acc >> PRECISION_BITS]
  int v = acc >> kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// transforms.ToTensor(): byte.div(255) as a correctly rounded fp32 quotient in three instructions: q = v * RN(1/255)
// and one Markstein correction with the exact residual.  Equal to __fdiv_rn(v, 255) for all 256 bytes (checked
// exhaustively on the host, and by every byte-exact test of the pipeline); the IEEE division costs ~10 instructions
// and made the conversion kernels instruction-bound.
__device__ __forceinline__ float div255(uint8_t v) {
  const float a = (float)v, y = 1.0f / 255.0f;
  const float q = __fmul_rn(a, y);
  return __fmaf_rn(__fmaf_rn(-255.0f, q, a), y, q);
}
template <class Out> __device__ __forceinline__ void store_tensor(Out* p, size_t i, uint8_t v);
template <> __device__ __forceinline__ void store_tensor<float>(float* p, size_t i, uint8_t v) {
  p[i] = div255(v);
}
template <> __device__ __forceinline__ void store_tensor<bf16_t>(bf16_t* p, size_t i, uint8_t v) {
  // bf16 image storage (BASELINE config 3): round-to-nearest-even of the fp32 value, like Tensor.bfloat16()
  const uint32_t u = __float_as_uint(div255(v));
  p[i].bits = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

// level 0: HWC uint8 -> CHW tensor
// `flip` (optional, [B] bytes): images whose flag is set are mirrored left-right while they are read, like
// `color.transpose(Image.FLIP_LEFT_RIGHT)` in MonoDataset.get_color (datasets/mono_dataset2.py via kitti_dataset.py)
template <class Out>
__global__ void __launch_bounds__(256) k_u8_to_tensor(const uint8_t* __restrict__ in, Out* __restrict__ out, int hw,
                                                     size_t total_px, int W, const uint8_t* __restrict__ flip) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // pixel index over [B, H*W]
  if (i >= total_px) return;
  const size_t b = i / hw, o = i - b * hw;
  size_t src = i;
  if (flip && flip[b]) {
    const int y = (int)(o / W), x = (int)(o - (size_t)y * W);
    src = b * hw + (size_t)y * W + (W - 1 - x);
  }
  const uint8_t* q = in + src * 3;
  Out* dst = out + b * 3 * hw + o;
  store_tensor<Out>(dst, 0, q[0]);
  store_tensor<Out>(dst, (size_t)hw, q[1]);
  store_tensor<Out>(dst, (size_t)2 * hw, q[2]);
}

// the same for four consecutive pixels per thread: 3 x 32-bit loads, one float4 (or 4 x bf16) store per plane.
// Needs H*W % 4 == 0 (then every 12-byte group and every plane segment is aligned).
__device__ __forceinline__ void store4(float* p, const uint8_t v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(div255(v[0]), div255(v[1]), div255(v[2]), div255(v[3]));
}
__device__ __forceinline__ void store4(bf16_t* p, const uint8_t v[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) store_tensor<bf16_t>(p, i, v[i]);
}
template <class Out>
__global__ void __launch_bounds__(256) k_u8_to_tensor_x4(const uint8_t* __restrict__ in, Out* __restrict__ out, int hw,
                                                        size_t total_quads, int W, const uint8_t* __restrict__ flip) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;  // group of 4 pixels
  if (i >= total_quads) return;
  const size_t px = i * 4;
  const size_t b = px / hw, o = px - b * hw;
  const bool fl = flip && flip[b];  // only with W % 4 == 0 (the launcher checks): a quad never straddles two rows
  size_t spx = px;
  if (fl) {
    const int y = (int)(o / W), x = (int)(o - (size_t)y * W);
    spx = b * hw + (size_t)y * W + (W - 4 - x);  // the four source pixels, read in reverse below
  }
  const uint32_t* q = reinterpret_cast<const uint32_t*>(in + spx * 3);
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];  // bytes r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
  uint8_t r[4] = {(uint8_t)w0, (uint8_t)(w0 >> 24), (uint8_t)(w1 >> 16), (uint8_t)(w2 >> 8)};
  uint8_t g[4] = {(uint8_t)(w0 >> 8), (uint8_t)w1, (uint8_t)(w1 >> 24), (uint8_t)(w2 >> 16)};
  uint8_t bl[4] = {(uint8_t)(w0 >> 16), (uint8_t)(w1 >> 8), (uint8_t)w2, (uint8_t)(w2 >> 24)};
  if (fl) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint8_t t;
      t = r[k]; r[k] = r[3 - k]; r[3 - k] = t;
      t = g[k]; g[k] = g[3 - k]; g[3 - k] = t;
      t = bl[k]; bl[k] = bl[3 - k]; bl[3 - k] = t;
    }
  }
  Out* dst = out + b * 3 * hw + o;
  store4(dst, r);
  store4(dst + hw, g);
  store4(dst + 2 * (size_t)hw, bl);
}

// level s-1 (HWC uint8, hi x wi) -> level s (HWC uint8, hi/2 x wi/2) + CHW tensor.  One CTA: 32 x 8 outputs.
// The input rows/columns the tile's taps touch are first copied to shared memory (32-bit loads when the rows
// are word-aligned), then both passes run out of shared memory with the coefficient row of the thread's
// output column / row held in registers and the 13 reserved taps fully unrolled (the table is zero beyond
// the taps in use).  Optionally the CTA also writes the CHW tensor of ITS part of the input level
// (out_in: the 64 x 16 input pixels under the tile), which saves the separate conversion launch.
// kTY output rows per tile: the horizontal pass covers the 2 kTY + 13 input rows the tile's vertical taps touch, so a
// taller tile repeats less of it (29 rows for 8 outputs, 45 for 16); the large level uses 16, the small ones 8 (more CTAs)
#ifndef VSL_LANCZOS_TALL
#define VSL_LANCZOS_TALL 1
#endif
constexpr int kTX = 32, kRowBytes = (2 * kTX + kKsize + 6) * 3 / 4 * 4 + 4;
template <int kTY>
struct LanczosSmem {
  static constexpr int kRowsMax = 2 * kTY + kKsize + 1;
  __align__(16) uint8_t tin[kRowsMax][kRowBytes];  // input window of the tile
  uint8_t hrow[kRowsMax][kTX][3];                  // its horizontal pass, rounded to 8 bits
  int32_t kx[kTX][kKsize], ky[kTY][kKsize];        // the tile's coefficient rows (13 words: conflict-free)
  int bx[kTX], by[kTY];                            // first tap of every output column / row
};
template <class Out, bool kWords, int kTY>
__device__ __forceinline__ void lanczos_tile(LanczosSmem<kTY>& sm, const uint8_t* __restrict__ in, uint8_t* __restrict__ out_u8,
                                             Out* __restrict__ out_t, Out* __restrict__ out_in, AxisTable tx, AxisTable ty,
                                             int hi, int wi, const uint8_t* __restrict__ flip, int b, int x0, int y0) {
  auto& tin = sm.tin; auto& hrow = sm.hrow; auto& kx = sm.kx; auto& ky = sm.ky; auto& bx = sm.bx; auto& by = sm.by;
  const int ho = hi >> 1, wo = wi >> 1;
  const int ylast = min(y0 + kTY, ho) - 1, xlast = min(x0 + kTX, wo) - 1;
  const int row_lo = ty.bounds[2 * y0];
  const int row_hi = ty.bounds[2 * ylast] + ty.bounds[2 * ylast + 1];  // exclusive
  const int nrows = row_hi - row_lo;                                    // <= 2 * 7 + 13
  // first input column of the window; rounded down to a multiple of 4 pixels (12 bytes) for the word loads
  const int col_lo = kWords ? (tx.bounds[2 * x0] & ~3) : tx.bounds[2 * x0];
  const int ncolb = (tx.bounds[2 * xlast] + tx.bounds[2 * xlast + 1] - col_lo) * 3;  // bytes per row, <= (3 + 2 * 31 + 13) * 3
  const size_t pitch = (size_t)wi * 3;
  const uint8_t* src = in + ((size_t)b * hi + row_lo) * pitch + (size_t)col_lo * 3;
  for (int item = threadIdx.x; item < kTX * kKsize; item += 256) {
    const int o = item / kKsize, j = item - o * kKsize;
    kx[o][j] = x0 + o < wo ? tx.coefs[(size_t)(x0 + o) * kKsize + j] : 0;
    if (o < kTY) ky[o][j] = y0 + o < ho ? ty.coefs[(size_t)(y0 + o) * kKsize + j] : 0;
  }
  if (threadIdx.x < kTX) {
    const int o = threadIdx.x;
    bx[o] = x0 + o < wo ? tx.bounds[2 * (x0 + o)] - col_lo : 0;
    if (o < kTY) by[o] = y0 + o < ho ? ty.bounds[2 * (y0 + o)] - row_lo : 0;
  }
  // Staging: every global load of a thread is issued before its first shared-memory store (a plain load -> store
  // loop serialises one memory round trip per row: eight in a row per tile, which is what the kernel used to cost)
  constexpr int kRowsMax = LanczosSmem<kTY>::kRowsMax;
  constexpr int kStageIt = (kRowsMax + 3) / 4;
  if (kWords && flip && flip[b]) {
    // mirrored read of the raw frame (level 1 only): window column c holds source column wi - 1 - (col_lo + c).
    // Word loads over the mirrored source span; each byte is stored at its pixel's mirrored place, channels in order.
    const int ncol = ncolb / 3;
    const int sb0 = (wi - col_lo - ncol) * 3, a0 = sb0 & ~3;     // source byte span [sb0, sb0 + ncolb), word-aligned start
    const int nw = (sb0 + ncolb - a0 + 3) >> 2;                    // <= 60 words; stays inside the row (wi * 3 % 4 == 0)
    const uint8_t* row0 = in + ((size_t)b * hi + row_lo) * pitch + a0;
    const int w = threadIdx.x & 63;
    uint32_t v[kStageIt];
#pragma unroll
    for (int k = 0; k < kStageIt; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      v[k] = (r < nrows && w < nw) ? reinterpret_cast<const uint32_t*>(row0 + r * pitch)[w] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kStageIt; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      if (r >= nrows || w >= nw) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int rel = a0 + 4 * w + q - sb0;
        if (rel < 0 || rel >= ncolb) continue;
        const int ps = rel / 3, ch = rel - 3 * ps;
        tin[r][(ncol - 1 - ps) * 3 + ch] = (uint8_t)(v[k] >> (8 * q));
      }
    }
  } else if (flip && flip[b]) {  // ragged widths: byte loads
    const uint8_t* row0 = in + ((size_t)b * hi + row_lo) * pitch;
    for (int idx = threadIdx.x; idx < nrows * ncolb; idx += 256) {
      const int r = idx / ncolb, j = idx - r * ncolb;
      const int c = j / 3, ch = j - 3 * c;
      tin[r][j] = row0[r * pitch + (size_t)(wi - 1 - (col_lo + c)) * 3 + ch];
    }
  } else if (kWords) {  // rows start on a word boundary (wi % 4 == 0; col_lo is a multiple of 4): <= 59 words per row
    const int nw = (ncolb + 3) >> 2;  // the last word may reach up to 3 bytes past the taps: still inside the row
    const int w = threadIdx.x & 63;
    uint32_t v[kStageIt];
#pragma unroll
    for (int k = 0; k < kStageIt; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      v[k] = (r < nrows && w < nw) ? reinterpret_cast<const uint32_t*>(src + r * pitch)[w] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kStageIt; ++k) {
      const int r = (threadIdx.x >> 6) + 4 * k;
      if (r < nrows && w < nw) reinterpret_cast<uint32_t*>(tin[r])[w] = v[k];
    }
  } else if ((int)threadIdx.x < ncolb) {  // one thread per byte column, independent loads down the rows
    const uint8_t* q = src + threadIdx.x;
#pragma unroll 5
    for (int r = 0; r < nrows; ++r) tin[r][threadIdx.x] = q[r * pitch];
  }
  __syncthreads();
  {
    const int xx = threadIdx.x & (kTX - 1);
    if (x0 + xx < wo) {
      int k[kKsize];
#pragma unroll
      for (int j = 0; j < kKsize; ++j) k[j] = kx[xx][j];
      const int lo3 = bx[xx] * 3;
      for (int r = threadIdx.x >> 5; r < nrows; r += 8) {
        const uint8_t* q = &tin[r][lo3];
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
        for (int j = 0; j < kKsize; ++j) {
          a0 += q[3 * j] * k[j]; a1 += q[3 * j + 1] * k[j]; a2 += q[3 * j + 2] * k[j];
        }
        hrow[r][xx][0] = clip8(a0); hrow[r][xx][1] = clip8(a1); hrow[r][xx][2] = clip8(a2);
      }
    }
  }
  if (out_in) {  // CHW tensor of the input pixels under this tile (2 kTY rows x 2 kTX columns)
    const size_t hwi = (size_t)hi * wi;
    for (int item = threadIdx.x; item < 2 * kTY * 2 * kTX; item += 256) {
      const int ry = item / (2 * kTX), rx = item - ry * (2 * kTX);
      const int y = 2 * y0 + ry, x = 2 * x0 + rx;
      if (y >= hi || x >= wi) continue;
      const uint8_t* q = &tin[y - row_lo][(x - col_lo) * 3];
      Out* dst = out_in + (size_t)b * 3 * hwi + (size_t)y * wi + x;
      store_tensor<Out>(dst, 0, q[0]);
      store_tensor<Out>(dst, hwi, q[1]);
      store_tensor<Out>(dst, 2 * hwi, q[2]);
    }
  }
  __syncthreads();
  const int xx = threadIdx.x & (kTX - 1), xo = x0 + xx;
  if (xo >= wo) return;
#pragma unroll
  for (int yy = threadIdx.x / kTX; yy < kTY; yy += 256 / kTX) {
    const int yo = y0 + yy;
    if (yo >= ho) break;
    const int lo = by[yy];
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
    for (int j = 0; j < kKsize; ++j) {
      const int w = ky[yy][j];
      a0 += hrow[lo + j][xx][0] * w; a1 += hrow[lo + j][xx][1] * w; a2 += hrow[lo + j][xx][2] * w;
    }
    const uint8_t v0 = clip8(a0), v1 = clip8(a1), v2 = clip8(a2);
    const size_t o = (size_t)yo * wo + xo, hw = (size_t)ho * wo;
    uint8_t* d8 = out_u8 + ((size_t)b * hw + o) * 3;
    d8[0] = v0; d8[1] = v1; d8[2] = v2;
    if (out_t) {
      Out* dt = out_t + (size_t)b * 3 * hw + o;
      store_tensor<Out>(dt, 0, v0);
      store_tensor<Out>(dt, hw, v1);
      store_tensor<Out>(dt, 2 * hw, v2);
    }
  }
}

// one 2:1 level per launch
template <class Out, bool kWords, int kTY>
__global__ void __launch_bounds__(256) k_lanczos_half(const uint8_t* __restrict__ in, uint8_t* __restrict__ out_u8,
                                                     Out* __restrict__ out_t, Out* __restrict__ out_in, AxisTable tx,
                                                     AxisTable ty, int hi, int wi, const uint8_t* __restrict__ flip) {
  __shared__ LanczosSmem<kTY> sm;
  lanczos_tile<Out, kWords, kTY>(sm, in, out_u8, out_t, out_in, tx, ty, hi, wi, flip, blockIdx.z, blockIdx.x * kTX,
                                 blockIdx.y * kTY);
}

template <class Out>
static int pyramid_forward_impl(const VslPyramidDesc* d, const PyramidPlan& pl, const uint8_t* frames,
                                const uint8_t* flip, void* const levels[VSL_MAX_SCALES], uint8_t* ws, cudaStream_t st) {
  const size_t px0 = (size_t)d->batch * d->height * d->width;
  // with more than one level the first 2:1 kernel also writes the level-0 tensor (its tiles cover level 0)
  const bool fuse0 = levels[0] && d->num_levels > 1;
  if (levels[0] && !fuse0) {
    const int hw = d->height * d->width;
    if (hw % 4 == 0 && (!flip || d->width % 4 == 0) && ((uintptr_t)frames & 3u) == 0 && ((uintptr_t)levels[0] & 15u) == 0)
      k_u8_to_tensor_x4<Out><<<(unsigned)((px0 / 4 + 255) / 256), 256, 0, st>>>(frames, (Out*)levels[0], hw, px0 / 4,
                                                                                 d->width, flip);
    else
      k_u8_to_tensor<Out><<<(unsigned)((px0 + 255) / 256), 256, 0, st>>>(frames, (Out*)levels[0], hw, px0, d->width, flip);
    VSL_CUDA_OK_IN(cudaGetLastError());
  }
  const uint8_t* prev = frames;
  for (int s = 1; s < d->num_levels; ++s) {
    const int hi = d->height >> (s - 1), wi = d->width >> (s - 1);
    AxisTable tx = {(const int32_t*)(ws + pl.off_xb[s]), (const int32_t*)(ws + pl.off_xc[s])};
    AxisTable ty = {(const int32_t*)(ws + pl.off_yb[s]), (const int32_t*)(ws + pl.off_yc[s])};
    uint8_t* cur = ws + pl.off_u8[s];
    Out* out_in = (s == 1 && fuse0) ? (Out*)levels[0] : nullptr;
    const uint8_t* fl = s == 1 ? flip : nullptr;  // later levels read the already mirrored 8-bit level
    const bool words = wi % 4 == 0 && ((uintptr_t)prev & 3u) == 0;
    // taller tiles where the level is large enough to fill the GPU with them (>= 4 CTAs per SM)
    const int ty16 = ((wi >> 1) + kTX - 1) / kTX * (((hi >> 1) + 15) / 16) * d->batch;
    if (VSL_LANCZOS_TALL && ty16 >= 4 * 148) {
      dim3 grid(((wi >> 1) + kTX - 1) / kTX, ((hi >> 1) + 15) / 16, d->batch);
      if (words) k_lanczos_half<Out, true, 16><<<grid, 256, 0, st>>>(prev, cur, (Out*)levels[s], out_in, tx, ty, hi, wi, fl);
      else k_lanczos_half<Out, false, 16><<<grid, 256, 0, st>>>(prev, cur, (Out*)levels[s], out_in, tx, ty, hi, wi, fl);
    } else {
      dim3 grid(((wi >> 1) + kTX - 1) / kTX, ((hi >> 1) + 7) / 8, d->batch);
      if (words) k_lanczos_half<Out, true, 8><<<grid, 256, 0, st>>>(prev, cur, (Out*)levels[s], out_in, tx, ty, hi, wi, fl);
      else k_lanczos_half<Out, false, 8><<<grid, 256, 0, st>>>(prev, cur, (Out*)levels[s], out_in, tx, ty, hi, wi, fl);
    }
    VSL_CUDA_OK_IN(cudaGetLastError());
    prev = cur;
  }
  return VSL_OK;
}

// ---- arbitrary-ratio resize: the decoded file image -> level 0 ---------------------------------------------
// `self.resize[0](inputs[(n, im, -1)])` (datasets/mono_dataset2.py:85-89, :107-109): PIL Image.resize with LANCZOS
// from the native resolution (e.g. 1242 x 375, 1280 x 1024) to (width, height).  Same two 8-bit passes as the pyramid
// levels (horizontal first, each rounded to 8 bits) with per-size tap counts; one thread per output pixel.
struct ResizePlan {
  int ksx, ksy;                       // taps reserved per output column / row
  size_t off_xb, off_xc, off_yb, off_yc, off_mid, total;
};
static bool resize_dims_ok(int batch, int in_h, int in_w, int out_h, int out_w) {
  return batch >= 1 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1 && in_h <= 16384 && in_w <= 16384 &&
         out_h <= 16384 && out_w <= 16384;
}
static ResizePlan make_resize_plan(int batch, int in_h, int in_w, int out_h, int out_w) {
  ResizePlan pl = {};
  pl.ksx = axis_ksize(in_w, out_w);
  pl.ksy = axis_ksize(in_h, out_h);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  pl.off_xb = take((size_t)out_w * 2 * 4);
  pl.off_xc = take((size_t)out_w * pl.ksx * 4);
  pl.off_yb = take((size_t)out_h * 2 * 4);
  pl.off_yc = take((size_t)out_h * pl.ksy * 4);
  pl.off_mid = take((size_t)batch * in_h * out_w * 3);   // the horizontal pass's output
  pl.total = off;
  return pl;
}
// one pass (ImagingResampleHorizontal_8bpc / Vertical_8bpc): out[b, y, x] from n taps along the axis
template <bool kHorizontal>
__global__ void __launch_bounds__(256) k_resample_axis(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                       const int32_t* __restrict__ bounds, const int32_t* __restrict__ coefs,
                                                       int ksize, int in_h, int in_w, int out_h, int out_w, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % out_w);
  const size_t t = i / out_w;
  const int y = (int)(t % out_h), b = (int)(t / out_h);
  const int o = kHorizontal ? x : y;
  const int lo = bounds[2 * o], n = bounds[2 * o + 1];
  const int32_t* k = coefs + (size_t)o * ksize;
  const uint8_t* src = kHorizontal ? in + (((size_t)b * in_h + y) * in_w + lo) * 3 : in + (((size_t)b * in_h + lo) * in_w + x) * 3;
  const size_t step = kHorizontal ? 3 : (size_t)in_w * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < n; ++j) {
    const int w = k[j];
    const uint8_t* q = src + j * step;
    a0 += q[0] * w; a1 += q[1] * w; a2 += q[2] * w;
  }
  uint8_t* d = out + i * 3;
  d[0] = clip8(a0); d[1] = clip8(a1); d[2] = clip8(a2);
}

}  // namespace vsl

using namespace vsl;

extern "C" {

size_t vsl_pyramid_workspace_bytes(const VslPyramidDesc* desc) {
  if (!pyr_desc_ok(desc)) return 0;
  const size_t t = make_pyr_plan(desc).total;
  return t ? t : 256;
}

int vsl_pyramid_plan(const VslPyramidDesc* d, void* workspace, size_t workspace_bytes, void* stream) {
  if (!pyr_desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (!workspace) return VSL_ERR_NULL_POINTER;
  if (((uintptr_t)workspace & 255u) != 0) return VSL_ERR_MISALIGNED;
  const PyramidPlan pl = make_pyr_plan(d);
  if (workspace_bytes < pl.total) return VSL_ERR_WORKSPACE;
  uint8_t* ws = (uint8_t*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int32_t> bounds, coefs;
  for (int s = 1; s < d->num_levels; ++s) {
    const int hi = d->height >> (s - 1), wi = d->width >> (s - 1);
    // pageable source: the runtime stages the bytes before cudaMemcpyAsync returns, so the vectors may be reused
    if (!axis_coeffs(wi, wi >> 1, bounds, coefs)) return VSL_ERR_UNSUPPORTED;
    VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_xb[s], bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
    VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_xc[s], coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, st));
    if (!axis_coeffs(hi, hi >> 1, bounds, coefs)) return VSL_ERR_UNSUPPORTED;
    VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_yb[s], bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice, st));
    VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_yc[s], coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, st));
  }
  return VSL_OK;
}

int vsl_pyramid_coefficients(int in_size, int out_size, int32_t* bounds, int32_t* coefs, int ksize_capacity) {
  if (in_size < 1 || out_size < 1 || !bounds || !coefs) return VSL_ERR_BAD_DESC;
  if (ksize_capacity != kKsize) return VSL_ERR_BAD_DESC;
  std::vector<int32_t> b, c;
  if (!axis_coeffs(in_size, out_size, b, c)) return VSL_ERR_UNSUPPORTED;
  for (size_t i = 0; i < b.size(); ++i) bounds[i] = b[i];
  for (size_t i = 0; i < c.size(); ++i) coefs[i] = c[i];
  return VSL_OK;
}

int vsl_pyramid_forward(const VslPyramidDesc* d, const uint8_t* frames_hwc, void* const levels[VSL_MAX_SCALES],
                        uint8_t* const levels_u8[VSL_MAX_SCALES], void* workspace, size_t workspace_bytes, void* stream) {
  return vsl_pyramid_forward_flip(d, frames_hwc, nullptr, levels, levels_u8, workspace, workspace_bytes, stream);
}

// stereo_T of MonoDataset.__getitem__ (datasets/mono_dataset2.py:197-203): identity with
// T[0,3] = side_sign * baseline_sign * baseline, side_sign = -1 for the left camera, baseline_sign = -1 when flipped
__global__ void k_stereo_T(int B, const uint8_t* __restrict__ flip, const uint8_t* __restrict__ side_left, float baseline,
                           float* __restrict__ T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 16) return;
  const int b = i >> 4, e = i & 15;
  float v = (e == 0 || e == 5 || e == 10 || e == 15) ? 1.0f : 0.0f;
  if (e == 3) {
    const bool neg = ((flip && flip[b]) ? 1 : 0) != ((side_left && side_left[b]) ? 1 : 0);
    v = neg ? -baseline : baseline;
  }
  T[i] = v;
}

int vsl_stereo_transform(int batch, const uint8_t* flip, const uint8_t* side_left, float baseline, float* T, void* stream) {
  if (batch < 1) return VSL_ERR_BAD_DESC;
  if (!T) return VSL_ERR_NULL_POINTER;
  k_stereo_T<<<(batch * 16 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(batch, flip, side_left, baseline, T);
  VSL_CUDA_OK_IN(cudaGetLastError());
  return VSL_OK;
}

int vsl_pyramid_forward_flip(const VslPyramidDesc* d, const uint8_t* frames_hwc, const uint8_t* flip,
                             void* const levels[VSL_MAX_SCALES], uint8_t* const levels_u8[VSL_MAX_SCALES], void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!pyr_desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (!frames_hwc || !levels || !workspace) return VSL_ERR_NULL_POINTER;
  if (((uintptr_t)workspace & 255u) != 0) return VSL_ERR_MISALIGNED;
  const PyramidPlan pl = make_pyr_plan(d);
  if (workspace_bytes < pl.total) return VSL_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = d->out_dtype == VSL_DTYPE_BF16
               ? pyramid_forward_impl<bf16_t>(d, pl, frames_hwc, flip, levels, (uint8_t*)workspace, st)
               : pyramid_forward_impl<float>(d, pl, frames_hwc, flip, levels, (uint8_t*)workspace, st);
  if (rc != VSL_OK) return rc;
  if (levels_u8) {  // optional copies of the 8-bit levels (what PIL would hold), for checking / logging
    for (int s = 1; s < d->num_levels; ++s)
      if (levels_u8[s])
        VSL_CUDA_OK_IN(cudaMemcpyAsync(levels_u8[s], (uint8_t*)workspace + pl.off_u8[s],
                                       (size_t)d->batch * (d->height >> s) * (d->width >> s) * 3,
                                       cudaMemcpyDeviceToDevice, st));
  }
  return VSL_OK;
}

size_t vsl_resize_workspace_bytes(int batch, int in_height, int in_width, int out_height, int out_width) {
  if (!resize_dims_ok(batch, in_height, in_width, out_height, out_width)) return 0;
  return make_resize_plan(batch, in_height, in_width, out_height, out_width).total;
}

int vsl_resize_plan(int batch, int in_height, int in_width, int out_height, int out_width, void* workspace,
                    size_t workspace_bytes, void* stream) {
  if (!resize_dims_ok(batch, in_height, in_width, out_height, out_width)) return VSL_ERR_BAD_DESC;
  if (!workspace) return VSL_ERR_NULL_POINTER;
  if (((uintptr_t)workspace & 255u) != 0) return VSL_ERR_MISALIGNED;
  const ResizePlan pl = make_resize_plan(batch, in_height, in_width, out_height, out_width);
  if (workspace_bytes < pl.total) return VSL_ERR_WORKSPACE;
  uint8_t* ws = (uint8_t*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<int32_t> b, c;
  axis_coeffs_stride(in_width, out_width, pl.ksx, b, c);
  VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_xb, b.data(), b.size() * 4, cudaMemcpyHostToDevice, st));
  VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_xc, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
  axis_coeffs_stride(in_height, out_height, pl.ksy, b, c);
  VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_yb, b.data(), b.size() * 4, cudaMemcpyHostToDevice, st));
  VSL_CUDA_OK_IN(cudaMemcpyAsync(ws + pl.off_yc, c.data(), c.size() * 4, cudaMemcpyHostToDevice, st));
  return VSL_OK;
}

int vsl_resize_forward(int batch, int in_height, int in_width, int out_height, int out_width, const uint8_t* frames_hwc,
                       uint8_t* out_hwc, void* workspace, size_t workspace_bytes, void* stream) {
  if (!resize_dims_ok(batch, in_height, in_width, out_height, out_width)) return VSL_ERR_BAD_DESC;
  if (!frames_hwc || !out_hwc || !workspace) return VSL_ERR_NULL_POINTER;
  if (((uintptr_t)workspace & 255u) != 0) return VSL_ERR_MISALIGNED;
  const ResizePlan pl = make_resize_plan(batch, in_height, in_width, out_height, out_width);
  if (workspace_bytes < pl.total) return VSL_ERR_WORKSPACE;
  uint8_t* ws = (uint8_t*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // Pillow: a pass whose size does not change is skipped altogether (no filtering, no rounding)
  const bool need_h = in_width != out_width, need_v = in_height != out_height;
  if (!need_h && !need_v) {
    VSL_CUDA_OK_IN(cudaMemcpyAsync(out_hwc, frames_hwc, (size_t)batch * in_height * in_width * 3, cudaMemcpyDeviceToDevice, st));
    return VSL_OK;
  }
  const uint8_t* cur = frames_hwc;
  if (need_h) {
    uint8_t* dst = need_v ? ws + pl.off_mid : out_hwc;
    const size_t total = (size_t)batch * in_height * out_width;
    k_resample_axis<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        cur, dst, (const int32_t*)(ws + pl.off_xb), (const int32_t*)(ws + pl.off_xc), pl.ksx, in_height, in_width, in_height,
        out_width, total);
    cur = dst;
  }
  if (need_v) {
    const size_t total = (size_t)batch * out_height * out_width;
    k_resample_axis<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        cur, out_hwc, (const int32_t*)(ws + pl.off_yb), (const int32_t*)(ws + pl.off_yc), pl.ksy, in_height, out_width,
        out_height, out_width, total);
  }
  VSL_CUDA_OK_IN(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
