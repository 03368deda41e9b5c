// Fused view-synthesis loss for sm_100a: kernels + C ABI (include/vsl.h).
//
// Launch sequence of vsl_loss_forward_backward (one stream, no host sync):
//   k_photometric   warp + SSIM/L1 + automask min + adjoint       (trainer.py:491-674)  <- the hot kernel
//                   + the edge-aware smoothness terms of every level on the tile's own pixels (layers.py:286-299)
//   k_epilogue      sums the tiles' up-sample-adjoint partials (s >= 1), deterministic reductions (losses, dL/dP,
//                   per-image mean of disp_s and smoothness sums, trainer.py:676-680), loss dict
// and of vsl_loss_combine_grads (the backward):
//   k_combine       smoothness chain rule through the per-image mean + upstream weights
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/vsl.h"
#include "vsl_tile.cuh"

// Three source frames run as one 512-thread CTA per SM (the tile needs > 114 KB).  With the whole SM's shared memory
// to itself the tile can be 32 x 24 (201 KB): halo recomputation 1.31x / 1.15x instead of 1.41x / 1.20x, and 1,008
// region pixels / 884 windows fill two passes of 512 threads (98 % / 86 %) where 720 / 612 filled 70 % / 60 %.
#ifndef VSL_F3_TILE_H
#define VSL_F3_TILE_H 24
#endif
#ifndef VSL_ADJ_PAIR_F3
#define VSL_ADJ_PAIR_F3 0   // paired adjoint for three frames: 54 accumulators next to 36 dL/dP sums (measured)
#endif

namespace vsl {

thread_local int g_last_cuda_error = 0;  // shared with vsl_layers.cu

#define VSL_CUDA_OK(expr)                         \
  do {                                            \
    cudaError_t e__ = (expr);                     \
    if (e__ != cudaSuccess) {                     \
      g_last_cuda_error = (int)e__;               \
      return VSL_ERR_CUDA;                        \
    }                                             \
  } while (0)

constexpr int kChunk = 1024;   // native-resolution pixels per block in the small kernels
constexpr int kSmallNT = 256;

struct SmallParams {  // smoothness + epilogue
  const float* disp[kMaxScales];
  const void* img[kMaxScales];    // target pyramid (fp32 or bf16)
  float* gsmooth[kMaxScales];     // d smooth_s / d disp_s
  float* gphoto[kMaxScales];      // d min_loss_s / d disp_s
  const float* gpart[kMaxScales]; // per-CTA partial up-sample adjoints (null for identity scales)
  float* norm;                    // [S][B][2]: 1/(mean disp + 1e-7), sum(g*d) * inv^2 / n   (for k_combine)
  int tw, th, tiles_x, tiles_y;
  int log_tw, log_th, level_shift[kMaxScales];
  const float* partials;          // photometric partials [numCTA][S][kPartial]
  float* lossb;                   // [S][B]
  float* smoothb;                 // [S][B][2]
  float* gradP;                   // [S][F][B][12]
  float* losses;                  // [3S+1]: min_loss (S), loss/s (S), loss, smooth (S)
  unsigned* counter;
  int B, H, W, S, F;
  int hs[kMaxScales], ws[kMaxScales], scale_id[kMaxScales], identity_scale[kMaxScales];
  float scale_h[kMaxScales], scale_w[kMaxScales];
  int chunks0, tiles_per_image, kpartial;
  unsigned epilogue_blocks;       // number of k_epilogue blocks that have work (the others exit at once)
  float smooth_weight;
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still draining; it must call pdl_wait() before it touches anything the
// predecessor wrote.  The predecessor calls pdl_launch_dependents() once its CTAs no longer need to be alone.
#ifndef VSL_PDL
#define VSL_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if VSL_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if VSL_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <class... Args>
static cudaError_t launch_after(bool programmatic, void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (programmatic && VSL_PDL) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum 32 per-lane values across the warp at once: each butterfly step halves the number of values a lane
// holds while adding its partner's copy of the half it keeps (31 shuffles instead of 32 x 5).  On return
// lane l holds the warp total of v[l].  Fixed order, so the result is reproducible.
template <int N>
__device__ __forceinline__ void butterfly_step(float (&v)[32], int lane) {
  constexpr int H = N / 2;
  const bool up = (lane & H) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float keep = up ? v[i + H] : v[i];
    const float send = up ? v[i] : v[i + H];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
  }
}
__device__ __forceinline__ float warp_sum32(float (&v)[32], int lane) {
  butterfly_step<32>(v, lane);
  butterfly_step<16>(v, lane);
  butterfly_step<8>(v, lane);
  butterfly_step<4>(v, lane);
  butterfly_step<2>(v, lane);
  return v[0];
}

// ---------------------------------------------------------------------------------------------
// kFastArith: the rounding-order selectors are the compile-time default (arith == 0, PyTorch-CUDA order
// for batch >= 2), so every variant branch in vsl_math.cuh folds away.
template <class C, bool kFastArith>
__global__ void __launch_bounds__(C::NT, C::NT >= 512 ? 1 : 2) k_photometric(const PhotoParams p) {
  extern __shared__ __align__(16) float sm[];
  GeoConst g = p.g;
  if (kFastArith) g.arith = 0;
  TileCtx t;
  {
    // read once and kept: left to itself the compiler re-reads the special registers inside the phase loops
    unsigned bx, by, bz;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
    asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
    asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
    t.b = (int)bz;
    t.x0 = (int)bx * C::TW;
    t.y0 = (int)by * C::TH;
    t.cta = (int)((bz * gridDim.y + by) * gridDim.x + bx);
  }
  const int tid = threadIdx.x;
  // Global reads whose addresses do not depend on computed values are issued as asynchronous copies well
  // before their phase: disp_s window (one scale ahead), tie-break noise (a phase ahead), image tiles.
  phase_stage_disp<C>(p, t, sm, 0, tid);
  if constexpr (sizeof(typename C::Img) == 4) {
    phase_stage_images<C>(p, t, sm, tid, p.automask != 0);
    stage_wait<0>();
    __syncthreads();
    phase_target_stats<C>(p, t, sm, tid);
    if (p.automask) phase_identity<C>(p, g, t, sm, tid);
  } else {
    phase_load_tiles<C>(p, t, sm, tid, p.automask != 0);
    __syncthreads();
    phase_target_stats<C>(p, t, sm, tid);
    if (p.automask) phase_identity<C>(p, g, t, sm, tid);
  }

  float* red = sm + C::oRed;
  for (int s = 0; s < p.S; ++s) {
    ThreadState<C> ts;
    ts.loss = 0.f;
#pragma unroll
    for (int k = 0; k < C::F * 12; ++k) ts.dP[k] = 0.f;
    // P of the previous scale is no longer read (sync after its adjoint); with one pose for all scales it is formed once
    if (s == 0 || p.pose_per_scale) phase_pose<C>(p, g, t, sm, s, tid);
    phase_stage_noise<C>(p, t, sm, s, tid);  // lands during the warp phase
    stage_wait<1>();                         // this scale's disp window (committed one scale earlier)
    __syncthreads();
    phase_warp<C>(p, g, t, sm, s, tid);
    {
      // smoothness of this level on the tile's own level pixels (reads the staged disp window before it is
      // replaced); per-warp sums parked in shared memory until the scale's block reduction
      float sacc[4] = {0.f, 0.f, 0.f, 0.f};
      phase_smooth<C>(p, t, sm, s, tid, sacc);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float tot = warp_sum(sacc[k]);
        if ((tid & 31) == 0) sm[C::oRedS + (tid >> 5) * 4 + k] = tot;
      }
    }
    stage_wait<0>();
    __syncthreads();
    if (s + 1 < p.S) phase_stage_disp<C>(p, t, sm, s + 1, tid);  // the adjoint reads the saved depth, not disp
    float pre[3];
    if constexpr (!C::kStageImg) {
      if (s + 1 < p.S) phase_prefetch_level<C>(p, t, s + 1, tid, pre);  // lands during the window phase
    }
    if constexpr (C::AVG) phase_windows_avg<C>(p, g, t, sm, s, tid, ts);
    else if constexpr (C::F == 2) phase_windows_paired<C>(p, g, t, sm, s, tid, ts);
    else phase_windows<C>(p, g, t, sm, s, tid, ts);
    if constexpr (!C::kStageImg) {
      if (s + 1 < p.S) phase_store_level<C>(sm, tid, pre);  // phase_smooth of this scale is done with the buffer (barrier above)
    }
    __syncthreads();
    if (!p.forward_only) {
      if constexpr (VSL_ADJ_PAIR && !C::AVG && C::TH % 2 == 0 && C::IN > C::NT && (C::F < 3 || VSL_ADJ_PAIR_F3)) phase_backward_pair<C>(p, g, t, sm, s, tid, ts);
      else phase_backward<C>(p, g, t, sm, s, tid, ts);
    }
    // deterministic block reduction of (loss, dP) -> one partial per CTA and scale
    const int w = tid >> 5, l = tid & 31;
#pragma unroll
    for (int k0 = 0; k0 < C::kPartial; k0 += 32) {  // (loss, dP[...]) in batches of 32 values
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int k = k0 + i;
        v[i] = k == 0 ? ts.loss : (k < C::kPartial ? ts.dP[k - 1 < C::F * 12 ? k - 1 : 0] : 0.f);
      }
      const float tot = warp_sum32(v, l);
      if (k0 + l < C::kPartial) red[w * C::kPartial + k0 + l] = tot;
    }
    __syncthreads();
    if (tid < C::kPartialAll) {
      float r = 0.f;
#pragma unroll
      for (int i = 0; i < C::NT / 32; ++i)
        r += tid < C::kPartial ? red[i * C::kPartial + tid] : sm[C::oRedS + i * 4 + (tid - C::kPartial)];
      p.partials[((size_t)t.cta * p.S + s) * C::kPartialAll + tid] = r;
    }
    if (!p.identity_scale[s] && !p.forward_only) {  // block-uniform; the tile's d/d(up-sampled disp) is complete (sync above)
      phase_adjoint_rows<C>(p, t, sm, s, tid);
      __syncthreads();
      phase_adjoint_cols<C>(p, t, sm, s, tid);
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSmallNT) k_epilogue(const SmallParams p) {
  __shared__ bool is_last;
  pdl_wait();  // everything k_photometric wrote is visible from here on
  pdl_launch_dependents();
  int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  int h = p.hs[s], w = p.ws[s], n = h * w;
  int nchunk = (n + kChunk - 1) / kChunk;
  // identity levels only have the per-image reductions of chunk 0; idle blocks must not touch the counter
  if (chunk >= (p.identity_scale[s] ? 1 : nchunk)) return;
  if (!p.identity_scale[s] && p.gphoto[s]) {  // gphoto is null with VSL_FLAG_FORWARD_ONLY
    // d(min_loss/s)/d disp_s: add the (<= 4) tile partials of every coarse pixel, tiles in a fixed order
    float* gp = p.gphoto[s] + (size_t)b * n;
    const int lcw = p.log_tw - p.level_shift[s], ch = p.th >> p.level_shift[s];
    const int lch = p.log_th >= 0 ? p.log_th - p.level_shift[s] : -1;
    // one division per thread instead of two per pixel: the pixels of a thread are kSmallNT apart
    int i = chunk * kChunk + threadIdx.x;
    int jy = i / w, jx = i - jy * w;
    const int dy = kSmallNT / w, dx = kSmallNT - dy * w;
    for (; i < min(n, (chunk + 1) * kChunk); i += kSmallNT) {
      gp[i] = gather_adjoint_partials(p.gpart[s], b, jy, jx, lcw, ch, lch, p.tiles_x, p.tiles_y);
      jy += dy; jx += dx;
      if (jx >= w) { jx -= w; ++jy; }
    }
  }
  if (chunk == 0) {
    // per (scale, image): the tiles' partials (loss sum, dL/dP, smoothness sums) in fp64, tile order fixed:
    // 8 lane-groups each add every 8th tile (independent loads, so they pipeline), then the 8 group sums
    __shared__ double gsum[kSmallNT / 32][40];
    __shared__ double ssum[4];
    const float* base = p.partials + ((size_t)b * p.tiles_per_image * p.S + s) * p.kpartial;
    const int grp = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int kphoto = p.kpartial - 4;  // 1 + 12 F photometric values, then sum d, sum |dx| e, sum |dy| e, sum g d
    for (int k0 = 0; k0 < p.kpartial; k0 += 32) {  // kpartial <= 41
      double acc = 0.0;
      if (k0 + k < p.kpartial) {
        // same order of additions as a plain loop over this group's tiles; the loads of eight tiles are issued
        // together (left to itself the loop waited out one L2 round trip per tile: 30 in a row at config 1)
        constexpr int kU = 8, kG = kSmallNT / 32;
        for (int tl = grp; tl < p.tiles_per_image; tl += kG * kU) {
          float v[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const int ti = tl + kG * u;
            v[u] = ti < p.tiles_per_image ? __ldcg(base + (size_t)ti * p.S * p.kpartial + k0 + k) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) acc += (double)v[u];
        }
      }
      __syncthreads();
      gsum[grp][k] = acc;
      __syncthreads();
      if (grp == 0 && k0 + k < p.kpartial) {
        double tot = 0.0;
#pragma unroll
        for (int gi = 0; gi < kSmallNT / 32; ++gi) tot += gsum[gi][k];
        const int kk = k0 + k;
        if (kk == 0) p.lossb[s * p.B + b] = (float)tot;
        else if (kk < kphoto) {
          int f = (kk - 1) / 12, e = (kk - 1) % 12;
          if (p.gradP) p.gradP[((size_t)(s * p.F + f) * p.B + b) * 12 + e] = (float)tot;
        } else {
          ssum[kk - kphoto] = tot;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // norm = disp * inv, inv = 1 / (mean + 1e-7) (trainer.py:676-677); the tiles summed |d_a - d_b| e and
      // g for inv > 0, so |inv| scales the sums and sgn(inv) the gradient (k_combine: g * norm[0] - norm[1])
      const float mean = (float)(ssum[0] / (double)n);
      const float inv = 1.0f / (mean + 1e-7f);
      const float ainv = fabsf(inv);
      if (p.norm) {
        p.norm[(s * p.B + b) * 2] = ainv;
        p.norm[(s * p.B + b) * 2 + 1] = (inv < 0.f ? -1.f : 1.f) * (float)ssum[3] * inv * inv / (float)n;
      }
      p.smoothb[(s * p.B + b) * 2] = (float)(ssum[1] * (double)ainv);
      p.smoothb[(s * p.B + b) * 2 + 1] = (float)(ssum[2] * (double)ainv);
    }
  }
  // last block done: assemble the loss dict (trainer.py:672-685)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    is_last = atomicAdd(p.counter, 1u) == p.epilogue_blocks - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // all (scale, image) sums are loaded in parallel (L2, bypassing L1), then added per scale in image order
  __shared__ float fin[3][kMaxScales * 64];
  const int nsb = p.S * p.B;
  for (int i = threadIdx.x; i < nsb; i += kSmallNT) {
    if (i < kMaxScales * 64) {
      fin[0][i] = __ldcg(p.lossb + i);
      fin[1][i] = __ldcg(p.smoothb + 2 * i);
      fin[2][i] = __ldcg(p.smoothb + 2 * i + 1);
    }
  }
  __syncthreads();
  __shared__ double level_loss[kMaxScales];
  if (threadIdx.x < p.S) {
    const int si = threadIdx.x;
    double ml = 0.0, sx = 0.0, sy = 0.0;
    for (int bi = 0; bi < p.B; ++bi) {
      const int i = si * p.B + bi;
      if (i < kMaxScales * 64) { ml += (double)fin[0][i]; sx += (double)fin[1][i]; sy += (double)fin[2][i]; }
      else { ml += (double)__ldcg(p.lossb + i); sx += (double)__ldcg(p.smoothb + 2 * i); sy += (double)__ldcg(p.smoothb + 2 * i + 1); }
    }
    int hh = p.hs[si], ww = p.ws[si];
    double min_loss = ml / ((double)p.B * p.H * p.W);
    double smooth = sx / ((double)p.B * hh * (ww - 1)) + sy / ((double)p.B * (hh - 1) * ww);
    double loss = min_loss + (double)p.smooth_weight * smooth / (double)(1 << p.scale_id[si]);
    p.losses[si] = (float)min_loss;
    p.losses[p.S + si] = (float)loss;
    p.losses[2 * p.S + 1 + si] = (float)smooth;
    level_loss[si] = loss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int si = 0; si < p.S; ++si) total += level_loss[si];
    p.losses[2 * p.S] = (float)(total / p.S);
    *p.counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
struct CombineParams {
  const float* up;  // [2S+1] device
  const float* gphoto[kMaxScales];
  const float* gsmooth[kMaxScales];
  float* out[kMaxScales];
  const float* gradP;  // [S][F][B][12]
  float* gradP_out;    // [F][B][12] or null
  float* gradT_out;    // [F][B][16] ([S][F][B][16] with pose_per_scale) or null: K[:3,:]^T @ dP
  int pose_per_scale;
  const float* K;      // [B,4,4]
  const float* norm;   // [S][B][2]
  int B, S, F, chunks0;
  int n[kMaxScales], scale_id[kMaxScales], vec4[kMaxScales];
  float smooth_weight;
};
#ifndef VSL_COMBINE_VEC
#define VSL_COMBINE_VEC 1   // measured at C1 (graph-replayed step): 1 -> 0.7968 ms, 4 -> 0.7989 ms, 8 -> 0.8071 ms
#endif
constexpr int kCombineVec = VSL_COMBINE_VEC;            // float4 groups per thread, all loaded before the first is used
constexpr int kCombineChunk = kChunk * kCombineVec;     // elements per k_combine block

__global__ void __launch_bounds__(kSmallNT) k_combine(const CombineParams p) {
  pdl_wait();  // launched programmatically behind k_epilogue when the two are adjacent in the stream (graph replay)
  int s = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  float tot = p.up[2 * p.S] / (float)p.S;
  if (chunk * kCombineChunk < p.n[s]) {
    float a = p.up[s] + p.up[p.S + s] + tot;
    float bb = (p.up[p.S + s] + tot) * p.smooth_weight / (float)(1 << p.scale_id[s]);
    const float* gp = p.gphoto[s] + (size_t)b * p.n[s];
    const float* gs = p.gsmooth[s] + (size_t)b * p.n[s];
    float* o = p.out[s] + (size_t)b * p.n[s];
    // d smooth_s/d disp = g * inv - sum(g d) inv^2 / n  (chain through norm_disp = disp / (mean + 1e-7), trainer.py:676-677)
    const float inv = p.norm[(s * p.B + b) * 2], corr = p.norm[(s * p.B + b) * 2 + 1];
    const int i0 = chunk * kCombineChunk, i1 = min(p.n[s], i0 + kCombineChunk);
    if (p.vec4[s]) {  // level size and the three buffers 16-byte aligned: four elements per access
      float4 g1[kCombineVec], g2[kCombineVec];
#pragma unroll
      for (int k = 0; k < kCombineVec; ++k) {
        const int i = i0 + 4 * (threadIdx.x + k * kSmallNT);
        if (i < i1) { g1[k] = *reinterpret_cast<const float4*>(gp + i); g2[k] = *reinterpret_cast<const float4*>(gs + i); }
      }
#pragma unroll
      for (int k = 0; k < kCombineVec; ++k) {
        const int i = i0 + 4 * (threadIdx.x + k * kSmallNT);
        if (i < i1)
          *reinterpret_cast<float4*>(o + i) =
              make_float4(a * g1[k].x + bb * (g2[k].x * inv - corr), a * g1[k].y + bb * (g2[k].y * inv - corr),
                          a * g1[k].z + bb * (g2[k].z * inv - corr), a * g1[k].w + bb * (g2[k].w * inv - corr));
      }
    } else {
      for (int i = i0 + threadIdx.x; i < i1; i += kSmallNT) o[i] = a * gp[i] + bb * (gs[i] * inv - corr);
    }
  }
  if (chunk == 0 && (s == 0 || p.pose_per_scale) && (p.gradP_out || p.gradT_out)) {
    // shared pose: sum the scales' dL/dP with their upstream weights; posecnn: one pose per scale, no sum
    __shared__ float gP[kMaxSrc * 12];
    for (int k = threadIdx.x; k < p.F * 12; k += kSmallNT) {
      int f = k / 12, e = k % 12;
      float acc = 0.f;
      for (int si = (p.pose_per_scale ? s : 0); si < (p.pose_per_scale ? s + 1 : p.S); ++si) {
        float a = p.up[si] + p.up[p.S + si] + tot;
        acc += a * p.gradP[((size_t)(si * p.F + f) * p.B + b) * 12 + e];
      }
      gP[k] = acc;
      if (p.gradP_out && !p.pose_per_scale) p.gradP_out[((size_t)f * p.B + b) * 12 + e] = acc;
    }
    __syncthreads();
    if (p.gradT_out) {  // P = K[:3,:] @ T  ->  dL/dT = K[:3,:]^T @ dL/dP
      const size_t base = p.pose_per_scale ? (size_t)s * p.F * p.B * 16 : 0;
      for (int k = threadIdx.x; k < p.F * 16; k += kSmallNT) {
        int f = k / 16, e = k % 16, r = e >> 2, n = e & 3;
        const float* Kb = p.K + b * 16;
        p.gradT_out[base + ((size_t)f * p.B + b) * 16 + e] =
            Kb[r] * gP[f * 12 + n] + Kb[4 + r] * gP[f * 12 + 4 + n] + Kb[8 + r] * gP[f * 12 + 8 + n];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// side outputs of generate_images_pred for one scale (trainer.py:500-537)
struct WarpParams {
  const float* disp; const float* invK; const float* P[kMaxSrc]; const void* src[kMaxSrc];
  float* depth; float* sample[kMaxSrc]; float* color[kMaxSrc];
  int B, H, W, F, hs, ws, identity;
  float scale_h, scale_w;
  GeoConst g;
};

template <class Img>
__global__ void __launch_bounds__(256) k_warp_forward(const WarpParams p) {
  int HW = p.H * p.W;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)p.B * HW) return;
  int b = (int)(gid / HW), i = (int)(gid - (size_t)b * HW);
  int v = i / p.W, u = i - v * p.W;
  float D = upsample_disp(p.disp + (size_t)b * p.hs * p.ws, p.hs, p.ws, p.scale_h, p.scale_w, p.identity != 0, v, u,
                          p.g.arith);
  Cam cam = backproject_pixel(D, p.invK + b * 16, u, v, p.g);
  if (p.depth) p.depth[gid] = cam.z;
  for (int f = 0; f < p.F; ++f) {
    Proj pr = project_pixel(cam, p.P[f] + b * 12, p.g);
    if (p.sample[f]) {
      p.sample[f][gid * 2] = pr.gx;
      p.sample[f][gid * 2 + 1] = pr.gy;
    }
    if (p.color[f]) {
      Taps tp = bilinear_taps(pr, p.W, p.H);
      const Img* img = (const Img*)p.src[f] + (size_t)b * 3 * HW + pr.y0 * p.W + pr.x0;
      int dx = tp.x1ok ? 1 : 0, dy = tp.y1ok ? p.W : 0;
      for (int c = 0; c < 3; ++c) {
        const Img* q = img + c * HW;
        p.color[f][(size_t)b * 3 * HW + c * HW + i] =
            bilinear_value(tp, ldimg(q, 0), ldimg(q, dx), ldimg(q, dy), ldimg(q, dy + dx), p.g.arith);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
static bool desc_ok(const VslDesc* d) {
  if (!d || d->abi_version != VSL_ABI_VERSION) return false;
  if (d->batch < 1 || d->height < 2 || d->width < 2) return false;
  if (d->num_scales < 1 || d->num_scales > VSL_MAX_SCALES) return false;
  if (d->num_src < 1 || d->num_src > VSL_MAX_SRC) return false;
  for (int s = 0; s < d->num_scales; ++s) {
    int e = d->scale_ids[s];
    if (e < 0 || e > 3) return false;  // up-sample adjoint uses 2*2^e <= 16 lanes per coarse pixel
    if (e + d->smooth_level_bias < 0 || e + d->smooth_level_bias > 16) return false;
    if ((d->height >> e) < 2 || (d->width >> e) < 2) return false;
    if (((d->height >> e) << e) != d->height || ((d->width >> e) << e) != d->width) return false;
  }
  return true;
}

static GeoConst make_geo(const VslDesc* d) {
  GeoConst g;
  g.min_disp = d->min_disp; g.disp_range = d->disp_range; g.eps = d->eps; g.one = 1.0f;
  g.W = d->width; g.H = d->height;
  g.wm1 = (float)(d->width - 1); g.hm1 = (float)(d->height - 1);
  g.inv_wm1 = 1.0f / g.wm1; g.inv_hm1 = 1.0f / g.hm1;
  g.arith = d->arith;
  return g;
}

struct Plan {  // sizes derived from the descriptor; identical in workspace_bytes() and the launcher
  int tw, th, tiles_x, tiles_y, num_cta, kpartial, chunks0;
  size_t off_partials, off_gpart[kMaxScales], off_lossb, off_smoothb, off_counter, total;
};

// Tile height of the kernel variant that will run.  Two source frames: 32x16 tiles, 256 threads, two CTAs per
// SM.  Three: the 32x16 tile needs 143 KB, so either one 512-thread CTA per SM on 32x16 (the default fp32 /
// bf16 kernels: 11 % faster than 32x8, whose halo overhead is 1.69x / 1.33x instead of 1.41x / 1.20x) or, for
// the rarely used --avg_reprojection / --predictive_mask kernels, 32x8 tiles with 256 threads.
static int tile_height(const VslDesc* d, bool avg, bool pmask) {
  if (d->num_src >= 3) return (avg || pmask) ? 8 : VSL_F3_TILE_H;
  return 16;
}

static Plan make_plan(const VslDesc* d, int th) {
  Plan pl;
  pl.tw = 32;
  pl.th = th;
  pl.tiles_x = (d->width + pl.tw - 1) / pl.tw;
  pl.tiles_y = (d->height + pl.th - 1) / pl.th;
  pl.num_cta = pl.tiles_x * pl.tiles_y * d->batch;
  pl.kpartial = 1 + d->num_src * 12 + 4;  // TileCfg::kPartialAll
  pl.chunks0 = (d->height * d->width + kChunk - 1) / kChunk;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off += (floats + 63) / 64 * 64; return o; };
  pl.off_partials = take((size_t)pl.num_cta * d->num_scales * pl.kpartial);
  for (int s = 0; s < d->num_scales; ++s) {
    int r = 1 << d->scale_ids[s];
    pl.off_gpart[s] = r == 1 ? 0 : take((size_t)pl.num_cta * (pl.tw / r + 2) * (pl.th / r + 2));
  }
  pl.off_lossb = take((size_t)d->num_scales * d->batch);
  pl.off_smoothb = take((size_t)d->num_scales * d->batch * 2);
  pl.off_counter = take(64);
  pl.total = off * sizeof(float);
  return pl;
}

template <class C, bool kFastArith>
static int launch_photometric_impl(const PhotoParams& pp, const Plan& pl, int batch, cudaStream_t st) {
  // the attribute is per device: remember which devices have it (idempotent; a race only repeats the call)
  static bool attr_done[64] = {};
  int dev = 0;
  VSL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    VSL_CUDA_OK(cudaFuncSetAttribute(k_photometric<C, kFastArith>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::kBytes));
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid(pl.tiles_x, pl.tiles_y, batch);
  k_photometric<C, kFastArith><<<grid, C::NT, C::kBytes, st>>>(pp);
  VSL_CUDA_OK(cudaGetLastError());
  return VSL_OK;
}
template <class C>
static int launch_photometric(const PhotoParams& pp, const Plan& pl, int batch, cudaStream_t st) {
  return pp.g.arith == 0 ? launch_photometric_impl<C, true>(pp, pl, batch, st)
                         : launch_photometric_impl<C, false>(pp, pl, batch, st);
}

}  // namespace vsl

using namespace vsl;

extern "C" {

int vsl_abi_version(void) { return VSL_ABI_VERSION; }

const char* vsl_status_string(int status) {
  switch (status) {
    case VSL_OK: return "ok";
    case VSL_ERR_BAD_DESC: return "bad descriptor";
    case VSL_ERR_NULL_POINTER: return "null pointer";
    case VSL_ERR_MISALIGNED: return "misaligned buffer";
    case VSL_ERR_UNSUPPORTED: return "unsupported option";
    case VSL_ERR_WORKSPACE: return "workspace too small";
    case VSL_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

int vsl_last_cuda_error(void) { return g_last_cuda_error; }

size_t vsl_loss_workspace_bytes(const VslDesc* desc) {
  if (!desc_ok(desc)) return 0;
  // the variant (hence the tile height) depends on buffers the caller passes later: size for either
  size_t best = 0;
  for (int th : {8, 16, VSL_F3_TILE_H}) {
    const size_t t = make_plan(desc, th).total;
    if (t > best) best = t;
  }
  return best;
}

int vsl_loss_workspace_init(const VslDesc* d, void* workspace, size_t workspace_bytes, void* stream) {
  if (!desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (!workspace) return VSL_ERR_NULL_POINTER;
  if (workspace_bytes < vsl_loss_workspace_bytes(d)) return VSL_ERR_WORKSPACE;
  VSL_CUDA_OK(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));
  return VSL_OK;
}

int vsl_loss_forward_backward(const VslDesc* d, const VslLossBuffers* buf, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return vsl_loss_forward_backward_timed(d, buf, workspace, workspace_bytes, stream, nullptr, nullptr);
}

int vsl_event_create(void** event) {
  if (!event) return VSL_ERR_NULL_POINTER;
  cudaEvent_t e;
  VSL_CUDA_OK(cudaEventCreate(&e));
  *event = (void*)e;
  return VSL_OK;
}
int vsl_event_destroy(void* event) {
  if (!event) return VSL_ERR_NULL_POINTER;
  VSL_CUDA_OK(cudaEventDestroy((cudaEvent_t)event));
  return VSL_OK;
}
int vsl_event_elapsed_ms(void* start, void* stop, float* ms) {
  if (!start || !stop || !ms) return VSL_ERR_NULL_POINTER;
  VSL_CUDA_OK(cudaEventSynchronize((cudaEvent_t)stop));
  VSL_CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return VSL_OK;
}

int vsl_loss_forward_backward_timed(const VslDesc* d, const VslLossBuffers* buf, void* workspace,
                                    size_t workspace_bytes, void* stream, void* event_before, void* event_after) {
  if (!desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (!buf || !workspace) return VSL_ERR_NULL_POINTER;
  if (d->flags & ~(VSL_FLAG_AUTOMASK | VSL_FLAG_NO_SSIM | VSL_FLAG_AVG_REPROJECTION | VSL_FLAG_FORWARD_ONLY))
    return VSL_ERR_UNSUPPORTED;
  const bool fwd_only = (d->flags & VSL_FLAG_FORWARD_ONLY) != 0;
  const bool avg = (d->flags & VSL_FLAG_AVG_REPROJECTION) && d->num_src > 1;  // the mean of one frame is the frame
  if (d->image_dtype != VSL_DTYPE_F32 && d->image_dtype != VSL_DTYPE_BF16) return VSL_ERR_UNSUPPORTED;
  if (d->num_src > 3) return VSL_ERR_UNSUPPORTED;
  const bool automask = (d->flags & VSL_FLAG_AUTOMASK) != 0;
  const int S = d->num_scales, F = d->num_src;
  bool pmask = false;
  for (int s = 0; s < d->num_scales; ++s) pmask |= !automask && buf && buf->predictive_mask[s] != nullptr;
  Plan pl = make_plan(d, tile_height(d, avg, pmask));
  if (workspace_bytes < vsl_loss_workspace_bytes(d)) return VSL_ERR_WORKSPACE;
  if (((uintptr_t)workspace & 15u) != 0) return VSL_ERR_MISALIGNED;
  if (!buf->inv_K || !buf->losses) return VSL_ERR_NULL_POINTER;
  if (!fwd_only && (!buf->grad_P || !buf->smooth_norm)) return VSL_ERR_NULL_POINTER;
  for (int s = 0; s < S; ++s) {
    if (!buf->target[s] || !buf->disp[s] || (automask && !buf->noise[s])) return VSL_ERR_NULL_POINTER;
    if (!fwd_only && (!buf->grad_disp_photo[s] || !buf->grad_disp_smooth[s])) return VSL_ERR_NULL_POINTER;
  }
  for (int f = 0; f < F; ++f) {
    if (!buf->source[f]) return VSL_ERR_NULL_POINTER;
    for (int s = 0; s < S; ++s) {
      const float* T = buf->T_scale[s][f] ? buf->T_scale[s][f] : buf->T[f];
      if (!buf->P[f] && !T) return VSL_ERR_NULL_POINTER;
      if (T && !buf->K) return VSL_ERR_NULL_POINTER;
    }
  }
  if (d->scale_ids[0] != 0) return VSL_ERR_UNSUPPORTED;  // level 0 is the photometric target

  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;

  SmallParams sp = {};
  PhotoParams pp = {};
  pp.tgt = buf->target[0];
  pp.automask = automask ? 1 : 0;
  pp.pose_per_scale = 0;
  for (int s = 0; s < S; ++s)
    for (int f = 0; f < F; ++f) pp.pose_per_scale |= buf->T_scale[s][f] != nullptr;
  pp.no_ssim = (d->flags & VSL_FLAG_NO_SSIM) ? 1 : 0;
  pp.forward_only = fwd_only ? 1 : 0;
  pp.invK = buf->inv_K;
  pp.B = sp.B = d->batch; pp.H = sp.H = d->height; pp.W = sp.W = d->width; pp.S = sp.S = S; pp.F = sp.F = F;
  pp.g = make_geo(d);
  pp.wpix = 1.0f / ((float)d->batch * d->height * d->width);
  pp.partials = ws + pl.off_partials;
  pp.K = buf->K;
  for (int f = 0; f < F; ++f) {
    pp.src[f] = buf->source[f]; pp.P[f] = buf->P[f];
    for (int s = 0; s < S; ++s) pp.T[s][f] = buf->T_scale[s][f] ? buf->T_scale[s][f] : buf->T[f];
  }
  for (int s = 0; s < S; ++s) {
    int e = d->scale_ids[s];
    int hs = d->height >> e, wsz = d->width >> e;
    pp.hs[s] = sp.hs[s] = hs; pp.ws[s] = sp.ws[s] = wsz;
    pp.scale_h[s] = sp.scale_h[s] = (float)hs / (float)d->height;
    pp.scale_w[s] = sp.scale_w[s] = (float)wsz / (float)d->width;
    pp.identity_scale[s] = sp.identity_scale[s] = (e == 0);
    pp.level_shift[s] = sp.level_shift[s] = e;
    pp.disp[s] = sp.disp[s] = buf->disp[s];
    pp.noise[s] = buf->noise[s];
    pp.mask[s] = automask ? buf->mask[s] : nullptr;
    pp.winner[s] = buf->winner[s];
    pp.pmask[s] = automask ? nullptr : buf->predictive_mask[s];  // the reference only uses it without automasking
    pp.gpmask[s] = (pp.pmask[s] && !fwd_only) ? buf->grad_predictive_mask[s] : nullptr;
    if (pp.pmask[s] && !pp.gpmask[s] && !fwd_only) return VSL_ERR_NULL_POINTER;
    pp.side_depth[s] = buf->side_depth[s];
    pp.side_any |= buf->side_depth[s] != nullptr;
    for (int f = 0; f < F; ++f) {
      pp.side_sample[s][f] = buf->side_sample[s][f];
      pp.side_color[s][f] = buf->side_color[s][f];
      pp.side_any |= buf->side_sample[s][f] != nullptr || buf->side_color[s][f] != nullptr;
    }
    pp.gD[s] = (e == 0 && !fwd_only) ? buf->grad_disp_photo[s] : nullptr;
    pp.gpart[s] = (e == 0) ? nullptr : ws + pl.off_gpart[s];
    sp.gpart[s] = pp.gpart[s];
    sp.img[s] = buf->target[s];
    sp.gsmooth[s] = buf->grad_disp_smooth[s];
    pp.tgts[s] = buf->target[s];
    // phase_smooth is switched on by a non-null gsmooth; forward-only runs it for the sums and writes nothing there
    pp.gsmooth[s] = fwd_only ? (float*)buf->losses : buf->grad_disp_smooth[s];
    sp.gphoto[s] = fwd_only ? nullptr : buf->grad_disp_photo[s];
    sp.scale_id[s] = e + d->smooth_level_bias;
  }
  sp.partials = pp.partials;
  sp.lossb = ws + pl.off_lossb;
  sp.smoothb = ws + pl.off_smoothb;
  sp.gradP = fwd_only ? nullptr : buf->grad_P;
  sp.norm = fwd_only ? nullptr : buf->smooth_norm;
  sp.tw = pl.tw; sp.th = pl.th; sp.tiles_x = pl.tiles_x; sp.tiles_y = pl.tiles_y;
  sp.log_tw = ilog2(pl.tw); sp.log_th = (pl.th & (pl.th - 1)) == 0 ? ilog2(pl.th) : -1;
  sp.losses = buf->losses;
  sp.counter = (unsigned*)(ws + pl.off_counter);
  sp.chunks0 = pl.chunks0;
  sp.tiles_per_image = pl.tiles_x * pl.tiles_y;
  sp.kpartial = pl.kpartial;
  sp.smooth_weight = d->smooth_weight;

  // sp.counter: zero on entry (vsl_loss_workspace_init, once per workspace), reset by the last epilogue block
  if (event_before) VSL_CUDA_OK(cudaEventRecord((cudaEvent_t)event_before, st));
  int rc;
  const bool bf16 = d->image_dtype == VSL_DTYPE_BF16;
  if (pmask) {  // --predictive_mask kernels: run-time rounding selectors only (one instantiation each)
#define VSL_PMASK_LAUNCH(IMG)                                                                                                  \
    if (avg && F == 2) rc = launch_photometric_impl<TileCfg<32, 16, 2, 256, IMG, true, true>, false>(pp, pl, d->batch, st);    \
    else if (avg) rc = launch_photometric_impl<TileCfg<32, 8, 3, 256, IMG, true, true>, false>(pp, pl, d->batch, st);          \
    else if (F == 1) rc = launch_photometric_impl<TileCfg<32, 16, 1, 256, IMG, false, true>, false>(pp, pl, d->batch, st);     \
    else if (F == 2) rc = launch_photometric_impl<TileCfg<32, 16, 2, 256, IMG, false, true>, false>(pp, pl, d->batch, st);     \
    else rc = launch_photometric_impl<TileCfg<32, 8, 3, 256, IMG, false, true>, false>(pp, pl, d->batch, st);
    if (bf16) { VSL_PMASK_LAUNCH(bf16_t) } else { VSL_PMASK_LAUNCH(float) }
#undef VSL_PMASK_LAUNCH
  } else if (avg) {
    if (bf16) {  // rarely used together: run-time rounding selectors only
      if (F == 2) rc = launch_photometric_impl<TileCfg<32, 16, 2, 256, bf16_t, true>, false>(pp, pl, d->batch, st);
      else rc = launch_photometric_impl<TileCfg<32, 8, 3, 256, bf16_t, true>, false>(pp, pl, d->batch, st);
    } else if (F == 2) rc = launch_photometric<TileCfg<32, 16, 2, 256, float, true>>(pp, pl, d->batch, st);
    else rc = launch_photometric<TileCfg<32, 8, 3, 256, float, true>>(pp, pl, d->batch, st);
  } else if (d->image_dtype == VSL_DTYPE_BF16) {
    if (F == 1) rc = launch_photometric<TileCfg<32, 16, 1, 256, bf16_t>>(pp, pl, d->batch, st);
    else if (F == 2) rc = launch_photometric<TileCfg<32, 16, 2, 256, bf16_t>>(pp, pl, d->batch, st);
    else rc = launch_photometric<TileCfg<32, VSL_F3_TILE_H, 3, 512, bf16_t>>(pp, pl, d->batch, st);
  } else {
    if (F == 1) rc = launch_photometric<TileCfg<32, 16, 1, 256>>(pp, pl, d->batch, st);
    else if (F == 2) rc = launch_photometric<TileCfg<32, 16, 2, 256>>(pp, pl, d->batch, st);
    else rc = launch_photometric<TileCfg<32, VSL_F3_TILE_H, 3, 512>>(pp, pl, d->batch, st);
  }
  if (rc != VSL_OK) return rc;
  if (event_after) VSL_CUDA_OK(cudaEventRecord((cudaEvent_t)event_after, st));
  {
    int gx = 1;
    sp.epilogue_blocks = 0;
    for (int s = 0; s < S; ++s) {
      int nchunk = (sp.hs[s] * sp.ws[s] + kChunk - 1) / kChunk;
      int active = sp.identity_scale[s] ? 1 : nchunk;
      if (active > gx) gx = active;
      sp.epilogue_blocks += (unsigned)active * d->batch;
    }
    // programmatic launch right behind k_photometric (not when an event has to sit between the two)
    VSL_CUDA_OK(launch_after(event_after == nullptr, k_epilogue, dim3(gx, d->batch, S), dim3(kSmallNT), 0, st, sp));
  }
  VSL_CUDA_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_loss_combine_grads(const VslDesc* d, const float* upstream, const VslLossBuffers* buf,
                           float* const grad_disp[VSL_MAX_SCALES], float* grad_P_out, float* grad_T_out, void* stream) {
  if (!desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (d->flags & VSL_FLAG_FORWARD_ONLY) return VSL_ERR_UNSUPPORTED;  // a forward-only pass left no gradients behind
  if (!upstream || !buf || !grad_disp) return VSL_ERR_NULL_POINTER;
  CombineParams cp = {};
  cp.up = upstream; cp.B = d->batch; cp.S = d->num_scales; cp.F = d->num_src;
  cp.chunks0 = (d->height * d->width + kCombineChunk - 1) / kCombineChunk;
  cp.smooth_weight = d->smooth_weight;
  cp.gradP = buf->grad_P; cp.gradP_out = grad_P_out; cp.gradT_out = grad_T_out; cp.K = buf->K;
  cp.pose_per_scale = 0;
  for (int s = 0; s < d->num_scales; ++s)
    for (int f = 0; f < d->num_src; ++f) cp.pose_per_scale |= buf->T_scale[s][f] != nullptr;
  if (cp.pose_per_scale && grad_P_out) return VSL_ERR_UNSUPPORTED;  // per-scale poses report dL/dT only
  if (grad_T_out && !buf->K) return VSL_ERR_NULL_POINTER;
  cp.norm = buf->smooth_norm;
  if (!cp.norm) return VSL_ERR_NULL_POINTER;
  for (int s = 0; s < d->num_scales; ++s) {
    if (!buf->grad_disp_photo[s] || !buf->grad_disp_smooth[s] || !grad_disp[s]) return VSL_ERR_NULL_POINTER;
    int e = d->scale_ids[s];
    cp.n[s] = (d->height >> e) * (d->width >> e);
    cp.scale_id[s] = e + d->smooth_level_bias;
    cp.gphoto[s] = buf->grad_disp_photo[s]; cp.gsmooth[s] = buf->grad_disp_smooth[s]; cp.out[s] = grad_disp[s];
    cp.vec4[s] = cp.n[s] % 4 == 0 && (((uintptr_t)cp.gphoto[s] | (uintptr_t)cp.gsmooth[s] | (uintptr_t)cp.out[s]) & 15u) == 0;
  }
  if ((grad_P_out || grad_T_out) && !buf->grad_P) return VSL_ERR_NULL_POINTER;
  dim3 grid(cp.chunks0, d->batch, d->num_scales);
  VSL_CUDA_OK(launch_after(true, k_combine, grid, dim3(kSmallNT), 0, (cudaStream_t)stream, cp));
  VSL_CUDA_OK(cudaGetLastError());
  return VSL_OK;
}

int vsl_warp_forward(const VslDesc* d, int scale_index, const float* disp, const float* inv_K,
                     const float* const P[VSL_MAX_SRC], const void* const source[VSL_MAX_SRC], float* depth,
                     float* const sample[VSL_MAX_SRC], float* const color[VSL_MAX_SRC], void* stream) {
  if (!desc_ok(d)) return VSL_ERR_BAD_DESC;
  if (scale_index < 0 || scale_index >= d->num_scales) return VSL_ERR_BAD_DESC;
  if ((d->image_dtype != VSL_DTYPE_F32 && d->image_dtype != VSL_DTYPE_BF16) || (d->flags & VSL_FLAG_V1_MULTISCALE))
    return VSL_ERR_UNSUPPORTED;
  if (!disp || !inv_K || !P || !source) return VSL_ERR_NULL_POINTER;
  WarpParams wp = {};
  int e = d->scale_ids[scale_index];
  wp.disp = disp; wp.invK = inv_K; wp.depth = depth;
  wp.B = d->batch; wp.H = d->height; wp.W = d->width; wp.F = d->num_src;
  wp.hs = d->height >> e; wp.ws = d->width >> e; wp.identity = (e == 0);
  wp.scale_h = (float)wp.hs / (float)d->height; wp.scale_w = (float)wp.ws / (float)d->width;
  wp.g = make_geo(d);
  for (int f = 0; f < d->num_src; ++f) {
    if (!P[f]) return VSL_ERR_NULL_POINTER;
    wp.P[f] = P[f];
    wp.src[f] = source[f];
    wp.sample[f] = sample ? sample[f] : nullptr;
    wp.color[f] = color ? color[f] : nullptr;
    if (wp.color[f] && !wp.src[f]) return VSL_ERR_NULL_POINTER;
  }
  size_t n = (size_t)d->batch * d->height * d->width;
  if (d->image_dtype == VSL_DTYPE_BF16) k_warp_forward<bf16_t><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wp);
  else k_warp_forward<float><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wp);
  VSL_CUDA_OK(cudaGetLastError());
  return VSL_OK;
}

}  // extern "C"
