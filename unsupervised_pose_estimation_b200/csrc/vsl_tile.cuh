// Tile logic of the fused photometric kernel (forward + backward in one pass).
//
// One CTA owns a TW x TH tile of target pixels of one image and walks every scale and source
// frame for it.  Per CTA, once: target tile (+2 halo, reflect-mapped), target window moments,
// identity-reprojection losses (trainer.py:620-633; independent of scale and pose).  Per scale:
// warp all source frames on the tile +2 halo (trainer.py:500-537), SSIM+L1 per window on the tile
// +1 halo (trainer.py:543-555), tie-break noise + min/arg-min (trainer.py:654-670), then the
// adjoint: SSIM 3x3 gather with the reflection-pad fold, bilinear, projection, depth.
//
// Each phase is a function of (tile, thread id) so the same code runs as CUDA threads between
// __syncthreads() and, in tests/emul, as a host loop over thread ids.
#pragma once
#include "vsl_math.cuh"

namespace vsl {

constexpr int kMaxScales = 4;
constexpr int kMaxSrc = 4;

struct PhotoParams {
  const void* tgt;                // [B,3,H,W]  fp32 or bf16 (TileCfg::Img)
  const void* tgts[kMaxScales];   // target pyramid inputs[("color", 0, s)] [B,3,hs,ws] (smoothness edge weights)
  float* gsmooth[kMaxScales];     // [B,hs,ws] d smooth_s / d(norm disp_s), or null: smoothness not evaluated here
  const void* src[kMaxSrc];       // [B,3,H,W]
  const float* disp[kMaxScales];  // [B,1,hs,ws]
  const float* invK;              // [B,4,4]
  const float* P[kMaxSrc];        // [B,3,4] = (K@T_f)[:, :3, :], or null when T[f] is given
  const float* K;                 // [B,4,4]  inputs[("K", 0)]
  const float* T[kMaxScales][kMaxSrc];  // [B,4,4] cam_T_cam / stereo_T per (scale, frame): P is then formed in the
                                  // kernel (layers.py:254).  The same pointer for every scale except with posecnn,
                                  // whose translation is rescaled by each level's mean inverse depth (trainer.py:516-525)
  const float* noise[kMaxScales]; // [B,F,H,W]
  float* mask[kMaxScales];        // [B,H,W] or null
  unsigned char* winner[kMaxScales];  // [B,H,W] or null: arg-min channel (identity f: f, warped f: F + f; AVG: 0 / 1)
  const float* pmask[kMaxScales]; // [B,F,H,W] --predictive_mask, already at the warp resolution; or null
  float* gpmask[kMaxScales];      // [B,F,H,W] d(min_loss/s)/d pmask
  float* gD[kMaxScales];          // identity levels: [B,H,W] d(min_loss/s)/d disp_s, written directly
  float* gpart[kMaxScales];       // other levels: per-CTA partial up-sample adjoints [numCTA][ncy][ncx]
  float* partials;                // [numCTA][S][1 + F*12]
  // optional side outputs of generate_images_pred for the interior pixels (trainer.py:506, :532-537); fp32
  float* side_depth[kMaxScales];
  float* side_sample[kMaxScales][kMaxSrc];
  float* side_color[kMaxScales][kMaxSrc];
  int side_any;                   // some side output pointer is set (one uniform test in the warp loop)
  int forward_only;               // VSL_FLAG_FORWARD_ONLY: no adjoint state is kept, no gradient is written
  int B, H, W, S, F;
  int hs[kMaxScales], ws[kMaxScales];
  float scale_h[kMaxScales], scale_w[kMaxScales];
  int identity_scale[kMaxScales];  // up-sample is the identity (level size == H x W)
  int level_shift[kMaxScales];     // log2(W / ws[s]): levels are exact power-of-two reductions (checked by the ABI)
  GeoConst g;
  float wpix;                     // 1 / (B*H*W): weight of one pixel in min_loss/s
  int automask;                   // 0: --disable_automasking (no identity candidates, no noise, no mask)
  int no_ssim;                    // 1: --no_ssim (reprojection loss = mean_c L1)
  int pose_per_scale;             // 1: T differs per scale (posecnn), P is re-formed at every scale
};

template <int TW_, int TH_, int F_, int NT_, class Img_ = float, bool AVG_ = false, bool PMASK_ = false>
struct TileCfg {
  typedef Img_ Img;  // storage type of the colour images
  static constexpr bool PMASK = PMASK_;  // --predictive_mask compiled in (costs ~2 % when merely present)
  static constexpr int TW = TW_, TH = TH_, F = F_, NT = NT_;
  static constexpr int kLogTW = TW_ == 32 ? 5 : (TW_ == 16 ? 4 : (TW_ == 64 ? 6 : -1));
  // --avg_reprojection (trainer.py:629-630, 649-650): the frames' losses are averaged before the minimum,
  // so when the warped average wins EVERY frame receives gradient: one coefficient record per frame
  static constexpr bool AVG = AVG_;
  static constexpr int NREC = AVG_ ? F_ : 1;
  static constexpr int RW = TW + 4, RH = TH + 4, RN = RW * RH;  // region: tile + 2 halo (warped / target pixels)
  static constexpr int WW = TW + 2, WH = TH + 2, WN = WW * WH;  // windows: tile + 1 halo (SSIM centres)
  static constexpr int IN = TW * TH;                            // interior
  // shared memory layout, in floats
  static constexpr int oT = 0;                  // target region           [3][RN]
  static constexpr int oTS = oT + 3 * RN;       // target mu_y, sigma_y    [6][WN]
  static constexpr int oX = oTS + 6 * WN;       // warped / source region  [F][3][RN]
  static constexpr int oId = oX + F * 3 * RN;   // identity losses         [F][WN]
  static constexpr int oCoef = oId + F * WN;    // per window: CoefRec (winner's SSIM adjoint coefficients + winner id)
  static constexpr int oG = oCoef + 12 * WN * NREC;  // d warped / d(ix,iy)  [F][3][IN] pairs (d/d ix, d/d iy)
  static constexpr int oRed = oG + F * 6 * IN;  // block-reduction scratch [NT/32][1 + F*12]
  static constexpr int oGD = oRed + (NT / 32) * (1 + F * 12);  // d/d(up-sampled disp) of the tile [IN]
  static constexpr int oH = oGD + IN;           // row-reduced adjoint [TH][TW/2 + 2]
  static constexpr int oP = oH + TH * (TW / 2 + 2);  // this image's projection matrices [F][12]
  static constexpr int oInvK = oP + F * 12;     // first three rows of inv_K [12]
  // staged by asynchronous copies (stage4) while earlier phases compute:
  static constexpr int NZC = AVG_ ? 1 : F_;     // tie-break noise channels (trainer.py:654-657)
  static constexpr int oNz = oInvK + 12;        // noise of the current scale on the window grid [NZC][WN]
  static constexpr int DW = TW + 4, DH = TH + 4;  // window of disp_s under the region (clamped to the level)
  static constexpr int oDisp = oNz + NZC * WN;  // [DH][DW]
  static constexpr int oZ = oDisp + DW * DH;    // depth of the interior pixels at the current scale [IN]
  static constexpr int oRedS = oZ + IN;         // per-warp smoothness sums [NT/32][4]
  // target pyramid pixels under the tile's own pixels of a level s >= 1, +1 halo (smoothness edge weights)
  static constexpr int SW = TW / 2 + 2, SH = TH / 2 + 2, SN = SW * SH;
  // fp32 images are staged with asynchronous copies; bf16 images need a conversion, so their tiles are loaded into
  // registers in one batch (phase_load_tiles) and their level pixels are prefetched into registers a phase ahead
  static constexpr bool kStageImg = sizeof(Img_) == 4;
  static constexpr int oImgS = oRedS + (NT / 32) * 4;   // [3][SN]
  static constexpr int kFloats = oImgS + 3 * SN;
  static constexpr int kBytes = kFloats * 4;
  static constexpr int kPartial = 1 + F * 12;   // photometric partials per (CTA, scale): loss, dL/dP
  static constexpr int kPartialAll = kPartial + 4;  // + smoothness: sum d, sum |dx| e, sum |dy| e, sum g d
  static_assert(oCoef % 4 == 0, "CoefRec needs 16-byte alignment");
  static_assert(oG % 2 == 0 && oX % 2 == 0, "pair storage needs 8-byte alignment");
};

// One record per SSIM window: the 9 adjoint coefficients (A,B,C per channel) of the frame that won the
// per-pixel minimum, zeros when an identity candidate won; 48 bytes so the adjoint gathers it with
// three 128-bit shared loads.
struct alignas(16) CoefRec {
  float c[9];
  float m;     // predictive-mask value of the frame at this window (1 without --predictive_mask)
  float pad1;
  int idx;  // winning source frame, or -1 (identity won / window outside the image)
};

template <class C>
struct ThreadState {
  float loss;
  float dP[C::F * 12];
};

struct TileCtx {
  int b, x0, y0;  // image index, tile origin (pixels)
  int cta;        // linear CTA id
};

// ---- phase: projection matrices of this image, P_f = (K @ T_f)[:3,:]  (layers.py:254) -------------
template <class C>
VSL_HD void phase_pose(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  for (int k = tid; k < C::F * 12; k += C::NT) {
    const int f = k / 12, e = k - f * 12, i = e >> 2, n = e & 3;
    float v;
    if (p.T[s][f]) {
      const float* Kb = p.K + t.b * 16 + i * 4;
      const float* Tb = p.T[s][f] + t.b * 16 + n;
      v = dot4kt(Kb[0], Tb[0], Kb[1], Tb[4], Kb[2], Tb[8], Kb[3], Tb[12], g.arith);
    } else {
      v = p.P[f][t.b * 12 + e];
    }
    sm[C::oP + k] = v;
  }
  if (s == 0 && tid < 12) sm[C::oInvK + tid] = p.invK[t.b * 16 + tid];
}

// ---- staging (asynchronous): window of disp_s under the region, noise of the window grid --------------
// Level coordinates of the staged disp window: origin (y0/r - 2, x0/r - 2), (TH/r + 4) x (TW/r + 4) entries
// (row stride DW), r = 2^e the level's ratio; entries outside the level repeat the border (never selected
// with a non-zero weight by ups_tap, but they keep every tap index in range).  H and W are multiples of r.
template <class C>
VSL_HD void disp_window(const PhotoParams& p, const TileCtx& t, int s, int& cy0, int& cx0, int& rows, int& cols) {
  const int e = p.level_shift[s];
  cy0 = (t.y0 >> e) - 2; cx0 = (t.x0 >> e) - 2;
  rows = (C::TH >> e) + 4; cols = (C::TW >> e) + 4;
}
template <class C>
VSL_HD void phase_stage_disp(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  int cy0, cx0, rows, cols;
  disp_window<C>(p, t, s, cy0, cx0, rows, cols);
  const int hs = p.hs[s], ws = p.ws[s];
  const float* disp = p.disp[s] + (size_t)t.b * hs * ws;
  // 64 threads per row (cols <= TW + 4 <= 64): no division by the run-time window width
  static_assert(C::DW <= 64 && C::NT % 64 == 0, "one row of the disp window per 64 threads");
  for (int ry = tid >> 6, rx = tid & 63; ry < rows; ry += C::NT / 64) {
    if (rx >= cols) continue;
    int yy = cy0 + ry, xx = cx0 + rx;
    yy = yy < 0 ? 0 : (yy > hs - 1 ? hs - 1 : yy);
    xx = xx < 0 ? 0 : (xx > ws - 1 ? ws - 1 : xx);
    stage4(sm + C::oDisp + ry * C::DW + rx, disp + yy * ws + xx);
  }
  if (C::kStageImg && p.gsmooth[s] && !p.identity_scale[s]) {
    // target pyramid level s under the tile's own level pixels, +1 halo, for phase_smooth
    const int e = p.level_shift[s];
    const int oy = (t.y0 >> e) - 1, ox = (t.x0 >> e) - 1, nrow = (C::TH >> e) + 2, ncol = (C::TW >> e) + 2;
    const float* img = (const float*)p.tgts[s] + (size_t)t.b * 3 * hs * ws;
    static_assert(C::SW <= 32, "one row of the image window per warp");
    for (int ry = tid >> 5, rx = tid & 31; ry < nrow; ry += C::NT / 32) {
      if (rx >= ncol) continue;
      int yy = oy + ry, xx = ox + rx;
      yy = yy < 0 ? 0 : (yy > hs - 1 ? hs - 1 : yy);
      xx = xx < 0 ? 0 : (xx > ws - 1 ? ws - 1 : xx);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        stage4(sm + C::oImgS + c * C::SN + ry * C::SW + rx, img + (size_t)c * hs * ws + yy * ws + xx);
    }
  }
  stage_commit();
}
template <class C>
VSL_HD void phase_stage_noise(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  if (p.automask) {
    const int HW = p.H * p.W;
    const float* nz = p.noise[s] + (size_t)t.b * C::NZC * HW;
    for (int i = tid; i < C::WN; i += C::NT) {
      const int wy = i / C::WW, wx = i - wy * C::WW;
      const int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
#pragma unroll
        for (int f = 0; f < C::NZC; ++f) stage4(sm + C::oNz + f * C::WN + i, nz + (size_t)f * HW + gy * p.W + gx);
      }
    }
  }
  stage_commit();
}

template <class C>
struct XLayout;  // defined below

// ---- bf16 images: the level-s target pixels phase_smooth reads (same window as phase_stage_disp stages for
// fp32), loaded into registers a phase ahead and parked in shared memory once the phase in between is done
template <class C>
VSL_HD void phase_prefetch_level(const PhotoParams& p, const TileCtx& t, int s, int tid, float (&pre)[3]) {
  pre[0] = pre[1] = pre[2] = 0.f;
  if (!(p.gsmooth[s] && !p.identity_scale[s])) return;
  const int e = p.level_shift[s], hs = p.hs[s], ws = p.ws[s];
  const int oy = (t.y0 >> e) - 1, ox = (t.x0 >> e) - 1, nrow = (C::TH >> e) + 2, ncol = (C::TW >> e) + 2;
  const int ry = tid / C::SW, rx = tid - ry * C::SW;
  if (ry >= nrow || rx >= ncol) return;
  int yy = oy + ry, xx = ox + rx;
  yy = yy < 0 ? 0 : (yy > hs - 1 ? hs - 1 : yy);
  xx = xx < 0 ? 0 : (xx > ws - 1 ? ws - 1 : xx);
  const typename C::Img* img = (const typename C::Img*)p.tgts[s] + (size_t)t.b * 3 * hs * ws + yy * ws + xx;
#pragma unroll
  for (int c = 0; c < 3; ++c) pre[c] = ldimg(img, (size_t)c * hs * ws);
}
template <class C>
VSL_HD void phase_store_level(float* __restrict__ sm, int tid, const float (&pre)[3]) {
  static_assert(C::NT >= C::SN, "one level pixel per thread");
  if (tid < C::SN) {
#pragma unroll
    for (int c = 0; c < 3; ++c) sm[C::oImgS + c * C::SN + tid] = pre[c];
  }
}

// ---- bf16 images: target and source tiles (+2 halo, reflect-mapped) in ONE batch of loads -------------
// Same placement as phase_load_region + phase_load_sources; every load of the thread is issued before the
// first converted value is stored, so the tile costs one global-memory round trip instead of one per image.
template <class C>
VSL_HD void phase_load_tiles(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int tid, bool sources) {
  using XL = XLayout<C>;
  typedef typename C::Img Img;
  constexpr int IT = (C::RN + C::NT - 1) / C::NT;
  float* T = sm + C::oT;
  float* X = sm + C::oX;
  const int HW = p.H * p.W;
  const size_t img_off = (size_t)t.b * 3 * HW;
  float v[IT][(1 + C::F) * 3];
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * C::NT;
    const int ry = i / C::RW, rx = i - ry * C::RW;
    const int gy = t.y0 - 2 + ry, gx = t.x0 - 2 + rx;
    const bool valid = i < C::RN && gy >= -1 && gy <= p.H && gx >= -1 && gx <= p.W;
    const int o = reflect1(gy, p.H) * p.W + reflect1(gx, p.W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      v[it][c] = valid ? ldimg((const Img*)p.tgt, img_off + c * HW + o) : 0.f;
#pragma unroll
      for (int f = 0; f < C::F; ++f)
        v[it][(1 + f) * 3 + c] = (valid && sources) ? ldimg((const Img*)p.src[f], img_off + c * HW + o) : 0.f;
    }
  }
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * C::NT;
    if (i >= C::RN) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T[c * C::RN + i] = v[it][c];
      if (sources) {
#pragma unroll
        for (int f = 0; f < C::F; ++f) X[XL::at(f, c, i)] = v[it][(1 + f) * 3 + c];
      }
    }
  }
}

// ---- phase: load an image tile + 2 halo into a region buffer, reflect-mapped --------------------
template <class C>
VSL_HD void phase_load_region(const PhotoParams& p, const TileCtx& t, const typename C::Img* __restrict__ img_b,
                              float* __restrict__ dst, int tid) {
  const int HW = p.H * p.W;
  for (int i = tid; i < C::RN; i += C::NT) {
    int ry = i / C::RW, rx = i - ry * C::RW;
    int gy = t.y0 - 2 + ry, gx = t.x0 - 2 + rx;
    bool valid = gy >= -1 && gy <= p.H && gx >= -1 && gx <= p.W;
    int o = reflect1(gy, p.H) * p.W + reflect1(gx, p.W);
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c * C::RN + i] = valid ? ldimg(img_b, (size_t)c * HW + o) : 0.f;
  }
}

// ---- phase: target window moments mu_y, sigma_y on the window grid ------------------------------
template <class C>
VSL_HD void phase_target_stats(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int tid) {
  const float* T = sm + C::oT;
  float* TS = sm + C::oTS;
  for (int i = tid; i < C::WN; i += C::NT) {
    int wy = i / C::WW, wx = i - wy * C::WW;
    int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
    bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float mu = 0.f, sig = 0.f;
      if (inside) {
        const float* y = T + c * C::RN + (wy + 1) * C::RW + (wx + 1);
        float sy = 0.f, syy = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            float v = y[dy * C::RW + dx];
            if (dy == -1 && dx == -1) { sy = v; syy = mul_rn(v, v); }  // 0 + v: same bits downstream (see reproj_window)
            else { sy = add_rn(sy, v); syy = add_rn(syy, mul_rn(v, v)); }
          }
        mu = div9(sy);
        sig = sub_rn(div9(syy), mul_rn(mu, mu));
      }
      TS[c * C::WN + i] = mu;
      TS[(3 + c) * C::WN + i] = sig;
    }
  }
}

// ---- warped / source pixel storage -----------------------------------------------------------------
// Source frames are kept in PAIRS: frames 2k and 2k+1 interleaved as one fp32 pair per (channel, pixel),
// so the SSIM window loop loads and processes both with packed instructions.  An odd last frame is stored
// on its own.  Float offsets relative to sm + C::oX:
template <class C>
struct XLayout {
  static constexpr int NP = C::F / 2;           // pairs
  static constexpr int R = C::F % 2;            // leftover single frame
  VSL_HD static int pair_base(int pr, int c) { return (pr * 3 + c) * 2 * C::RN; }  // F2 array [RN]
  VSL_HD static int single_base(int c) { return NP * 6 * C::RN + c * C::RN; }      // float array [RN]
  VSL_HD static int at(int f, int c, int pos) {
    return f < 2 * NP ? pair_base(f >> 1, c) + 2 * pos + (f & 1) : single_base(c) + pos;
  }
};

// Tail of a window's reprojection loss from its nine window sums (per channel c: sum x, sum x^2, sum xy at
// 3c..3c+2) and the centre pixel's |y - x|: pooled means, SSIM per channel, 0.85 mean_c SSIM + 0.15 mean_c L1
// (layers.py:318-332, trainer.py:546-553).  One definition for every window-phase variant, so they agree bit for bit.
template <class C>
VSL_HD float window_loss_from_sums(const float (&sums)[9], const float (&l1)[3], const float* __restrict__ TS, int widx,
                                   int arith, SsimOut so[3], bool no_ssim = false) {
  if (no_ssim) {  // trainer.py:549-550: reprojection loss = mean_c |target - pred|
#pragma unroll
    for (int c = 0; c < 3; ++c) so[c].live = false;
    return mean3(l1[0], l1[1], l1[2], arith);
  }
  float mean[9], ss[3];
  div9_all<9>(sums, mean);  // one guard branch for the nine quotients
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    so[c] = ssim_from_means(mean[3 * c], mean[3 * c + 1], mean[3 * c + 2], TS[c * C::WN + widx], TS[(3 + c) * C::WN + widx]);
    ss[c] = so[c].val;
  }
  float ms = mean3(ss[0], ss[1], ss[2], arith);
  float ml = mean3(l1[0], l1[1], l1[2], arith);
  return add_rn(mul_rn(0.85f, ms), mul_rn(0.15f, ml));
}

// SSIM + L1 of one window for one frame X (3 channels `cs` floats apart, pixels XS floats apart: XS = 1
// for a single-frame buffer, 2 for one half of a pair buffer); returns the reprojection loss
// (trainer.py:546-553) and leaves the per-channel SSIM state in `so`.
template <class C, int XS = 1>
VSL_HD float reproj_window(const float* __restrict__ X, int cs, const float* __restrict__ T,
                           const float* __restrict__ TS, int wy, int wx, int widx, int arith, SsimOut so[3],
                           bool no_ssim = false) {
  float l1[3];
  const int center = (wy + 1) * C::RW + (wx + 1);
  if (no_ssim) {  // trainer.py:549-550: reprojection loss = mean_c |target - pred|
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      l1[c] = fabsf(sub_rn(T[c * C::RN + center], X[c * cs + XS * center]));
      so[c].live = false;
    }
    return mean3(l1[0], l1[1], l1[2], arith);
  }
  // Window sums in avg_pool2d's order (row-major, sequential).  The accumulator's initial 0 + v is skipped:
  // it changes the result only when every term is -0, and a -0 instead of +0 sum leaves every later value of
  // the SSIM chain unchanged (each is added to a non-zero constant before it is used).
  float sums[9];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* x = X + c * cs + XS * center;
    const float* y = T + c * C::RN + center;
    float sx = 0.f, sxx = 0.f, sxy = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        float xv = x[XS * (dy * C::RW + dx)], yv = y[dy * C::RW + dx];
        if (dy == -1 && dx == -1) {
          sx = xv; sxx = mul_rn(xv, xv); sxy = mul_rn(xv, yv);
        } else {
          sx = add_rn(sx, xv);
          sxx = add_rn(sxx, mul_rn(xv, xv));
          sxy = add_rn(sxy, mul_rn(xv, yv));
        }
      }
    sums[3 * c] = sx; sums[3 * c + 1] = sxx; sums[3 * c + 2] = sxy;
    l1[c] = fabsf(sub_rn(y[0], x[0]));
  }
  return window_loss_from_sums<C>(sums, l1, TS, widx, arith, so);
}

// The same for a PAIR of frames at once (X2: pair storage of channel 0, channels 2*RN apart).  Returns
// both losses; `sums` keeps the window sums (sx, sxx, sxy per channel) so the caller can rebuild the
// SSIM state of whichever frame wins without holding both in registers.
struct PairSums { F2 sx[3], sxx[3], sxy[3]; };
template <class C>
VSL_HD F2 reproj_window_pair(const float* __restrict__ X2, const float* __restrict__ T,
                             const float* __restrict__ TS, int wy, int wx, int widx, int arith, float onef,
                             PairSums& sums, bool no_ssim = false) {
  F2 ss[3], l1[3];
  const F2 one = splat(onef);
  const int center = (wy + 1) * C::RW + (wx + 1);
  if (no_ssim) {  // trainer.py:549-550
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      F2 d = sub2(splat(T[c * C::RN + center]), reinterpret_cast<const F2*>(X2 + c * 2 * C::RN)[center]);
      l1[c] = f2(fabsf(d.x), fabsf(d.y));
    }
    return mean3_2(l1[0], l1[1], l1[2], arith);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const F2* x = reinterpret_cast<const F2*>(X2 + c * 2 * C::RN) + center;
    const float* y = T + c * C::RN + center;
    F2 sx = splat(0.f), sxx = splat(0.f), sxy = splat(0.f);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        F2 xv = x[dy * C::RW + dx];
        F2 yv = splat(y[dy * C::RW + dx]);
        if (dy == -1 && dx == -1) {  // 0 + v skipped, see reproj_window
          sx = xv; sxx = mul2(xv, xv); sxy = mul2(xv, yv);
        } else {
          sx = add2(sx, xv);
          sxx = addp(sxx, xv, xv, one);
          sxy = addp(sxy, xv, yv, one);
        }
      }
    sums.sx[c] = sx; sums.sxx[c] = sxx; sums.sxy[c] = sxy;
    F2 d = sub2(splat(y[0]), x[0]);
    l1[c] = f2(fabsf(d.x), fabsf(d.y));
  }
  {
    F2 sv[9], mean[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) { sv[3 * c] = sums.sx[c]; sv[3 * c + 1] = sums.sxx[c]; sv[3 * c + 2] = sums.sxy[c]; }
    div9_2_all<9>(sv, mean);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      ss[c] = ssim_val2_means(mean[3 * c], mean[3 * c + 1], mean[3 * c + 2], TS[c * C::WN + widx], TS[(3 + c) * C::WN + widx], one);
  }
  F2 ms = mean3_2(ss[0], ss[1], ss[2], arith);
  F2 ml = mean3_2(l1[0], l1[1], l1[2], arith);
  return addp(mul2_packed(splat(0.15f), ml), splat(0.85f), ms, one);  // 0.85 ms + 0.15 ml, each product rounded
}

// ---- phase: stage every source frame's tile (+2 halo, reflect-mapped) in the X storage -----------
template <class C>
VSL_HD void phase_load_sources(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int tid) {
  using XL = XLayout<C>;
  float* X = sm + C::oX;
  const int HW = p.H * p.W;
  const size_t img_off = (size_t)t.b * 3 * HW;
  for (int i = tid; i < C::RN; i += C::NT) {
    int ry = i / C::RW, rx = i - ry * C::RW;
    int gy = t.y0 - 2 + ry, gx = t.x0 - 2 + rx;
    bool valid = gy >= -1 && gy <= p.H && gx >= -1 && gx <= p.W;
    int o = reflect1(gy, p.H) * p.W + reflect1(gx, p.W);
#pragma unroll
    for (int pr = 0; pr < XL::NP; ++pr)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        F2 v = splat(0.f);
        if (valid) v = f2(ldimg((const typename C::Img*)p.src[2 * pr], img_off + c * HW + o),
                          ldimg((const typename C::Img*)p.src[2 * pr + 1], img_off + c * HW + o));
        reinterpret_cast<F2*>(X + XL::pair_base(pr, c))[i] = v;
      }
    if (XL::R) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        X[XL::single_base(c) + i] = valid ? ldimg((const typename C::Img*)p.src[C::F - 1], img_off + c * HW + o) : 0.f;
    }
  }
}

// ---- staging (asynchronous, fp32 images): target and source tiles in one batch of copies ---------------
// Same placement as phase_load_region + phase_load_sources; pixels outside the padded image are zero-filled
// with ordinary stores (disjoint addresses).
template <class C>
VSL_HD void phase_stage_images(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int tid, bool sources) {
  using XL = XLayout<C>;
  float* T = sm + C::oT;
  float* X = sm + C::oX;
  const int HW = p.H * p.W;
  const size_t img_off = (size_t)t.b * 3 * HW;
  const float* tg = (const float*)p.tgt + img_off;
  for (int i = tid; i < C::RN; i += C::NT) {
    int ry = i / C::RW, rx = i - ry * C::RW;
    int gy = t.y0 - 2 + ry, gx = t.x0 - 2 + rx;
    bool valid = gy >= -1 && gy <= p.H && gx >= -1 && gx <= p.W;
    int o = reflect1(gy, p.H) * p.W + reflect1(gx, p.W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (valid) stage4(T + c * C::RN + i, tg + c * HW + o);
      else T[c * C::RN + i] = 0.f;
    }
    if (sources) {
#pragma unroll
      for (int f = 0; f < C::F; ++f) {
        const float* sf = (const float*)p.src[f] + img_off;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float* dst = X + XL::at(f, c, i);
          if (valid) stage4(dst, sf + c * HW + o);
          else *dst = 0.f;
        }
      }
    }
  }
  stage_commit();
}

// ---- phase: identity reprojection losses of all source frames (trainer.py:620-633) --------------
template <class C>
VSL_HD void phase_identity(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int tid) {
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* TS = sm + C::oTS;
  const float* X = sm + C::oX;
  float* Id = sm + C::oId;
  for (int i = tid; i < C::WN; i += C::NT) {
    int wy = i / C::WW, wx = i - wy * C::WW;
    int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
    bool inside = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
#pragma unroll
    for (int pr = 0; pr < XL::NP; ++pr) {
      F2 v = splat(0.f);
      if (inside) {
        PairSums sums;
        v = reproj_window_pair<C>(X + XL::pair_base(pr, 0), T, TS, wy, wx, i, g.arith, g.one, sums, p.no_ssim != 0);
      }
      Id[(2 * pr) * C::WN + i] = v.x;
      Id[(2 * pr + 1) * C::WN + i] = v.y;
    }
    if (XL::R) {
      float v = 0.f;
      if (inside) {
        SsimOut so[3];
        v = reproj_window<C>(X + XL::single_base(0), C::RN, T, TS, wy, wx, i, g.arith, so, p.no_ssim != 0);
      }
      Id[(C::F - 1) * C::WN + i] = v;
    }
  }
}

// ---- phase: warp every source frame on the region for scale s -----------------------------------
template <class C>
VSL_HD void phase_warp(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int s,
                       int tid) {
  using XL = XLayout<C>;
  float* X = sm + C::oX;
  float* G = sm + C::oG;
  const int HW = p.H * p.W;
  const float* invK = sm + C::oInvK;
  int cy0, cx0, drows, dcols;
  disp_window<C>(p, t, s, cy0, cx0, drows, dcols);
  for (int i = tid; i < C::RN; i += C::NT) {
    int ry = i / C::RW, rx = i - ry * C::RW;
    int gy = t.y0 - 2 + ry, gx = t.x0 - 2 + rx;
    bool valid = gy >= -1 && gy <= p.H && gx >= -1 && gx <= p.W;
    float val[C::F][3];
#pragma unroll
    for (int f = 0; f < C::F; ++f)
#pragma unroll
      for (int c = 0; c < 3; ++c) val[f][c] = 0.f;
    if (valid) {
      int v = reflect1(gy, p.H), u = reflect1(gx, p.W);
      float D = upsample_disp_staged(sm + C::oDisp, C::DW, cy0, cx0, p.hs[s], p.ws[s], p.scale_h[s], p.scale_w[s],
                                     p.identity_scale[s] != 0, v, u, g.arith);
      Cam cam = backproject_pixel(D, invK, u, v, g);
      bool interior = ry >= 2 && ry < C::TH + 2 && rx >= 2 && rx < C::TW + 2 && gy < p.H && gx < p.W;
      int j = (ry - 2) * C::TW + (rx - 2);
      const bool keep = interior && !p.forward_only;  // state the adjoint needs
      if (keep) sm[C::oZ + j] = cam.z;  // the adjoint re-forms the camera point from it
      // all frames' taps are requested before any is consumed, so their latencies overlap
      Proj pr[C::F];
      Taps tp[C::F];
      float tap[C::F][3][4];
#pragma unroll
      for (int f = 0; f < C::F; ++f) {
        pr[f] = project_pixel(cam, sm + C::oP + f * 12, g);
        tp[f] = bilinear_taps(pr[f], p.W, p.H);
        const typename C::Img* img = (const typename C::Img*)p.src[f] + (size_t)t.b * 3 * HW + pr[f].y0 * p.W + pr[f].x0;
        int dx = tp[f].x1ok ? 1 : 0, dy = tp[f].y1ok ? p.W : 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const typename C::Img* q = img + c * HW;
          tap[f][c][0] = ldimg(q, 0); tap[f][c][1] = ldimg(q, dx);
          tap[f][c][2] = ldimg(q, dy); tap[f][c][3] = ldimg(q, dy + dx);
        }
      }
#pragma unroll
      for (int f = 0; f < C::F; ++f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float vnw = tap[f][c][0], vne = tap[f][c][1], vsw = tap[f][c][2], vse = tap[f][c][3];
          val[f][c] = bilinear_value(tp[f], vnw, vne, vsw, vse, g.arith);
          // grid_sampler_2d_backward's d out / d(ix, iy); zero where the border clip is active
          float ddx = pr[f].inx ? ((vne - vnw) * tp[f].wy1 + (vse - vsw) * tp[f].wy0) : 0.f;
          float ddy = pr[f].iny ? ((vsw - vnw) * tp[f].wx1 + (vse - vne) * tp[f].wx0) : 0.f;
          if (keep) reinterpret_cast<F2*>(G)[(f * 3 + c) * C::IN + j] = f2(ddx, ddy);
        }
      }
      if (p.side_any && interior) {  // the reference's outputs[("depth" | "sample" | "color", ...)] for this pixel
        const size_t o = (size_t)t.b * HW + (size_t)gy * p.W + gx;
        if (p.side_depth[s]) p.side_depth[s][o] = cam.z;
#pragma unroll
        for (int f = 0; f < C::F; ++f) {
          if (p.side_sample[s][f]) {
            p.side_sample[s][f][2 * o] = pr[f].gx;
            p.side_sample[s][f][2 * o + 1] = pr[f].gy;
          }
          if (p.side_color[s][f]) {
            float* q = p.side_color[s][f] + (size_t)t.b * 3 * HW + (size_t)gy * p.W + gx;
#pragma unroll
            for (int c = 0; c < 3; ++c) q[(size_t)c * HW] = val[f][c];
          }
        }
      }
    }
#pragma unroll
    for (int pr = 0; pr < XL::NP; ++pr)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        reinterpret_cast<F2*>(X + XL::pair_base(pr, c))[i] = f2(val[2 * pr][c], val[2 * pr + 1][c]);
    if (XL::R) {
#pragma unroll
      for (int c = 0; c < 3; ++c) X[XL::single_base(c) + i] = val[C::F - 1][c];
    }
  }
}

// ---- phase: edge-aware smoothness of disp_s on the CTA's own pixels of level s -----------------------
// get_smooth_loss (layers.py:286-299) with the normalisation of trainer.py:676-677 factored out: the
// per-image mean of disp_s is only known once every tile is done, and norm = disp * inv with
// inv = 1 / (mean + 1e-7), so |norm_a - norm_b| = |inv| |d_a - d_b|.  The tile accumulates the un-normalised
// sums; k_epilogue applies inv, k_combine the chain rule through the mean.  Per level pixel: the four
// incident edges (the left / up ones are re-evaluated rather than exchanged between threads).
//   acc[0] += d, acc[1] += |d - d_right| e, acc[2] += |d - d_down| e, acc[3] += g d
//   g = d smooth / d norm (for inv > 0) = cx (w_right - w_left) + cy (w_down - w_up), w = sgn(delta) e
VSL_HD float smooth_edge_weight(const float a[3], const float b[3]) {  // exp(-mean_c |img_a - img_b|), layers.py:293-297
  float g = fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2]);
  return expf(-g * (1.0f / 3.0f));
}
VSL_HD float sgn_of(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
template <class C>
VSL_HD void phase_smooth(const PhotoParams& p, const TileCtx& t, const float* __restrict__ sm, int s, int tid,
                         float (&acc)[4]) {
  if (!p.gsmooth[s]) return;
  static_assert((C::TW & (C::TW - 1)) == 0, "tile width must be a power of two");
  const int e = p.level_shift[s];
  const int cw = C::TW >> e, ch = C::TH >> e;  // the tile's own pixels at this level
  const int hs = p.hs[s], ws = p.ws[s], n = hs * ws;
  int cy0, cx0, rows, cols;
  disp_window<C>(p, t, s, cy0, cx0, rows, cols);
  float* g_out = p.gsmooth[s] + (size_t)t.b * n;
  const float cx = 1.0f / ((float)p.B * hs * (ws - 1)), cy = 1.0f / ((float)p.B * (hs - 1) * ws);
  const bool from_tile = p.identity_scale[s] != 0;  // level 0: the target tile is in shared memory
  const float* T = sm + C::oT;
  for (int i = tid; i < cw * ch; i += C::NT) {
    const int iy = i >> (C::kLogTW - e), ix = i & (cw - 1);  // cw is a power of two: the compiler cannot know, the mask helps
    const int y = (t.y0 >> e) + iy, x = (t.x0 >> e) + ix;
    if (y >= hs || x >= ws) continue;
    const int o = y * ws + x;
    const float* d = sm + C::oDisp + (y - cy0) * C::DW + (x - cx0);
    auto pixel = [&](int dy, int dx, float v[3]) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[c] = from_tile ? T[c * C::RN + (iy + 2 + dy) * C::RW + (ix + 2 + dx)]
               : sm[C::oImgS + c * C::SN + (iy + 1 + dy) * C::SW + (ix + 1 + dx)];
    };
    float c0[3], cn[3];
    pixel(0, 0, c0);
    const float d0 = d[0];
    float g = 0.f;
    if (x + 1 < ws) {
      pixel(0, 1, cn);
      const float e = smooth_edge_weight(c0, cn), dl = d0 - d[1];
      acc[1] += fabsf(dl) * e;
      g += cx * sgn_of(dl) * e;
    }
    if (y + 1 < hs) {
      pixel(1, 0, cn);
      const float e = smooth_edge_weight(c0, cn), dl = d0 - d[C::DW];
      acc[2] += fabsf(dl) * e;
      g += cy * sgn_of(dl) * e;
    }
    if (x > 0) {
      pixel(0, -1, cn);
      g -= cx * sgn_of(d[-1] - d0) * smooth_edge_weight(cn, c0);
    }
    if (y > 0) {
      pixel(-1, 0, cn);
      g -= cy * sgn_of(d[-C::DW] - d0) * smooth_edge_weight(cn, c0);
    }
    if (!p.forward_only) g_out[o] = g;
    acc[0] += d0;
    acc[3] += g * d0;
  }
}

// ---- phase: per-window losses, auto-mask arg-min, adjoint coefficients ---------------------------
// adjoint coefficients of one frame's SSIM at a window: A (d/d mu_x), B (d/d E[x^2]), C (d/d E[xy]) per
// channel, pre-scaled by the pixel weight, 0.85/3, the clamp mask, -1/2 and the 1/9 of the mean filter
VSL_HD void window_coefs(const SsimOut so[3], const float* __restrict__ TS, int WN, int i, float kc, float coef[9]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float dmu, dexx, dexy;
    ssim_r_grads(so[c], TS[c * WN + i], dmu, dexx, dexy);
    float k = so[c].live ? kc : 0.f;
    coef[c] = k * dmu;
    coef[3 + c] = k * dexx;
    coef[6 + c] = k * dexy;
  }
}

VSL_HD void store_rec(CoefRec* __restrict__ rec, const float coef[9], int idx, float m = 1.0f) {
  CoefRec r;
#pragma unroll
  for (int k = 0; k < 9; ++k) r.c[k] = coef[k];
  r.m = m; r.pad1 = 0.f; r.idx = idx;
  *rec = r;
}

// A record is always read whole, as three 128-bit shared loads.  Left to itself the compiler fetches the two words
// of the last quarter it needs (c[8], idx) with scalar loads, which at the records' 48-byte pitch are 4-way bank
// conflicted: 16 shared-memory wavefronts per record instead of 12 (r2c profile, vsl_tile.cuh Rec loads).
VSL_HD CoefRec load_rec(const CoefRec* __restrict__ rec) {
#if defined(__CUDA_ARCH__)
  CoefRec r;
  const unsigned a = (unsigned)__cvta_generic_to_shared(rec);
  float4 q0, q1, q2;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w) : "r"(a));
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+32];" : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w) : "r"(a));
  r.c[0] = q0.x; r.c[1] = q0.y; r.c[2] = q0.z; r.c[3] = q0.w;
  r.c[4] = q1.x; r.c[5] = q1.y; r.c[6] = q1.z; r.c[7] = q1.w;
  r.c[8] = q2.x; r.m = q2.y; r.pad1 = q2.z; r.idx = __float_as_int(q2.w);
  return r;
#else
  return *rec;
#endif
}

// --predictive_mask (trainer.py:635-642, only reached with --disable_automasking): every frame's
// reprojection loss is multiplied by a learnt per-pixel mask before the minimum.  The kernel receives the
// mask at the warp resolution, multiplies (one rounding, like `reprojection_losses *= mask`), scales the
// frame's adjoint by it and returns d/d mask = loss of the winning frame.
template <class C>
VSL_HD bool has_pmask(const PhotoParams& p, int s) { return C::PMASK && p.pmask[s] != nullptr; }
template <class C>
VSL_HD float pmask_at(const PhotoParams& p, const TileCtx& t, int s, int f, int gy, int gx) {
  if (!C::PMASK) return 1.0f;
  return p.pmask[s] ? p.pmask[s][((size_t)t.b * C::F + f) * p.H * p.W + gy * p.W + gx] : 1.0f;
}
template <class C>
VSL_HD void store_gpmask(const PhotoParams& p, const TileCtx& t, int s, int f, int gy, int gx, float v) {
  if (C::PMASK && p.gpmask[s]) p.gpmask[s][((size_t)t.b * C::F + f) * p.H * p.W + gy * p.W + gx] = v;
}

// torch.mean over the frame dimension (sequential sum times float(1/F)), trainer.py:629-630, 649-650
template <int F>
VSL_HD float mean_frames(const float* l) {
  float acc = l[0];
#pragma unroll
  for (int f = 1; f < F; ++f) acc = add_rn(acc, l[f]);
  return mul_rn(acc, 1.0f / (float)F);
}

// --avg_reprojection: candidates are (mean_f identity_f + noise, mean_f reprojection_f); when the warped
// mean wins, every frame's record is live with its coefficients scaled by 1/F.
template <class C>
VSL_HD void phase_windows_avg(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int s,
                              int tid, ThreadState<C>& ts) {
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* TS = sm + C::oTS;
  const float* X = sm + C::oX;
  const float* Id = sm + C::oId;
  CoefRec* Rec = reinterpret_cast<CoefRec*>(sm + C::oCoef);
  const int HW = p.H * p.W;
  const float kc = p.wpix * (0.85f / 3.0f) * (-0.5f) * (1.0f / 9.0f) * (1.0f / (float)C::F);
  for (int i = tid; i < C::WN; i += C::NT) {
    int wy = i / C::WW, wx = i - wy * C::WW;
    int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
    float coef[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) coef[k] = 0.f;
    if (!(gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)) {
#pragma unroll
      for (int f = 0; f < C::F; ++f) store_rec(Rec + f * C::WN + i, coef, -1);
      continue;
    }
    float best = INFINITY;
    if (p.automask) {
      float idl[C::F];
#pragma unroll
      for (int f = 0; f < C::F; ++f) idl[f] = Id[f * C::WN + i];
      best = add_rn(mean_frames<C::F>(idl), mul_rn(sm[C::oNz + i], 1e-5f));
    }
    float l[C::F], lraw[C::F];
#pragma unroll
    for (int f = 0; f < C::F; ++f) {
      SsimOut so[3];
      if (f < 2 * XL::NP)  // one half of a pair buffer (pixel stride 2) or the un-paired odd last frame
        lraw[f] = reproj_window<C, 2>(X + XL::pair_base(f >> 1, 0) + (f & 1), 2 * C::RN, T, TS, wy, wx, i, g.arith, so,
                                      p.no_ssim != 0);
      else
        lraw[f] = reproj_window<C, 1>(X + XL::single_base(0), C::RN, T, TS, wy, wx, i, g.arith, so, p.no_ssim != 0);
      const float m = pmask_at<C>(p, t, s, f, gy, gx);
      l[f] = has_pmask<C>(p, s) ? mul_rn(lraw[f], m) : lraw[f];
#pragma unroll
      for (int k = 0; k < 9; ++k) coef[k] = 0.f;
      if (!p.no_ssim && !p.forward_only) window_coefs(so, TS, C::WN, i, kc * m, coef);
      if (!p.forward_only) store_rec(Rec + f * C::WN + i, coef, f, m);
    }
    const float avg = mean_frames<C::F>(l);
    const bool warped = avg < best;  // identity first: ties keep the identity channel
    if (!warped && !p.forward_only) {
#pragma unroll
      for (int f = 0; f < C::F; ++f) Rec[f * C::WN + i].idx = -1;
    }
    bool interior = wy >= 1 && wy <= C::TH && wx >= 1 && wx <= C::TW;
    if (interior) {
      ts.loss += warped ? avg : best;
      if (p.mask[s]) p.mask[s][(size_t)t.b * HW + gy * p.W + gx] = warped ? 1.f : 0.f;
      if (p.winner[s]) p.winner[s][(size_t)t.b * HW + gy * p.W + gx] = warped ? 1 : 0;
#pragma unroll
      for (int f = 0; f < C::F; ++f)
        store_gpmask<C>(p, t, s, f, gy, gx, warped ? p.wpix * (1.0f / (float)C::F) * lraw[f] : 0.f);
    }
  }
}

template <class C>
VSL_HD void phase_windows(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int s,
                          int tid, ThreadState<C>& ts) {
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* TS = sm + C::oTS;
  const float* X = sm + C::oX;
  const float* Id = sm + C::oId;
  CoefRec* Rec = reinterpret_cast<CoefRec*>(sm + C::oCoef);
  const int HW = p.H * p.W;
  const float kc = p.wpix * (0.85f / 3.0f) * (-0.5f) * (1.0f / 9.0f);
  for (int i = tid; i < C::WN; i += C::NT) {
    int wy = i / C::WW, wx = i - wy * C::WW;
    int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
    float coef[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) coef[k] = 0.f;
    if (!(gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)) {
      store_rec(Rec + i, coef, -1);
      continue;
    }
    // identity candidates first (trainer.py:654-659: identity + 1e-5 * randn; cat(identity, reprojection));
    // strict '<' everywhere: ties keep the lower index, like torch.min
    float best = INFINITY;
    int bidx = -1;
    if (p.automask) {
#pragma unroll
      for (int f = 0; f < C::F; ++f) {
        float cand = add_rn(Id[f * C::WN + i], mul_rn(sm[C::oNz + f * C::WN + i], 1e-5f));
        if (cand < best) { best = cand; bidx = f; }
      }
    }
    float lraw[C::F], wm = 1.0f;  // un-masked losses (for d/d pmask), mask value of the current winner
#pragma unroll
    for (int pr = 0; pr < XL::NP; ++pr) {
      PairSums sums;
      F2 l = reproj_window_pair<C>(X + XL::pair_base(pr, 0), T, TS, wy, wx, i, g.arith, g.one, sums, p.no_ssim != 0);
      lraw[2 * pr] = l.x; lraw[2 * pr + 1] = l.y;
      const float m0 = pmask_at<C>(p, t, s, 2 * pr, gy, gx), m1 = pmask_at<C>(p, t, s, 2 * pr + 1, gy, gx);
      if (has_pmask<C>(p, s)) l = f2(mul_rn(l.x, m0), mul_rn(l.y, m1));
      int win = -1;
      if (l.x < best) { best = l.x; win = 0; }
      if (l.y < best) { best = l.y; win = 1; }
      if (win >= 0) { bidx = C::F + 2 * pr + win; wm = win == 0 ? m0 : m1; }
      if (win >= 0 && !p.no_ssim && !p.forward_only) {  // rebuild the winner's SSIM state from its window sums (same ops, same bits)
        SsimOut so[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          so[c] = win == 0 ? ssim_from_sums(sums.sx[c].x, sums.sxx[c].x, sums.sxy[c].x, TS[c * C::WN + i], TS[(3 + c) * C::WN + i])
                           : ssim_from_sums(sums.sx[c].y, sums.sxx[c].y, sums.sxy[c].y, TS[c * C::WN + i], TS[(3 + c) * C::WN + i]);
        window_coefs(so, TS, C::WN, i, kc * wm, coef);
      }
    }
    if (XL::R) {
      SsimOut so[3];
      float l = reproj_window<C>(X + XL::single_base(0), C::RN, T, TS, wy, wx, i, g.arith, so, p.no_ssim != 0);
      lraw[C::F - 1] = l;
      const float m = pmask_at<C>(p, t, s, C::F - 1, gy, gx);
      if (has_pmask<C>(p, s)) l = mul_rn(l, m);
      if (l < best) {
        best = l;
        bidx = 2 * C::F - 1;
        wm = m;
#pragma unroll
        for (int k = 0; k < 9; ++k) coef[k] = 0.f;  // a pair frame may have set them before losing to this one
        if (!p.no_ssim && !p.forward_only) window_coefs(so, TS, C::WN, i, kc * wm, coef);
      }
    }
    bool warped = bidx >= C::F;
    if (!p.forward_only) store_rec(Rec + i, coef, warped ? bidx - C::F : -1, wm);
    bool interior = wy >= 1 && wy <= C::TH && wx >= 1 && wx <= C::TW;
    if (interior) {
      ts.loss += best;
      if (p.mask[s]) p.mask[s][(size_t)t.b * HW + gy * p.W + gx] = warped ? 1.f : 0.f;
      if (p.winner[s]) p.winner[s][(size_t)t.b * HW + gy * p.W + gx] = (unsigned char)bidx;
#pragma unroll
      for (int f = 0; f < C::F; ++f)
        store_gpmask<C>(p, t, s, f, gy, gx, (warped && bidx - C::F == f) ? p.wpix * lraw[f] : 0.f);
    }
  }
}

#if defined(__CUDACC__)
// Two-source-frame variant (frames [0,-1,1], the reference default): lanes 2k / 2k+1 evaluate frame 0 / 1
// of the same window (the two halves of the pair storage) and exchange the losses by shuffle, so the
// 2*WN (window, frame) items fill the CTA's threads evenly (95 % of the slots instead of 80 %).  Same
// decisions as phase_windows; measured faster than the packed-pair evaluation for F = 2.
template <class C>
__device__ __forceinline__ void phase_windows_paired(const PhotoParams& p, const GeoConst& g, const TileCtx& t,
                                                     float* __restrict__ sm, int s, int tid, ThreadState<C>& ts) {
  static_assert(C::F == 2, "paired variant is for two source frames");
  static_assert(C::kLogTW > 0, "tile width must be a power of two");
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* TS = sm + C::oTS;
  const float* X = sm + C::oX;
  const float* Id = sm + C::oId;
  CoefRec* Rec = reinterpret_cast<CoefRec*>(sm + C::oCoef);
  const int HW = p.H * p.W;
  const float kc = p.wpix * (0.85f / 3.0f) * (-0.5f) * (1.0f / 9.0f);
  for (int base = 0; base < 2 * C::WN; base += C::NT) {  // uniform trip count: every lane reaches the shuffle
    const int item = base + tid;
    const int f = item & 1;
    const bool live = item < 2 * C::WN;
    // Item order: first the TW leftmost window columns of every row (a warp = 16 consecutive windows of ONE row x 2
    // frames = 32 consecutive words of the pair storage: conflict-free), then the two rightmost columns.  In plain
    // row-major order over the (TW + 2)-wide grid a warp straddles a row end every other time, and the jump of the
    // region pitch puts 4 of its 32 words on banks already taken (1.38 wavefronts per load in the r2d profile).
    constexpr int kMain = C::WH * 2 * C::TW;
    int wy, wx;
    if (item < kMain) { wy = item >> (C::kLogTW + 1); wx = (item & (2 * C::TW - 1)) >> 1; }
    else { const int rest = item - kMain; wy = rest >> 2; wx = C::TW + ((rest >> 1) & 1); }
    const int i = wy * C::WW + wx;
    int gy = t.y0 - 1 + wy, gx = t.x0 - 1 + wx;
    const bool inside = live && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
    float best = INFINITY, l = INFINITY, lraw = 0.f, m = 1.0f;
    bool id1 = false;  // the second identity candidate is the better one
    SsimOut so[3];
    if (inside) {
      if (p.automask) {
        float c0 = add_rn(Id[i], mul_rn(sm[C::oNz + i], 1e-5f));
        float c1 = add_rn(Id[C::WN + i], mul_rn(sm[C::oNz + C::WN + i], 1e-5f));
        id1 = c1 < c0;
        best = id1 ? c1 : c0;
      }
      l = lraw = reproj_window<C, 2>(X + XL::pair_base(0, 0) + f, 2 * C::RN, T, TS, wy, wx, i, g.arith, so, p.no_ssim != 0);
      if (has_pmask<C>(p, s)) {
        m = pmask_at<C>(p, t, s, f, gy, gx);
        l = mul_rn(lraw, m);
      }
    }
    const float other = __shfl_xor_sync(0xffffffffu, l, 1);
    if (!live) continue;
    // arg-min over (identity..., frame 0, frame 1) with ties to the lower index
    const float l0 = f == 0 ? l : other, l1 = f == 0 ? other : l;
    int win = -1;
    float mn = best;
    if (inside) {
      if (l0 < mn) { mn = l0; win = 0; }
      if (l1 < mn) { mn = l1; win = 1; }
    }
    float coef[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) coef[k] = 0.f;
    if (p.forward_only) {
      // no adjoint: nothing to record
    } else if (win == f) {
      if (!p.no_ssim) window_coefs(so, TS, C::WN, i, kc * m, coef);
      store_rec(Rec + i, coef, f, m);
    } else if (win < 0 && f == 0) {
      store_rec(Rec + i, coef, -1);
    }
    const bool interior = inside && wy >= 1 && wy <= C::TH && wx >= 1 && wx <= C::TW;
    if (interior) {
      store_gpmask<C>(p, t, s, f, gy, gx, win == f ? p.wpix * lraw : 0.f);
      if (f == 0) {
        ts.loss += mn;
        if (p.mask[s]) p.mask[s][(size_t)t.b * HW + gy * p.W + gx] = win >= 0 ? 1.f : 0.f;
        if (p.winner[s]) p.winner[s][(size_t)t.b * HW + gy * p.W + gx] = (unsigned char)(win >= 0 ? 2 + win : (id1 ? 1 : 0));
      }
    }
  }
}

#endif

// ---- phase: adjoint for the interior pixels of scale s -------------------------------------------
template <class C>
VSL_HD void phase_backward(const PhotoParams& p, const GeoConst& g, const TileCtx& t, float* __restrict__ sm, int s,
                           int tid, ThreadState<C>& ts) {
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* X = sm + C::oX;
  const CoefRec* Rec = reinterpret_cast<const CoefRec*>(sm + C::oCoef);
  const float* G = sm + C::oG;
  const int HW = p.H * p.W;
  const float* invK = sm + C::oInvK;
  const float kl1 = (p.no_ssim ? p.wpix * (1.0f / 3.0f) : p.wpix * (0.15f / 3.0f)) / (float)(C::AVG ? C::F : 1);
  for (int j = tid; j < C::IN; j += C::NT) {
    int iy = j / C::TW, ix = j - iy * C::TW;
    int gy = t.y0 + iy, gx = t.x0 + ix;
    if (gy >= p.H || gx >= p.W) {
      sm[C::oGD + j] = 0.f;
      continue;
    }
    // masked 3x3 gather of the winners' coefficients; reflected border rows/cols count twice
    float acc[C::F][9];
#pragma unroll
    for (int f = 0; f < C::F; ++f)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[f][k] = 0.f;
    unsigned used = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      float cy = ((dy == -1 && gy == 1) || (dy == 1 && gy == p.H - 2)) ? 2.f : 1.f;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        float cx = ((dx == -1 && gx == 1) || (dx == 1 && gx == p.W - 2)) ? 2.f : 1.f;
        const float cnt = cy * cx;
#pragma unroll
        for (int rr = 0; rr < C::NREC; ++rr) {
          const CoefRec rec = load_rec(Rec + rr * C::WN + (iy + 1 + dy) * C::WW + (ix + 1 + dx));  // 3 x 128-bit shared loads
          if (rec.idx >= 0) used |= 1u << rec.idx;
#pragma unroll
          for (int f = 0; f < C::F; ++f) {
            const float cf = rec.idx == f ? cnt : 0.f;  // records of lost windows hold zeros / idx -1
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[f][k] = fmaf(cf, rec.c[k], acc[f][k]);
          }
        }
      }
    }
    float gz = 0.f;
    Cam cam;
    cam.z = 0.f;
    if (used) {
      cam = backproject_z(sm[C::oZ + j], invK, gx, gy, g);
      const int center = (iy + 2) * C::RW + (ix + 2);
      const int wc = (iy + 1) * C::WW + (ix + 1);  // this pixel's own window
#pragma unroll
      for (int f = 0; f < C::F; ++f) {
        if (!(used & (1u << f))) continue;
        float gix = 0.f, giy = 0.f;
        // L1 term of the pixel's own window, if frame f is live there; scaled by its predictive-mask value
        const int own_idx = Rec[(C::AVG ? f : 0) * C::WN + wc].idx;
        const float k1 = (C::AVG ? own_idx >= 0 : own_idx == f)
                             ? (C::PMASK ? kl1 * Rec[(C::AVG ? f : 0) * C::WN + wc].m : kl1) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float xq = X[XL::at(f, c, center)], yq = T[c * C::RN + center];
          float gc = acc[f][c] + 2.f * xq * acc[f][3 + c] + yq * acc[f][6 + c];
          gc += (xq > yq) ? k1 : ((xq < yq) ? -k1 : 0.f);
          const F2 dg = reinterpret_cast<const F2*>(G)[(f * 3 + c) * C::IN + j];  // d warped / d(ix, iy)
          gix += gc * dg.x;
          giy += gc * dg.y;
        }
        const float* P = sm + C::oP + f * 12;
        float c0 = P[0] * cam.X + P[1] * cam.Y + P[2] * cam.Z + P[3];
        float c1 = P[4] * cam.X + P[5] * cam.Y + P[6] * cam.Z + P[7];
        float c2 = P[8] * cam.X + P[9] * cam.Y + P[10] * cam.Z + P[11];
        float iz = fast_rcp(c2 + g.eps);
        float g0 = gix * iz, g1 = giy * iz;
        float g2 = -(g0 * c0 + g1 * c1) * iz;
        float* dP = ts.dP + f * 12;
        dP[0] += g0 * cam.X; dP[1] += g0 * cam.Y; dP[2] += g0 * cam.Z; dP[3] += g0;
        dP[4] += g1 * cam.X; dP[5] += g1 * cam.Y; dP[6] += g1 * cam.Z; dP[7] += g1;
        dP[8] += g2 * cam.X; dP[9] += g2 * cam.Y; dP[10] += g2 * cam.Z; dP[11] += g2;
        float gX = g0 * P[0] + g1 * P[4] + g2 * P[8];
        float gY = g0 * P[1] + g1 * P[5] + g2 * P[9];
        float gZ = g0 * P[2] + g1 * P[6] + g2 * P[10];
        gz += gX * cam.rx + gY * cam.ry + gZ * cam.rz;
      }
    }
    // z = 1/(min_disp + range*D)  ->  dz/dD = -range * z^2
    const float gDv = -gz * g.disp_range * cam.z * cam.z;
    if (p.identity_scale[s]) p.gD[s][(size_t)t.b * HW + gy * p.W + gx] = gDv;
    else sm[C::oGD + j] = gDv;  // folded into d/d disp_s by phase_adjoint_rows / _cols
  }
}

#if defined(__CUDACC__)
// Two vertically adjacent interior pixels per thread: their 3x3 record neighbourhoods share two of three rows, so
// the pair reads 12 records instead of 18 (the adjoint phase waits on shared memory more than on anything else).
// Per pixel the accumulation order (rows top to bottom, columns left to right) and every operation are those of
// phase_backward, so the gradients have the same bits.  One record per window only (not --avg_reprojection).
#ifndef VSL_ADJ_PAIR
#define VSL_ADJ_PAIR 1
#endif
template <class C>
__device__ __forceinline__ void phase_backward_pair(const PhotoParams& p, const GeoConst& g, const TileCtx& t,
                                                    float* __restrict__ sm, int s, int tid, ThreadState<C>& ts) {
  static_assert(!C::AVG && C::TH % 2 == 0, "one record per window, even tile height");
  using XL = XLayout<C>;
  const float* T = sm + C::oT;
  const float* X = sm + C::oX;
  const CoefRec* Rec = reinterpret_cast<const CoefRec*>(sm + C::oCoef);
  const float* G = sm + C::oG;
  const int HW = p.H * p.W;
  const float* invK = sm + C::oInvK;
  const float kl1 = p.no_ssim ? p.wpix * (1.0f / 3.0f) : p.wpix * (0.15f / 3.0f);
  for (int item = tid; item < C::IN / 2; item += C::NT) {
    const int py = item / C::TW, ix = item - py * C::TW;
    const int iy0 = 2 * py;
    const int gx = t.x0 + ix;
    float acc[2][C::F][9];
    unsigned used[2] = {0u, 0u};
    int own_idx[2] = {-1, -1};
    float own_m[2] = {1.f, 1.f};
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int f = 0; f < C::F; ++f)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[q][f][k] = 0.f;
    const bool col_ok = gx < p.W;
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // record rows iy0 + r (window coordinates iy0 + r, i.e. pixel rows iy0 - 1 + r)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const CoefRec rec = load_rec(Rec + (iy0 + r) * C::WW + (ix + 1 + dx));
        const float cx = ((dx == -1 && gx == 1) || (dx == 1 && gx == p.W - 2)) ? 2.f : 1.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int dy = r - 1 - q;  // this record row relative to pixel q
          if (dy < -1 || dy > 1) continue;
          const int gy = t.y0 + iy0 + q;
          const float cy = ((dy == -1 && gy == 1) || (dy == 1 && gy == p.H - 2)) ? 2.f : 1.f;
          const float cnt = cy * cx;
          if (rec.idx >= 0) used[q] |= 1u << rec.idx;
          if (dy == 0 && dx == 0) { own_idx[q] = rec.idx; own_m[q] = rec.m; }
#pragma unroll
          for (int f = 0; f < C::F; ++f) {
            const float cf = rec.idx == f ? cnt : 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[q][f][k] = fmaf(cf, rec.c[k], acc[q][f][k]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int iy = iy0 + q, gy = t.y0 + iy, j = iy * C::TW + ix;
      if (gy >= p.H || !col_ok) {
        sm[C::oGD + j] = 0.f;
        continue;
      }
      float gz = 0.f;
      Cam cam;
      cam.z = 0.f;
      if (used[q]) {
        cam = backproject_z(sm[C::oZ + j], invK, gx, gy, g);
        const int center = (iy + 2) * C::RW + (ix + 2);
#pragma unroll
        for (int f = 0; f < C::F; ++f) {
          if (!(used[q] & (1u << f))) continue;
          float gix = 0.f, giy = 0.f;
          const float k1 = own_idx[q] == f ? (C::PMASK ? kl1 * own_m[q] : kl1) : 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float xq = X[XL::at(f, c, center)], yq = T[c * C::RN + center];
            float gc = acc[q][f][c] + 2.f * xq * acc[q][f][3 + c] + yq * acc[q][f][6 + c];
            gc += (xq > yq) ? k1 : ((xq < yq) ? -k1 : 0.f);
            const F2 dg = reinterpret_cast<const F2*>(G)[(f * 3 + c) * C::IN + j];  // d warped / d(ix, iy)
            gix += gc * dg.x;
            giy += gc * dg.y;
          }
          const float* P = sm + C::oP + f * 12;
          float c0 = P[0] * cam.X + P[1] * cam.Y + P[2] * cam.Z + P[3];
          float c1 = P[4] * cam.X + P[5] * cam.Y + P[6] * cam.Z + P[7];
          float c2 = P[8] * cam.X + P[9] * cam.Y + P[10] * cam.Z + P[11];
          float iz = fast_rcp(c2 + g.eps);
          float g0 = gix * iz, g1 = giy * iz;
          float g2 = -(g0 * c0 + g1 * c1) * iz;
          float* dP = ts.dP + f * 12;
          dP[0] += g0 * cam.X; dP[1] += g0 * cam.Y; dP[2] += g0 * cam.Z; dP[3] += g0;
          dP[4] += g1 * cam.X; dP[5] += g1 * cam.Y; dP[6] += g1 * cam.Z; dP[7] += g1;
          dP[8] += g2 * cam.X; dP[9] += g2 * cam.Y; dP[10] += g2 * cam.Z; dP[11] += g2;
          float gX = g0 * P[0] + g1 * P[4] + g2 * P[8];
          float gY = g0 * P[1] + g1 * P[5] + g2 * P[9];
          float gZ = g0 * P[2] + g1 * P[6] + g2 * P[10];
          gz += gX * cam.rx + gY * cam.ry + gZ * cam.rz;
        }
      }
      const float gDv = -gz * g.disp_range * cam.z * cam.z;
      if (p.identity_scale[s]) p.gD[s][(size_t)t.b * HW + gy * p.W + gx] = gDv;
      else sm[C::oGD + j] = gDv;
    }
  }
}
#endif

// ---- phases: adjoint of the bilinear up-sample of disp_s (trainer.py:500-501), tile-local part -----
// d/d disp_s[jy,jx] = sum over fine pixels o of wy(oy,jy) wx(ox,jx) gD[o]; the footprint of a coarse pixel is
// [r j - r/2, r j + 3r/2) per axis, so a TW x TH tile touches (TW/r + 2) x (TH/r + 2) coarse pixels.  The tile
// reduces its own fine pixels separably (rows, then columns) and stores one partial per touched coarse
// pixel; k_epilogue adds the <= 4 partials of every coarse pixel in a fixed order (no atomics).
// weight of fine pixel o in coarse pixel j for ratio R = 2^e: a triangle over the footprint
// t = o - (R j - R/2) in [0, 2R): (t + 0.5)/R rising, (2R - t - 0.5)/R falling — exactly the weights
// ups_tap() produces (all values are multiples of 2^-(e+1)) — except at the two borders, where the
// clamped source index gives the whole weight to the edge pixel.
template <int R>
VSL_HD float ups_weight(int o, int j, int size) {
  const int t = o - (R * j - R / 2);
  float w = ((t < R ? t : 2 * R - 1 - t) + 0.5f) * (1.0f / R);
  if (j == 0 && t < R) w = 1.0f;           // o < R/2: source index clamped to 0
  if (j == size - 1 && t >= R) w = 1.0f;   // last coarse pixel: no right/bottom neighbour
  return w;
}
template <class C, int R>
VSL_HD void adjoint_rows_r(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  const float* GD = sm + C::oGD;
  float* Hs = sm + C::oH;
  constexpr int ncx = C::TW / R + 2;
  for (int item = tid; item < ncx * C::TH; item += C::NT) {
    const int y = item / ncx, cj = item - y * ncx;
    const int jx = t.x0 / R - 1 + cj;
    float acc = 0.f;
    if (jx >= 0 && jx < p.ws[s]) {
      const int base = R * jx - R / 2 - t.x0;  // footprint start in tile coordinates
#pragma unroll
      for (int k = 0; k < 2 * R; ++k) {
        const int x = base + k;
        if (x >= 0 && x < C::TW) acc += ups_weight<R>(t.x0 + x, jx, p.ws[s]) * GD[y * C::TW + x];
      }
    }
    Hs[item] = acc;
  }
}
template <class C, int R>
VSL_HD void adjoint_cols_r(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  const float* Hs = sm + C::oH;
  constexpr int ncx = C::TW / R + 2, ncy = C::TH / R + 2;
  float* out = p.gpart[s] + (size_t)t.cta * ncx * ncy;
  for (int item = tid; item < ncx * ncy; item += C::NT) {
    const int cjy = item / ncx, cj = item - cjy * ncx;
    const int jy = t.y0 / R - 1 + cjy;
    float acc = 0.f;
    if (jy >= 0 && jy < p.hs[s]) {
      const int base = R * jy - R / 2 - t.y0;
#pragma unroll
      for (int k = 0; k < 2 * R; ++k) {
        const int y = base + k;
        if (y >= 0 && y < C::TH) acc += ups_weight<R>(t.y0 + y, jy, p.hs[s]) * Hs[y * ncx + cj];
      }
    }
    out[item] = acc;
  }
}
template <class C>
VSL_HD void phase_adjoint_rows(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  const int r = 1 << p.level_shift[s];
  if (r == 2) adjoint_rows_r<C, 2>(p, t, sm, s, tid);
  else if (r == 4) adjoint_rows_r<C, 4>(p, t, sm, s, tid);
  else adjoint_rows_r<C, 8>(p, t, sm, s, tid);
}
template <class C>
VSL_HD void phase_adjoint_cols(const PhotoParams& p, const TileCtx& t, float* __restrict__ sm, int s, int tid) {
  const int r = 1 << p.level_shift[s];
  if (r == 2) adjoint_cols_r<C, 2>(p, t, sm, s, tid);
  else if (r == 4) adjoint_cols_r<C, 4>(p, t, sm, s, tid);
  else adjoint_cols_r<C, 8>(p, t, sm, s, tid);
}

VSL_HD int ilog2(int v) {  // v a power of two
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
// sum of the tile partials that touch coarse pixel (jy, jx) of image b; tiles in row-major order
// lcw: log2 of the tile's width in pixels of this level (tile width and ratio are powers of two); ch: the tile's height
// in pixels of this level, lch its log2 or -1 when it is not a power of two (the 24-row tiles of three source frames)
VSL_HD float gather_adjoint_partials(const float* __restrict__ gpart, int b, int jy, int jx, int lcw, int ch, int lch,
                                     int tiles_x, int tiles_y) {
  const int cw = 1 << lcw, ncx = cw + 2, ncy = ch + 2;
  int ty_hi = lch >= 0 ? (jy + 1) >> lch : (jy + 1) / ch, tx_hi = (jx + 1) >> lcw;
  int ty_lo = jy - ch <= 0 ? 0 : (lch >= 0 ? (jy - 1) >> lch : (jy - 1) / ch), tx_lo = jx - cw <= 0 ? 0 : (jx - 1) >> lcw;
  if (ty_hi >= tiles_y) ty_hi = tiles_y - 1;
  if (tx_hi >= tiles_x) tx_hi = tiles_x - 1;
  float acc = 0.f;
  for (int ty = ty_lo; ty <= ty_hi; ++ty)
    for (int tx = tx_lo; tx <= tx_hi; ++tx) {
      int cjy = jy - (ty * ch - 1), cjx = jx - (tx * cw - 1);
      size_t cta = ((size_t)b * tiles_y + ty) * tiles_x + tx;
      acc += gpart[(cta * ncy + cjy) * ncx + cjx];
    }
  return acc;
}

}  // namespace vsl
