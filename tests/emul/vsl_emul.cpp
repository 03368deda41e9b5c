// TEST-ONLY host build of the fused photometric kernel's tile logic.
//
// Compiles unsupervised_pose_estimation_b200/csrc/vsl_tile.cuh as plain C++ and runs every CTA /
// thread of k_photometric as nested host loops, so that indexing, halo handling, the reflection-pad
// adjoint and the analytic gradients can be checked against the oracle on a machine without a GPU.
// It is NOT part of libvsl_b200.so and nothing in the package loads it: the product has no CPU path.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../unsupervised_pose_estimation_b200/csrc/vsl_tile.cuh"

using namespace vsl;

template <class C>
static bool run_tiles(const PhotoParams& p, int tiles_x, int tiles_y) {
  // the "shared memory" sits between two guard zones filled with a sentinel; a phase that writes outside
  // its CTA's allocation trips the check after the tile (compute-sanitizer is not available on the GPU pool)
  constexpr int kGuard = 256;
  std::vector<float> sm_store(C::kFloats + 2 * kGuard + 4);
  float* sm_base = sm_store.data() + kGuard;
  while (reinterpret_cast<uintptr_t>(sm_base) & 15u) ++sm_base;  // CoefRec is 16-byte aligned
  const float kSentinel = -12345.678f;
  for (float& v : sm_store) v = kSentinel;
  struct { float* p; float* data() const { return p; } } sm{sm_base};
  auto guards_intact = [&]() {
    for (float* q = sm_store.data(); q < sm_base; ++q) if (*q != kSentinel) return false;
    for (float* q = sm_base + C::kFloats; q < sm_store.data() + sm_store.size(); ++q) if (*q != kSentinel) return false;
    return true;
  };
  std::vector<ThreadState<C>> ts(C::NT);
  for (int b = 0; b < p.B; ++b)
    for (int ty = 0; ty < tiles_y; ++ty)
      for (int tx = 0; tx < tiles_x; ++tx) {
        TileCtx t;
        t.b = b; t.x0 = tx * C::TW; t.y0 = ty * C::TH;
        t.cta = (b * tiles_y + ty) * tiles_x + tx;
        size_t img_off = (size_t)b * 3 * p.H * p.W;
        for (int tid = 0; tid < C::NT; ++tid) phase_load_region<C>(p, t, (const float*)p.tgt + img_off, sm.data() + C::oT, tid);
        for (int tid = 0; tid < C::NT; ++tid) phase_target_stats<C>(p, t, sm.data(), tid);
        for (int tid = 0; tid < C::NT; ++tid) phase_load_sources<C>(p, t, sm.data(), tid);
        for (int tid = 0; tid < C::NT; ++tid) phase_identity<C>(p, p.g, t, sm.data(), tid);
        for (int s = 0; s < p.S; ++s) {
          for (int tid = 0; tid < C::NT; ++tid) {
            ts[tid].loss = 0.f;
            for (int k = 0; k < C::F * 12; ++k) ts[tid].dP[k] = 0.f;
          }
          for (int tid = 0; tid < C::NT; ++tid) phase_pose<C>(p, p.g, t, sm.data(), s, tid);
          for (int tid = 0; tid < C::NT; ++tid) phase_stage_noise<C>(p, t, sm.data(), s, tid);
          for (int tid = 0; tid < C::NT; ++tid) phase_stage_disp<C>(p, t, sm.data(), s, tid);
          for (int tid = 0; tid < C::NT; ++tid) phase_warp<C>(p, p.g, t, sm.data(), s, tid);
          for (int tid = 0; tid < C::NT; ++tid) phase_windows<C>(p, p.g, t, sm.data(), s, tid, ts[tid]);
          for (int tid = 0; tid < C::NT; ++tid) phase_backward<C>(p, p.g, t, sm.data(), s, tid, ts[tid]);
          if (!p.identity_scale[s]) {
            for (int tid = 0; tid < C::NT; ++tid) phase_adjoint_rows<C>(p, t, sm.data(), s, tid);
            for (int tid = 0; tid < C::NT; ++tid) phase_adjoint_cols<C>(p, t, sm.data(), s, tid);
          }
          float* out = p.partials + ((size_t)t.cta * p.S + s) * C::kPartial;
          for (int k = 0; k < C::kPartial; ++k) out[k] = 0.f;
          for (int tid = 0; tid < C::NT; ++tid) {
            out[0] += ts[tid].loss;
            for (int k = 0; k < C::F * 12; ++k) out[1 + k] += ts[tid].dP[k];
          }
        }
        if (!guards_intact()) return false;
      }
  return true;
}

extern "C" int vsl_emul_photometric(int B, int H, int W, int S, int F, const int* scale_ids, const float* tgt,
                                    const float* const* src, const float* const* disp, const float* invK,
                                    const float* const* P, const float* const* noise, float min_disp,
                                    float disp_range, float eps, int arith, int tw, int th,
                                    float* const* mask, float* const* gdisp, float* gradP, double* loss_sums) {
  PhotoParams p;
  std::memset(&p, 0, sizeof(p));
  p.tgt = tgt; p.invK = invK; p.B = B; p.H = H; p.W = W; p.S = S; p.F = F;
  p.g.one = 1.0f; p.automask = 1; p.no_ssim = 0; p.pose_per_scale = 0; p.g.min_disp = min_disp; p.g.disp_range = disp_range; p.g.eps = eps; p.g.W = W; p.g.H = H;
  p.g.wm1 = (float)(W - 1); p.g.hm1 = (float)(H - 1);
  p.g.inv_wm1 = 1.0f / p.g.wm1; p.g.inv_hm1 = 1.0f / p.g.hm1; p.g.arith = arith;
  p.wpix = 1.0f / ((float)B * H * W);
  std::vector<std::vector<float>> gD(S), gpart(S);
  const int tiles_x0 = (W + tw - 1) / tw, tiles_y0 = (H + th - 1) / th;
  for (int f = 0; f < F; ++f) { p.src[f] = src[f]; p.P[f] = P[f]; }
  for (int s = 0; s < S; ++s) {
    int e = scale_ids[s];
    p.hs[s] = H >> e; p.ws[s] = W >> e;
    p.scale_h[s] = (float)p.hs[s] / (float)H; p.scale_w[s] = (float)p.ws[s] / (float)W;
    p.identity_scale[s] = e == 0;
    p.level_shift[s] = e;
    p.disp[s] = disp[s]; p.noise[s] = noise[s]; p.mask[s] = mask[s];
    gD[s].assign((size_t)B * H * W, 0.f);
    p.gD[s] = gD[s].data();
    int r = 1 << e;
    gpart[s].assign((size_t)tiles_x0 * tiles_y0 * B * (tw / r + 2) * (th / r + 2), 0.f);
    p.gpart[s] = gpart[s].data();
  }
  int tiles_x = (W + tw - 1) / tw, tiles_y = (H + th - 1) / th;
  int kpartial = 1 + F * 12;
  std::vector<float> partials((size_t)tiles_x * tiles_y * B * S * kpartial, 0.f);
  p.partials = partials.data();
  bool ok = false;
  bool guards = true;
#define TRY(TW, TH, FF) if (tw == TW && th == TH && F == FF) { guards = run_tiles<TileCfg<TW, TH, FF, 256>>(p, tiles_x, tiles_y); ok = true; }
  TRY(32, 16, 1) TRY(32, 16, 2) TRY(32, 16, 3) TRY(16, 8, 2) TRY(16, 8, 3)
#undef TRY
  if (!ok) return -4;
  if (!guards) return -7;  // a phase wrote outside the CTA's shared-memory allocation
  int tpi = tiles_x * tiles_y;
  for (int s = 0; s < S; ++s) {
    loss_sums[s] = 0.0;
    for (int b = 0; b < B; ++b) {
      for (int k = 0; k < kpartial; ++k) {
        double acc = 0.0;
        for (int tl = 0; tl < tpi; ++tl) acc += partials[(((size_t)b * tpi + tl) * S + s) * kpartial + k];
        if (k == 0) loss_sums[s] += acc;
        else gradP[((size_t)(s * F + (k - 1) / 12) * B + b) * 12 + (k - 1) % 12] = (float)acc;
      }
      int hs = p.hs[s], ws = p.ws[s];
      for (int i = 0; i < hs * ws; ++i)
        gdisp[s][(size_t)b * hs * ws + i] =
            p.identity_scale[s] ? gD[s][(size_t)b * H * W + i]
                                : gather_adjoint_partials(gpart[s].data(), b, i / ws, i % ws, ilog2(tw) - ilog2(W / ws),
                                                          th / (W / ws), ilog2(th) - ilog2(W / ws), tiles_x, tiles_y);
    }
  }
  return 0;
}
