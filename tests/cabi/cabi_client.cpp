// A caller of libvsl_b200.so that knows nothing about Python or torch: plain C++ + the CUDA runtime, the
// entry points exactly as include/vsl.h declares them (tests/test_cabi_client.py builds and runs it).
//
//   cabi_client pyramid <in.bin> <out.bin>     in:  int32 B,H,W,levels ; uint8 frames [B,H,W,3]
//                                              out: uint8 levels 1.. [B,h,w,3] concatenated ; float32 last level [B,3,h,w]
//   cabi_client ssim <in.bin> <out.bin>        in:  int32 B,C,H,W ; float32 x[B,C,H,W] ; float32 y[B,C,H,W]
//                                              out: float32 [B,C,H,W]   (layers.py:318-332)
//   cabi_client loss <in.bin> <out.bin>        the fused loss, forward + backward (trainer.py:491-686, :312):
//                                              in:  int32 B,H,W,S,F ; float32 min_disp, disp_range, smooth_weight ;
//                                                   float32 target[s] [B,3,H>>s,W>>s] (S) ; source[f] [B,3,H,W] (F) ;
//                                                   disp[s] [B,1,H>>s,W>>s] (S) ; inv_K [B,4,4] ; K [B,4,4] ; T[f] [B,4,4] (F) ;
//                                                   noise[s] [B,F,H,W] (S) ; upstream [2S+1]
//                                              out: float32 losses [3S+1] ; mask[s] [B,H,W] (S) ; grad_disp[s] (S) ; grad_T [F][B][16]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/vsl.h"

#define CHECK_CUDA(e)                                                               \
  do {                                                                              \
    cudaError_t err__ = (e);                                                        \
    if (err__ != cudaSuccess) { std::fprintf(stderr, "cuda: %s\n", cudaGetErrorString(err__)); return 3; } \
  } while (0)
#define CHECK_VSL(e)                                                                \
  do {                                                                              \
    int rc__ = (e);                                                                 \
    if (rc__ != VSL_OK) { std::fprintf(stderr, "vsl: %s\n", vsl_status_string(rc__)); return 4; } \
  } while (0)

static bool read_file(const char* path, std::vector<unsigned char>& data) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  data.resize((size_t)n);
  bool ok = std::fread(data.data(), 1, (size_t)n, f) == (size_t)n;
  std::fclose(f);
  return ok;
}

static int run_pyramid(const std::vector<unsigned char>& in, FILE* out) {
  int32_t hdr[4];
  std::memcpy(hdr, in.data(), sizeof(hdr));
  const int B = hdr[0], H = hdr[1], W = hdr[2], L = hdr[3];
  VslPyramidDesc d = {VSL_ABI_VERSION, B, H, W, L, VSL_DTYPE_F32};
  const size_t ws_bytes = vsl_pyramid_workspace_bytes(&d);
  if (!ws_bytes) return 2;
  unsigned char *frames = nullptr, *ws = nullptr;
  const size_t n0 = (size_t)B * H * W * 3;
  CHECK_CUDA(cudaMalloc(&frames, n0));
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));   // cudaMalloc is 256-byte aligned
  CHECK_CUDA(cudaMemcpy(frames, in.data() + sizeof(hdr), n0, cudaMemcpyHostToDevice));
  void* levels[VSL_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
  uint8_t* levels_u8[VSL_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
  for (int s = 0; s < L; ++s) {
    const size_t n = (size_t)B * 3 * (H >> s) * (W >> s);
    CHECK_CUDA(cudaMalloc(&levels[s], n * sizeof(float)));
    if (s) CHECK_CUDA(cudaMalloc(&levels_u8[s], n));
  }
  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  CHECK_VSL(vsl_pyramid_plan(&d, ws, ws_bytes, st));
  CHECK_VSL(vsl_pyramid_forward(&d, frames, levels, levels_u8, ws, ws_bytes, st));
  CHECK_CUDA(cudaStreamSynchronize(st));
  for (int s = 1; s < L; ++s) {
    std::vector<unsigned char> host((size_t)B * 3 * (H >> s) * (W >> s));
    CHECK_CUDA(cudaMemcpy(host.data(), levels_u8[s], host.size(), cudaMemcpyDeviceToHost));
    std::fwrite(host.data(), 1, host.size(), out);
  }
  std::vector<float> last((size_t)B * 3 * (H >> (L - 1)) * (W >> (L - 1)));
  CHECK_CUDA(cudaMemcpy(last.data(), levels[L - 1], last.size() * sizeof(float), cudaMemcpyDeviceToHost));
  std::fwrite(last.data(), sizeof(float), last.size(), out);
  return 0;
}

static int run_ssim(const std::vector<unsigned char>& in, FILE* out) {
  int32_t hdr[4];
  std::memcpy(hdr, in.data(), sizeof(hdr));
  const int B = hdr[0], C = hdr[1], H = hdr[2], W = hdr[3];
  const size_t n = (size_t)B * C * H * W;
  float *x = nullptr, *y = nullptr, *o = nullptr;
  CHECK_CUDA(cudaMalloc(&x, n * sizeof(float)));
  CHECK_CUDA(cudaMalloc(&y, n * sizeof(float)));
  CHECK_CUDA(cudaMalloc(&o, n * sizeof(float)));
  CHECK_CUDA(cudaMemcpy(x, in.data() + sizeof(hdr), n * sizeof(float), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(y, in.data() + sizeof(hdr) + n * sizeof(float), n * sizeof(float), cudaMemcpyHostToDevice));
  CHECK_VSL(vsl_ssim_forward(B, C, H, W, x, y, o, nullptr));   // stream 0
  CHECK_CUDA(cudaDeviceSynchronize());
  std::vector<float> host(n);
  CHECK_CUDA(cudaMemcpy(host.data(), o, n * sizeof(float), cudaMemcpyDeviceToHost));
  std::fwrite(host.data(), sizeof(float), n, out);
  return 0;
}

static int run_loss(const std::vector<unsigned char>& in, FILE* out) {
  int32_t hdr[5];
  float opt[3];
  std::memcpy(hdr, in.data(), sizeof(hdr));
  std::memcpy(opt, in.data() + sizeof(hdr), sizeof(opt));
  const int B = hdr[0], H = hdr[1], W = hdr[2], S = hdr[3], F = hdr[4];
  VslDesc d = {};
  d.abi_version = VSL_ABI_VERSION; d.batch = B; d.height = H; d.width = W; d.num_scales = S; d.num_src = F;
  for (int s = 0; s < S; ++s) d.scale_ids[s] = s;
  d.flags = VSL_FLAG_AUTOMASK; d.image_dtype = VSL_DTYPE_F32; d.arith = 0;
  d.min_disp = opt[0]; d.disp_range = opt[1]; d.eps = 1e-7f; d.smooth_weight = opt[2];
  const float* src = (const float*)(in.data() + sizeof(hdr) + sizeof(opt));
  size_t off = 0;
  auto upload = [&](size_t n, float** dev) -> int {   // next n floats of the input file -> device
    CHECK_CUDA(cudaMalloc(dev, n * sizeof(float)));
    CHECK_CUDA(cudaMemcpy(*dev, src + off, n * sizeof(float), cudaMemcpyHostToDevice));
    off += n;
    return 0;
  };
  auto alloc = [&](size_t n, float** dev) -> int { CHECK_CUDA(cudaMalloc(dev, n * sizeof(float))); return 0; };
  VslLossBuffers buf = {};
  float* p = nullptr;
  size_t nlev[VSL_MAX_SCALES];
  const size_t n0 = (size_t)B * H * W;
  for (int s = 0; s < S; ++s) {
    nlev[s] = (size_t)B * (H >> s) * (W >> s);
    if (upload(3 * nlev[s], &p)) return 3;
    buf.target[s] = p;
  }
  for (int f = 0; f < F; ++f) { if (upload(3 * n0, &p)) return 3; buf.source[f] = p; }
  for (int s = 0; s < S; ++s) { if (upload(nlev[s], &p)) return 3; buf.disp[s] = p; }
  if (upload((size_t)B * 16, &p)) return 3; buf.inv_K = p;
  if (upload((size_t)B * 16, &p)) return 3; buf.K = p;
  for (int f = 0; f < F; ++f) { if (upload((size_t)B * 16, &p)) return 3; buf.T[f] = p; }
  for (int s = 0; s < S; ++s) { if (upload((size_t)F * n0, &p)) return 3; buf.noise[s] = p; }
  float* upstream = nullptr;
  if (upload((size_t)2 * S + 1, &upstream)) return 3;
  float *losses, *norm, *gradP, *gradT;
  if (alloc(3 * S + 1, &losses) || alloc((size_t)S * B * 2, &norm) || alloc((size_t)S * F * B * 12, &gradP) ||
      alloc((size_t)F * B * 16, &gradT)) return 3;
  buf.losses = losses; buf.smooth_norm = norm; buf.grad_P = gradP;
  float* grad_disp[VSL_MAX_SCALES] = {nullptr, nullptr, nullptr, nullptr};
  for (int s = 0; s < S; ++s) {
    if (alloc(n0, &buf.mask[s]) || alloc(nlev[s], &buf.grad_disp_photo[s]) || alloc(nlev[s], &buf.grad_disp_smooth[s]) ||
        alloc(nlev[s], &grad_disp[s])) return 3;
  }
  const size_t ws_bytes = vsl_loss_workspace_bytes(&d);
  if (!ws_bytes) return 2;
  void* ws = nullptr;
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  CHECK_VSL(vsl_loss_workspace_init(&d, ws, ws_bytes, st));
  CHECK_VSL(vsl_loss_forward_backward(&d, &buf, ws, ws_bytes, st));
  CHECK_VSL(vsl_loss_combine_grads(&d, upstream, &buf, grad_disp, nullptr, gradT, st));
  CHECK_CUDA(cudaStreamSynchronize(st));
  auto dump = [&](const float* dev, size_t n) -> int {
    std::vector<float> host(n);
    CHECK_CUDA(cudaMemcpy(host.data(), dev, n * sizeof(float), cudaMemcpyDeviceToHost));
    std::fwrite(host.data(), sizeof(float), n, out);
    return 0;
  };
  if (dump(losses, 3 * S + 1)) return 3;
  for (int s = 0; s < S; ++s) if (dump(buf.mask[s], n0)) return 3;
  for (int s = 0; s < S; ++s) if (dump(grad_disp[s], nlev[s])) return 3;
  if (dump(gradT, (size_t)F * B * 16)) return 3;
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 4) { std::fprintf(stderr, "usage: cabi_client pyramid|ssim|loss <in.bin> <out.bin>\n"); return 1; }
  if (vsl_abi_version() != VSL_ABI_VERSION) { std::fprintf(stderr, "ABI version mismatch\n"); return 1; }
  std::vector<unsigned char> in;
  if (!read_file(argv[2], in) || in.size() < 16) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
  FILE* out = std::fopen(argv[3], "wb");
  if (!out) return 1;
  int rc = std::strcmp(argv[1], "pyramid") == 0 ? run_pyramid(in, out)
           : std::strcmp(argv[1], "loss") == 0 ? run_loss(in, out) : run_ssim(in, out);
  std::fclose(out);
  return rc;
}
