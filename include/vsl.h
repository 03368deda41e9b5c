/*
 * vsl.h — C ABI of the B200-native view-synthesis loss library (libvsl_b200.so).
 *
 * Drop-in boundary for the view-synthesis loss path of
 * meghakalia/unsupervised_pose_estimation (a monodepth2 fork).  The reference has no
 * FFI layer: its boundary is the Python call surface of layers.py / trainer.py.  Each
 * entry point below names the reference interface it replaces (file:line in the
 * reference tree); the Python host side (unsupervised_pose_estimation_b200/layers.py,
 * trainer.py) binds these symbols with ctypes and mirrors the reference's classes.
 *
 * Conventions
 *  - plain C: pointers, sizes, PODs; no torch types.  `stream` is a cudaStream_t passed
 *    as void* (the caller passes torch.cuda.current_stream().cuda_stream).
 *  - the CALLER allocates and owns every buffer (device memory unless stated); the
 *    library never allocates, frees or keeps a pointer after the call returns.
 *  - all work is enqueued on `stream`; no call synchronises the device.
 *  - return value: VSL_OK (0) or a negative VslStatus; nothing throws across the ABI.
 *  - tensors are dense row-major ("contiguous") with the shapes given; images are NCHW.
 *  - arithmetic is fp32; images may be stored as fp32 or bf16 (VslDesc.image_dtype).
 *  - there is no CPU implementation behind this ABI: without a CUDA device every compute
 *    entry point returns VSL_ERR_CUDA.
 */
#ifndef VSL_H_
#define VSL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSL_ABI_VERSION 3
#define VSL_MAX_SCALES 4
#define VSL_MAX_SRC 4

typedef enum VslStatus {
  VSL_OK = 0,
  VSL_ERR_BAD_DESC = -1,      /* null / wrong abi_version / sizes out of range            */
  VSL_ERR_NULL_POINTER = -2,  /* a required buffer pointer is null                        */
  VSL_ERR_MISALIGNED = -3,    /* a buffer is not aligned to its element size              */
  VSL_ERR_UNSUPPORTED = -4,   /* flag / dtype combination not implemented by this build   */
  VSL_ERR_WORKSPACE = -5,     /* workspace smaller than vsl_loss_workspace_bytes()        */
  VSL_ERR_CUDA = -6           /* a CUDA runtime call failed (see vsl_last_cuda_error)     */
} VslStatus;

/* option bits, named after the reference flags (options.py:145-159) */
enum {
  VSL_FLAG_AUTOMASK = 1 << 0,         /* NOT --disable_automasking (trainer.py:620-633, 654-661) */
  VSL_FLAG_AVG_REPROJECTION = 1 << 1, /* --avg_reprojection (trainer.py:629-630, 649-650)        */
  VSL_FLAG_NO_SSIM = 1 << 2,          /* --no_ssim (trainer.py:549-550)                          */
  VSL_FLAG_V1_MULTISCALE = 1 << 3,    /* --v1_multiscale (trainer.py:497-498, 604-605)           */
  VSL_FLAG_FORWARD_ONLY = 1 << 4      /* losses, masks and side outputs only: Trainer.val() runs the path under
                                         torch.no_grad() (trainer.py:463-489).  The gradient buffers
                                         (grad_disp_*, grad_P, smooth_norm, grad_predictive_mask) may be null
                                         and are not written; vsl_loss_combine_grads must not follow.     */
};

enum { VSL_DTYPE_F32 = 0, VSL_DTYPE_BF16 = 1 };

/* arithmetic-order selectors (VslDesc.arith).  0 reproduces eager PyTorch-CUDA bit for bit
 * (the order of the reference on a GPU).  VSL_ARITH_TRUE_DIV reproduces PyTorch-CPU's
 * `x /= (W-1)` (a true division; CUDA multiplies by the rounded reciprocal).             */
enum {
  VSL_ARITH_TRUE_DIV = 1 << 0,
  VSL_ARITH_DOT_NOFMA = 1 << 1,    /* K=4 bmm (projection) adds un-fused products: cuBLAS, batch 1 */
  VSL_ARITH_DOT_REVERSE = 1 << 2,  /* probe: k-descending accumulation in the K=4 bmm       */
  VSL_ARITH_UPS_RIGHT = 1 << 3,    /* probe: up-sample fuses the right-hand product         */
  VSL_ARITH_UPS_NOFMA = 1 << 4,    /* probe: up-sample without FMA contraction              */
  VSL_ARITH_TAP_NOFMA = 1 << 5,    /* probe: bilinear tap accumulation without FMA          */
  VSL_ARITH_MEAN_DIV = 1 << 6,     /* probe: channel mean as sum/3 instead of sum*(1/3)     */
  VSL_ARITH_DOT3_NOFMA = 1 << 7,   /* K=3 bmm (rays) adds un-fused products: cuBLAS, batch 1 */
  VSL_ARITH_DOT3_REVERSE = 1 << 8, /* probe: k-descending accumulation in the K=3 bmm       */
  VSL_ARITH_DOTKT_NOFMA = 1 << 9,  /* K@T (4x4x4 bmm, layers.py:254) adds un-fused products  */
  VSL_ARITH_DOTKT_REVERSE = 1 << 10, /* probe: k-descending accumulation in K@T              */
  VSL_ARITH_NORM_SEQ = 1 << 11      /* torch.norm of the axis-angle as (x0^2+x1^2)+x2^2 instead of the
                                       4-lane shuffle tree (x0^2+x2^2)+x1^2 (vsl_pose_forward)        */
};

/* Problem descriptor: what Trainer.__init__ fixes once (trainer.py:245-259, options.py). */
typedef struct VslDesc {
  int32_t abi_version;            /* VSL_ABI_VERSION                                        */
  int32_t batch;                  /* opt.batch_size (local batch of this rank)              */
  int32_t height, width;          /* opt.height, opt.width (scale 0)                        */
  int32_t num_scales;             /* len(opt.scales), 1..VSL_MAX_SCALES                     */
  int32_t scale_ids[VSL_MAX_SCALES]; /* opt.scales: level s has size (H >> s, W >> s)       */
  int32_t num_src;                /* len(opt.frame_ids) - 1, 1..VSL_MAX_SRC                 */
  int32_t flags;                  /* VSL_FLAG_*                                             */
  int32_t image_dtype;            /* VSL_DTYPE_* of colour images                           */
  int32_t arith;                  /* VSL_ARITH_*; 0 = PyTorch-CUDA order                    */
  float min_disp;                 /* float32(1/opt.max_depth)           (layers.py:90)      */
  float disp_range;               /* float32(1/min_depth - 1/max_depth) (layers.py:92)      */
  float eps;                      /* Project3D eps, 1e-7                (layers.py:245)     */
  float smooth_weight;            /* opt.disparity_smoothness           (trainer.py:680)    */
  int32_t smooth_level_bias;      /* smoothness is divided by 2^(scale_id + bias) (trainer.py:680); non-zero
                                     only for --v1_multiscale, where each level runs as its own problem  */
} VslDesc;

/* ------------------------------------------------------------------------------------ */
int vsl_abi_version(void);
const char* vsl_status_string(int status);
/* cudaError_t of the last failing CUDA call made by this thread's previous vsl_* call. */
int vsl_last_cuda_error(void);

/* ------------------------------------------------------------------------------------
 * Fused loss, forward + backward in one pass.
 * Replaces Trainer.generate_images_pred + Trainer.compute_losses + the autograd backward
 * of that graph (trainer.py:491-541, 557-686, 312) for the default flags (automasking,
 * per-pixel minimum, SSIM, multi-scale at full resolution), any mix of temporal and
 * stereo source frames.
 * ------------------------------------------------------------------------------------ */
typedef struct VslLossBuffers {
  /* inputs */
  const void* target[VSL_MAX_SCALES];   /* inputs[("color",0,s)]  [B,3,H>>s,W>>s]; [0] is the photometric target */
  const void* source[VSL_MAX_SRC];      /* inputs[("color",f,0)]  [B,3,H,W] per source frame          */
  const float* disp[VSL_MAX_SCALES];    /* outputs[("disp",s)]    [B,1,H>>s,W>>s]                     */
  const float* inv_K;                   /* inputs[("inv_K",0)]    [B,4,4]                             */
  const float* P[VSL_MAX_SRC];          /* (K @ T_f)[:, :3, :]    [B,3,4]  (layers.py:254); may be null
                                           for a frame whose T is given                               */
  const float* K;                       /* inputs[("K",0)]        [B,4,4]  (needed with T)            */
  const float* T[VSL_MAX_SRC];          /* outputs[("cam_T_cam",0,f)] / inputs["stereo_T"] [B,4,4]: the
                                           kernel then forms P_f = (K @ T_f)[:3,:] itself             */
  const float* T_scale[VSL_MAX_SCALES][VSL_MAX_SRC]; /* optional per-(scale, frame) override of T: posecnn
                                           rescales the translation by each level's mean inverse depth
                                           (trainer.py:516-525); null entries fall back to T[f]        */
  const float* noise[VSL_MAX_SCALES];   /* torch.randn draw of trainer.py:656 per scale [B,F,H,W]
                                           ([B,1,H,W] with VSL_FLAG_AVG_REPROJECTION); unused without automask */
  const float* predictive_mask[VSL_MAX_SCALES]; /* optional, --predictive_mask with --disable_automasking
                                           (trainer.py:635-642): the mask at the warp resolution [B,F,H,W]  */
  /* outputs */
  float* losses;                        /* [3*S+1]: min_loss/s (S), loss/s (S), loss, smooth/s (S)    */
  float* mask[VSL_MAX_SCALES];          /* outputs["identity_selection/s"] [B,H,W]; may be null       */
  float* grad_disp_photo[VSL_MAX_SCALES];  /* d(min_loss/s)/d disp_s   [B,1,H>>s,W>>s]                */
  float* grad_disp_smooth[VSL_MAX_SCALES]; /* d(smooth_s)/d(norm disp_s) [B,1,H>>s,W>>s]; combine_grads applies
                                              the chain through the per-image mean with smooth_norm        */
  float* smooth_norm;                   /* [S][B][2] per-image normalisation terms for the backward          */
  float* grad_P;                        /* d(min_loss/s)/d P_f  [S][F][B][12]                         */
  float* grad_predictive_mask[VSL_MAX_SCALES]; /* d(min_loss/s)/d predictive_mask[s] [B,F,H,W] (with it)    */
  /* optional side outputs of Trainer.generate_images_pred (trainer.py:506, :532-537), written by the same
   * kernel that forms them anyway; every pointer may be null.  fp32 whatever the image storage.            */
  float* side_depth[VSL_MAX_SCALES];                /* outputs[("depth",0,s)]   [B,1,H,W]                    */
  float* side_sample[VSL_MAX_SCALES][VSL_MAX_SRC];  /* outputs[("sample",f,s)]  [B,H,W,2]                    */
  float* side_color[VSL_MAX_SCALES][VSL_MAX_SRC];   /* outputs[("color",f,s)]   [B,3,H,W]                    */
  /* optional: the arg-min CHANNEL of trainer.py:663-666 per pixel and scale [B,H,W] (the mask above only says
   * whether a warped frame won).  0..F-1: identity candidate of source frame f; F..2F-1: warped frame f-F; with
   * VSL_FLAG_AVG_REPROJECTION 0 = identity mean, 1 = warped mean.  Needed for the source-image gradient
   * (vsl_source_grad_upstream); may be null.                                                                */
  uint8_t* winner[VSL_MAX_SCALES];
} VslLossBuffers;

size_t vsl_loss_workspace_bytes(const VslDesc* desc);
/* Once per workspace, before its first use (and after any aborted call): zeroes it.  The per-call completion
 * counter inside is reset by the call itself, so the hot path carries no memset. */
int vsl_loss_workspace_init(const VslDesc* desc, void* workspace, size_t workspace_bytes, void* stream);
int vsl_loss_forward_backward(const VslDesc* desc, const VslLossBuffers* buf,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Same call, additionally recording two caller-owned CUDA events (cudaEvent_t as void*, either may
 * be null) on `stream` immediately before and after the photometric kernel, so a benchmark can time
 * the dominant kernel inside its own timed region.  Events come from vsl_event_create. */
int vsl_loss_forward_backward_timed(const VslDesc* desc, const VslLossBuffers* buf,
                                    void* workspace, size_t workspace_bytes, void* stream,
                                    void* event_before, void* event_after);
int vsl_event_create(void** event);
int vsl_event_destroy(void* event);
/* waits for `stop` to complete, then returns the time between the two records in milliseconds */
int vsl_event_elapsed_ms(void* start, void* stop, float* ms);

/* Chain rule from the loss dict to the leaves: `upstream` holds dL/d(losses[k]) on the DEVICE in
 * the order min_loss/0..S-1, loss/0..S-1, loss (2S+1 floats; trainer.py:672-685 defines how the
 * entries depend on each other).  Writes grad_disp[s] [B,1,H>>s,W>>s] and, each optional (null to skip),
 * grad_P_out [F][B][12] = dL/dP_f and grad_T_out [F][B][16] = K[:3,:]^T dL/dP_f (needs buf->K).  When
 * buf->T_scale is used the poses differ per scale: grad_T_out is then [S][F][B][16] and grad_P_out must be null. */
int vsl_loss_combine_grads(const VslDesc* desc, const float* upstream,
                           const VslLossBuffers* buf, float* const grad_disp[VSL_MAX_SCALES],
                           float* grad_P_out, float* grad_T_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Side outputs of Trainer.generate_images_pred (trainer.py:491-541) for one scale:
 * outputs[("depth",0,s)] [B,1,H,W], outputs[("sample",f,s)] [B,H,W,2], outputs[("color",f,s)]
 * [B,3,H,W].  Any output pointer may be null.  Only read by logging in the reference
 * (wandb_logging.py:134-143), so the fused loss never materialises them.
 * ------------------------------------------------------------------------------------ */
int vsl_warp_forward(const VslDesc* desc, int scale_index, const float* disp, const float* inv_K,
                     const float* const P[VSL_MAX_SRC], const void* const source[VSL_MAX_SRC],
                     float* depth, float* const sample[VSL_MAX_SRC], float* const color[VSL_MAX_SRC],
                     void* stream);

/* Calibration helper: out[b,i,n] = sum_k A[b,i,k] * X[b,k,n] (A [B,3,k], X [B,k,n], k = 3 or 4) with
 * the accumulation order `arith` selects, so the host layer can find which order torch.bmm (cuBLAS)
 * uses for a shape on this device.  Not used on the hot path. */
int vsl_probe_bmm(int batch, int k, int n, int arith, const float* A, const float* X, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Stand-alone layers (the layers.py call surface).  fp32 only.
 * ------------------------------------------------------------------------------------ */
/* transformation_from_parameters (layers.py:97-114, with rot_from_axisangle :133-172 and
 * get_translation_matrix :117-130): axisangle [B,3], translation [B,3] -> T [B,4,4]; invert != 0 gives
 * R^T T(-t) (used for frames before the target, trainer.py:437-438). */
int vsl_pose_forward(int batch, int invert, int arith, const float* axisangle, const float* translation,
                     float* T, void* stream);
/* its backward: grad_T [B,4,4] -> grad_axisangle [B,3], grad_translation [B,3] */
int vsl_pose_backward(int batch, int invert, const float* axisangle, const float* translation,
                      const float* grad_T, float* grad_axisangle, float* grad_translation, void* stream);
/* The posecnn pose tail (--pose_model_type posecnn, trainer.py:516-525) for every scale and frame at once:
 *   T[s][f] = transformation_from_parameters(axisangle_f, translation_f * mean_inv_depth_s, invert_f)   [S][F][B][16]
 *   mean_inv_depth_s[b] = mean over the H x W pixels of 1 / depth_s (depth_s = disp_to_depth of the up-sampled disp_s)
 * desc gives batch, sizes, scales and the depth range; axisangle / translation [B,3] per frame; invert [F] on the HOST.
 * The per-pixel values are the reference's bits, the mean is accumulated in fp64 in a fixed order (the reference takes
 * two fp32 torch means in a row), so T matches the torch ops to ~1e-7 relative rather than bit for bit.
 * workspace: vsl_posecnn_workspace_bytes, 8-byte aligned.  mean_inv [S][B] is kept for the backward. */
size_t vsl_posecnn_workspace_bytes(const VslDesc* desc);
int vsl_posecnn_forward(const VslDesc* desc, const float* const disp[VSL_MAX_SCALES], int num_frames,
                        const float* const axisangle[VSL_MAX_SRC], const float* const translation[VSL_MAX_SRC],
                        const int32_t* invert, int pose_arith, float* T, float* mean_inv, void* workspace,
                        size_t workspace_bytes, void* stream);
/* its backward: grad_T [S][F][B][16] -> grad_axisangle / grad_translation [B,3] per frame and grad_disp_const [S][B],
 * the value every pixel of d L / d disp_s[b] receives through the mean (the bilinear up-sample preserves the mean, so
 * d mean_inv_depth_s / d disp_s[j] = disp_range / (hs ws) for every j). */
int vsl_posecnn_backward(const VslDesc* desc, int num_frames, const float* const axisangle[VSL_MAX_SRC],
                         const float* const translation[VSL_MAX_SRC], const int32_t* invert, const float* mean_inv,
                         const float* grad_T, float* const grad_axisangle[VSL_MAX_SRC],
                         float* const grad_translation[VSL_MAX_SRC], float* grad_disp_const, void* stream);
/* BackprojectDepth.forward (layers.py:234-239): depth [B,1,h,w], inv_K [B,4,4] -> cam [B,4,hw] */
int vsl_backproject_forward(int batch, int height, int width, int arith, const float* depth,
                            const float* inv_K, float* cam_points, void* stream);
/* its backward: grad_cam [B,4,hw] -> grad_depth [B,1,h,w] */
int vsl_backproject_backward(int batch, int height, int width, const float* grad_cam,
                             const float* inv_K, float* grad_depth, void* stream);
/* Project3D.forward (layers.py:253-264) with P = (K@T)[:, :3, :] [B,3,4]: points [B,4,hw] -> pix [B,h,w,2] */
int vsl_project_forward(int batch, int height, int width, float eps, int arith, const float* points,
                        const float* P, float* pix, void* stream);
/* its backward: grad_pix [B,h,w,2] -> grad_points [B,4,hw] and grad_P [B,3,4] (ws: vsl_project_workspace_bytes) */
size_t vsl_project_workspace_bytes(int batch, int height, int width);
int vsl_project_backward(int batch, int height, int width, float eps, const float* points, const float* P,
                         const float* grad_pix, float* grad_points, float* grad_P,
                         void* workspace, size_t workspace_bytes, void* stream);
/* SSIM.forward (layers.py:318-332): x, y [B,C,H,W] -> [B,C,H,W] */
int vsl_ssim_forward(int batch, int channels, int height, int width, const float* x, const float* y,
                     float* out, void* stream);
/* its backward; grad_x / grad_y may be null; workspace: vsl_ssim_workspace_bytes (4 floats per element) */
size_t vsl_ssim_workspace_bytes(int batch, int channels, int height, int width);
int vsl_ssim_backward(int batch, int channels, int height, int width, const float* x, const float* y,
                      const float* grad_out, float* grad_x, float* grad_y,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Trainer.compute_reprojection_loss (trainer.py:543-555): pred, target [B,3,H,W] -> [B,1,H,W] */
int vsl_reprojection_loss_forward(int batch, int height, int width, int no_ssim, int arith,
                                  const float* pred, const float* target, float* out, void* stream);
/* workspace: vsl_ssim_workspace_bytes(batch, 3, height, width); not needed with no_ssim */
int vsl_reprojection_loss_backward(int batch, int height, int width, int no_ssim, const float* pred,
                                   const float* target, const float* grad_out, float* grad_pred,
                                   float* grad_target, void* workspace, size_t workspace_bytes, void* stream);
/* get_smooth_loss (layers.py:286-299): disp [B,1,h,w], img [B,3,h,w] -> scalar; ws: vsl_smooth_workspace_bytes */
size_t vsl_smooth_workspace_bytes(int batch, int height, int width);
int vsl_smooth_loss_forward(int batch, int height, int width, const float* disp, const float* img,
                            float* loss, void* workspace, size_t workspace_bytes, void* stream);
/* grad_loss: device scalar */
int vsl_smooth_loss_backward(int batch, int height, int width, const float* disp, const float* img,
                             const float* grad_loss, float* grad_disp, void* stream);

/* ------------------------------------------------------------------------------------
 * Gradient with respect to the SOURCE IMAGES (optional; the reference's autograd produces it whenever
 * inputs[("color", f, 0)] requires grad: through F.grid_sample's backward for the warped candidates,
 * trainer.py:534-537, and through the identity reprojection losses, trainer.py:620-633).  Composed by the
 * host layer from: the `winner` maps of the fused call, vsl_source_grad_upstream (what each candidate's
 * reprojection loss receives from the loss dict), vsl_reprojection_loss_backward (d loss / d pred), and
 * vsl_grid_sample_backward_source (the bilinear scatter into the source image).
 * ------------------------------------------------------------------------------------ */
/* upstream [2S+1] (device; same vector as vsl_loss_combine_grads).  Writes, per pixel,
 *   up_warped[s][f][B,H,W] = a_s / (B H W) where warped frame f won at scale s, else 0
 *   up_identity[f][B,H,W]  = sum_s a_s / (B H W) where identity candidate f won at scale s   (null without automask)
 * with a_s = upstream[s] + upstream[S+s] + upstream[2S]/S; with VSL_FLAG_AVG_REPROJECTION every frame receives
 * 1/F of its candidate's weight. */
int vsl_source_grad_upstream(const VslDesc* desc, const float* upstream, const uint8_t* const winner[VSL_MAX_SCALES],
                             float* up_identity, float* const up_warped[VSL_MAX_SCALES], void* stream);
/* Adjoint of F.grid_sample(source, grid, padding_mode="border", align_corners=True) with respect to `source`:
 * grad_source [B,3,H,W] += scatter of grad_pred [B,3,H,W] through the four bilinear taps of grid [B,H,W,2]
 * (ATen grid_sampler_2d_backward's safe_add_2d).  A CTA accumulates the taps that land near its tile in shared
 * memory and flushes them with one red.global.add.f32 per touched source pixel; far taps go to global memory
 * directly.  Float atomics: the summation order is not fixed (neither is PyTorch's). */
int vsl_grid_sample_backward_source(int batch, int height, int width, const float* grid, const float* grad_pred,
                                    float* grad_source, void* stream);

/* ------------------------------------------------------------------------------------
 * Depth metrics and the GAN prior's loss: the losses / metrics next to the path.
 * One caller-owned workspace of vsl_metrics_workspace_bytes() bytes (8-byte aligned) serves every call below;
 * results stay on the device.  Reductions run in a fixed order (reproducible); medians are exact.
 * ------------------------------------------------------------------------------------ */
size_t vsl_metrics_workspace_bytes(void);
/* compute_depth_errors (layers.py:335-353): gt, pred [n], all positive ->
 * out7 = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 */
int vsl_depth_errors(size_t n, const float* gt, const float* pred, float* out7, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Trainer.compute_depth_losses (trainer.py:688-716): depth_pred = outputs[("depth",0,0)] [B,1,h,w] is up-sampled
 * (bilinear, align_corners=False) to depth_gt's size [B,1,gt_h,gt_w] and clamped to [clamp_min, clamp_max]; pixels
 * with depth_gt > 0 inside crop = {y0, y1, x0, x1} (half-open; the reference's Garg/Eigen crop is
 * {153, 371, 44, 1197} at 375 x 1242) are kept, the prediction is scaled by median(gt) / median(pred) (torch.median:
 * the lower median), clamped again, and the seven metrics of compute_depth_errors are written to out7. */
int vsl_depth_losses(int batch, int height, int width, int gt_height, int gt_width, const int crop[4],
                     float clamp_min, float clamp_max, const float* depth_pred, const float* depth_gt, float* out7,
                     void* workspace, size_t workspace_bytes, void* stream);
/* SLlog.forward (layers.py:32-56; used as self.si_loss(fake_disp_scaled, disp), trainer.py:579): fake, real [n]
 * -> loss (device scalar); stats3 = (N, mean of the log differences, loss) is kept for the backward. */
int vsl_sllog_forward(size_t n, const float* fake, const float* real, float* loss, float* stats3, void* workspace,
                      size_t workspace_bytes, void* stream);
/* its backward: grad_loss (device scalar) -> grad_fake, grad_real [n] (either may be null) */
int vsl_sllog_backward(size_t n, const float* fake, const float* real, const float* stats3, const float* grad_loss,
                       float* grad_fake, float* grad_real, void* stream);

/* ------------------------------------------------------------------------------------
 * Input pipeline: 8-bit frames -> the ("color", f, s) pyramid.
 * Replaces MonoDataset.preprocess (datasets/mono_dataset2.py:103-124): level s =
 * transforms.Resize((H // 2^s, W // 2^s), Image.ANTIALIAS)(level s-1) on 8-bit PIL images
 * (:85-89), then transforms.ToTensor() (:113), followed by the fp32 host->device copy of
 * trainer.py:373-374.  The host ships the level-0 frames as uint8 HWC (what np.asarray(PIL
 * image) gives, a quarter of the fp32 bytes and no pyramid); the GPU reproduces Pillow's 8-bit
 * LANCZOS resampling (Resample.c, horizontal pass then vertical pass, each rounded to 8 bits,
 * 22-bit fixed-point coefficients) and the /255 conversion bit for bit.
 * ------------------------------------------------------------------------------------ */
typedef struct VslPyramidDesc {
  int32_t abi_version;  /* VSL_ABI_VERSION                                                   */
  int32_t batch;        /* frames per call                                                   */
  int32_t height, width;/* level 0; multiples of 2^(num_levels-1)                            */
  int32_t num_levels;   /* 1..VSL_MAX_SCALES (opt.scales = range(num_levels))                */
  int32_t out_dtype;    /* VSL_DTYPE_F32 or VSL_DTYPE_BF16 (image storage of the loss path)  */
} VslPyramidDesc;
/* workspace: coefficient tables + the 8-bit levels s >= 1; 256-byte aligned, caller-owned */
size_t vsl_pyramid_workspace_bytes(const VslPyramidDesc* desc);
/* once per (shape, workspace): evaluates Pillow's precompute_coeffs / normalize_coeffs_8bpc on the
 * host in double precision and copies the tables into the workspace (enqueued on `stream`) */
int vsl_pyramid_plan(const VslPyramidDesc* desc, void* workspace, size_t workspace_bytes, void* stream);
/* frames_hwc: [B,H,W,3] uint8 (device).  levels[s]: [B,3,H>>s,W>>s] out_dtype, or null to skip that
 * tensor (the 8-bit level is still formed, later levels need it).  levels_u8 (optional, may be null):
 * levels_u8[s] for s >= 1 receives the 8-bit level [B,H>>s,W>>s,3]. */
int vsl_pyramid_forward(const VslPyramidDesc* desc, const uint8_t* frames_hwc, void* const levels[VSL_MAX_SCALES],
                        uint8_t* const levels_u8[VSL_MAX_SCALES], void* workspace, size_t workspace_bytes,
                        void* stream);
/* The same with the horizontal flip of MonoDataset.get_color (`color.transpose(Image.FLIP_LEFT_RIGHT)` when the
 * item's do_flip is set, datasets/mono_dataset2.py:151-156 via kitti_dataset.py) applied to the raw frame while it is
 * read: flip [B] bytes on the device, non-zero = mirror that image; null = no flips.  Flipping first and resampling
 * afterwards is what the reference does (its fixed-point LANCZOS tables are not exactly mirror-symmetric). */
int vsl_pyramid_forward_flip(const VslPyramidDesc* desc, const uint8_t* frames_hwc, const uint8_t* flip,
                             void* const levels[VSL_MAX_SCALES], uint8_t* const levels_u8[VSL_MAX_SCALES],
                             void* workspace, size_t workspace_bytes, void* stream);
/* inputs["stereo_T"] of MonoDataset.__getitem__ (datasets/mono_dataset2.py:197-203) for a batch: identity with
 * T[0,3] = side_sign * baseline_sign * baseline; side_left [B] / flip [B] bytes on the device (null = all zero). */
int vsl_stereo_transform(int batch, const uint8_t* flip, const uint8_t* side_left, float baseline, float* T, void* stream);
/* host-only helper (no device work): the coefficient table of one axis, bounds [out_size][2] = (first
 * input index, tap count), coefs [out_size][13]; ksize_capacity must be 13.  Lets tests compare the
 * tables with Pillow's without a GPU. */
int vsl_pyramid_coefficients(int in_size, int out_size, int32_t* bounds, int32_t* coefs, int ksize_capacity);

/* The decoded file image -> level 0: `self.resize[0](inputs[(n, im, -1)])` (datasets/mono_dataset2.py:85-89, :107-109),
 * PIL Image.resize((out_width, out_height), LANCZOS) from any native resolution; byte-exact (the same two 8-bit
 * passes as the pyramid levels with per-size tap counts; a pass whose size does not change is skipped, like Pillow's).
 * workspace: coefficient tables + the horizontal pass's output; 256-byte aligned, caller-owned.  vsl_resize_plan once
 * per (shape, workspace), then vsl_resize_forward: frames_hwc [B,in_h,in_w,3] uint8 -> out_hwc [B,out_h,out_w,3] uint8
 * (device), which is what vsl_pyramid_forward takes. */
size_t vsl_resize_workspace_bytes(int batch, int in_height, int in_width, int out_height, int out_width);
int vsl_resize_plan(int batch, int in_height, int in_width, int out_height, int out_width, void* workspace,
                    size_t workspace_bytes, void* stream);
int vsl_resize_forward(int batch, int in_height, int in_width, int out_height, int out_width, const uint8_t* frames_hwc,
                       uint8_t* out_hwc, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Colour augmentation of the `color_aug` inputs (dataset-side remainder).
 * Replaces `self.to_tensor(color_aug(f))` (datasets/mono_dataset2.py:124) with color_aug =
 * transforms.Compose([ColorJitter(brightness, contrast, saturation, hue), RandomHorizontalFlip(0.5),
 * RandomAutocontrast()]) (datasets/mono_dataset2.py:92-97) on 8-bit PIL images: torchvision's
 * _functional_pil.py on top of Pillow's Image.blend / convert("L"|"HSV"|"RGB") / ImageOps.autocontrast,
 * reproduced byte for byte.  The random draws stay with the caller (torch's global generator); one record
 * per image describes the outcome of one call of the transform.
 * ------------------------------------------------------------------------------------ */
typedef struct VslAugParams {
  int32_t order[4];     /* ColorJitter.get_params' randperm(4): 0 brightness, 1 contrast, 2 saturation, 3 hue   */
  float factor[3];      /* brightness, contrast, saturation factors as C floats (Image.blend's alpha)           */
  int32_t hue_shift;    /* uint8(int32(hue_factor * 255)): what adjust_hue adds to the H channel (8-bit wrap)   */
  int32_t flip;         /* RandomHorizontalFlip's outcome                                                       */
  int32_t autocontrast; /* RandomAutocontrast's outcome                                                         */
  int32_t enabled;      /* 0: the item's do_color_aug is false (datasets/mono_dataset2.py:179-186): ToTensor only */
  int32_t reserved;
} VslAugParams;
/* workspace: per-image statistics + the jittered 8-bit images; 256-byte aligned, caller-owned */
size_t vsl_color_aug_workspace_bytes(int batch, int height, int width);
/* frames_hwc [B,H,W,3] uint8 and params [B] on the device.  out: [B,3,H,W] of out_dtype (VSL_DTYPE_F32 /
 * VSL_DTYPE_BF16) or null; out_u8: the augmented 8-bit images [B,H,W,3] or null (at least one of the two). */
int vsl_color_aug_forward(int batch, int height, int width, int out_dtype, const uint8_t* frames_hwc,
                          const VslAugParams* params, void* out, uint8_t* out_u8, void* workspace,
                          size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSL_H_ */
