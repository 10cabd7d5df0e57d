/*
 * polcue.h -- C ABI of libpolcue.so: the per-pixel polarization hot path on B200 (sm_100a).
 *
 * The reference (kkaytekin/Supervised-Depth-Estimation-from-Polarized-Images) is pure Python and has
 * no FFI; its boundary for this path is a set of Python function signatures.  Every
 * entry point below names the reference function (file:line, relative to the reference root) whose
 * arithmetic it replaces.  The Python mirror of those signatures lives in
 * supervised-depth-estimation-from-polarized-images_b200/polcue/ and binds this header with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - All image pointers are DEVICE pointers owned by the caller unless the function name ends in
 *     `_host`; tensors are dense, row-major, batch outermost.  Nothing is allocated per call.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and never synchronise.
 *   - Return value: 0 on success; a NEGATIVE errno-style code for a rejected argument (nothing was
 *     launched); a POSITIVE value is the cudaError_t reported by the launch.
 *   - Re-entrant for distinct streams/outputs; no process-global state changes numerics.  The `_host` entry points
 *     cache their device rings per (device, geometry): calls for different devices or geometries run concurrently,
 *     two threads using the same (device, geometry) serialise on that ring.  Not for use from forked DataLoader workers.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef POLCUE_H_
#define POLCUE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define POLCUE_API __attribute__((visibility("default")))
#else
#define POLCUE_API
#endif

#define POLCUE_OK 0
#define POLCUE_EINVAL (-22)  /* null pointer, non-positive or odd size, misaligned buffer */
#define POLCUE_ENOMEM (-12)  /* device / pinned allocation failed */
#define POLCUE_ERANGE (-34)  /* refractive index whose tables cannot be represented */
#define POLCUE_E2BIG (-7)    /* problem exceeds the 32-bit work-item range of one launch */

typedef void* polcue_stream_t;      /* cudaStream_t */
typedef struct polcue_lut polcue_lut; /* opaque: zenith-angle lookup tables for one refractive index */

POLCUE_API const char* polcue_version(void);
POLCUE_API const char* polcue_error_string(int code);
/* How many depth->normals launches took the TMA-staged kernel (the rest used the manually staged one). */
POLCUE_API unsigned long long polcue_debug_stencil_tma_launches(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
POLCUE_API unsigned long long polcue_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Zenith-angle tables.  Replaces the table construction + scipy interp1d(linear, extrapolate) of
 *   manydepth/normals_vec.py:11-22 (rho_diffuse) and :25-50 (rho_spec),
 *   polarisation/xolp_and_normals.py:41-83, ppp_code/physical_normals_channels.py:39-72.
 * The three 1000-knot piecewise-linear tables are built on the host in float64 and re-indexed into
 * uniform cells of g(rho) (one knot per cell) so the device needs no search; see DESIGN.md.
 * `polcue_lut_create` allocates on the CURRENT device.  Returns POLCUE_ERANGE if no cell grid
 * reproduces the reference interpolant to 1e-6 rad inside its shared-memory budget.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_lut_create(double n, polcue_lut** out);
POLCUE_API void polcue_lut_destroy(polcue_lut* lut);
/* Zenith-angle sincos used by every later launch that is given THIS handle: 1 = MUFU sin/cos (3.6e-7 abs, the default of
 * a new handle), 0 = polynomial (1.4e-7 abs).  A property of the handle, not of the process: callers that want both keep
 * two handles.  Not to be changed while a call using the handle is being issued from another thread. */
POLCUE_API int polcue_lut_set_trig(polcue_lut* lut, int mufu);
/* Host-side introspection (no GPU needed): cell counts / knot counts of table t (0 diffuse, 1 spec1, 2 spec2). */
POLCUE_API int polcue_lut_host_build(double n, polcue_lut** out); /* same tables, no device copy (CPU tests) */
POLCUE_API int polcue_lut_cells(const polcue_lut* lut, int table);
POLCUE_API int polcue_lut_knots(const polcue_lut* lut, int table, double* x, double* y, int capacity);
/* Evaluates the float32 cell tables on the host exactly as the kernels do (float32 arithmetic). */
POLCUE_API int polcue_lut_eval_host(const polcue_lut* lut, int table, const float* rho, size_t count, float* theta);
/* Steep end segment of table t: where the last segment of the reference's table is steeper than 512 rad per unit rho
 * (second specular branch at n = 1.8: slope -10797, theta -> -1e4 rad for rho -> 2) float32 holds neither rho nor
 * theta to the parity bound, so every kernel evaluates queries beyond the second-to-last knot in float64.
 * Returns 1 and fills xys = {x_lo, y_lo, slope} if table t has one, 0 if not. */
POLCUE_API int polcue_lut_steep(const polcue_lut* lut, int table, double* xys);

/* ---------------------------------------------------------------------------------------------
 * Quadrant split ("demosaic" of the stored 2x2-tiled polarizer image), bit-exact.
 *   polarisation/pol_split_and_save.py:10-27  split_pol(img) -> (im00, im10, im01, im11)
 * img: B x H x W pixels of px_bytes bytes each; H and W even.  Each output: B x H/2 x W/2.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_split_pol(const void* img, int B, int H, int W, int px_bytes,
                     void* im00, void* im10, void* im01, void* im11, polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused pipeline: quadrant split -> Stokes -> Iun/DoLP/AoLP -> three physics normal candidates.
 * One launch replaces the chain of polarisation/xolp_and_normals.py:107-121:
 *   split_pol (pol_split_and_save.py:10-27), np.stack order (indoor_dataset.py:435-439),
 *   Iun_and_xolp (polarisation/xolp.py:8-34, canonical angles, closed form of the 4x3 least squares),
 *   rho_diffuse / rho_spec (normals_vec.py:11-50), calc_normals x3 (normals_vec.py:53-60) in the
 *   channel order of ShallowNormalsEncoder.get_normals (networks/pre_encoders.py:99-113).
 * mosaic : B x H x W uint8, quadrants TL/TR/BL/BR = 0/45/90/135 degrees.  Hs = H/2, Ws = W/2.
 * planes : B x 4 x Hs x Ws uint8 (I0, I45, I90, I135) or NULL
 * iun    : B x Hs x Ws float32 or NULL
 * xolp   : B x 2 x Hs x Ws float32 (DoLP, AoLP)                      [required]
 * normals: B x 9 x Hs x Ws float32 (diffuse xyz, spec1 xyz, spec2 xyz) or NULL
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_fused_mosaic_u8(const uint8_t* mosaic, int B, int H, int W, const polcue_lut* lut,
                           uint8_t* planes, float* iun, float* xolp, float* normals,
                           polcue_stream_t stream);

/* The fused pipeline with the dataset statistics / output checksum as a BY-PRODUCT of the same launch:
 *   polarisation/xolp_mean_and_std_dev.py:26-32  mean / std of DoLP and AoLP over a set of frames (the constants of
 *   manydepth/networks/pre_encoders.py:79); per-plane float64 output checksums of the sequence benchmark.
 * stats13 (13 device doubles): sum rho, sum phi, the sums of the nine normal channels, sum rho^2, sum phi^2 over the batch.
 * The sums follow the library's canonical order (group of 4 pixels -> tile of 256 groups -> segment of 1024 tiles -> total,
 * polcue_device.cuh): bitwise reproducible under the kernel's dynamic tile scheduling, and bit-identical to
 * polcue_channel_stats_f32 of the stored xolp / normals.  The outputs are not read again (one extra launch folds the
 * 64-byte tile records).  Shapes without 4-pixel groups (Ws % 4 != 0) and tables with a steep end segment take the
 * separate statistics pass internally -- same results, two more reads of the outputs.
 * workspace: polcue_fused_stats_workspace_bytes(B, H, W) bytes, 64-byte aligned, first 8 bytes zero before the first use. */
POLCUE_API size_t polcue_fused_stats_workspace_bytes(int B, int H, int W);
POLCUE_API int polcue_fused_mosaic_stats_u8(const uint8_t* mosaic, int B, int H, int W, const polcue_lut* lut,
                                 uint8_t* planes, float* iun, float* xolp, float* normals,
                                 void* workspace, double* stats13, polcue_stream_t stream);

/* The same pipeline for the RAW sensor layout (not used by the reference, whose HAMMER images are stored pre-tiled):
 * an interleaved 2x2 super-pixel mosaic, output pixel (y, x) owning mosaic pixels (2y, 2x), (2y, 2x+1), (2y+1, 2x),
 * (2y+1, 2x+1).  angle_at: 4 HOST ints, the angle index (0, 1, 2, 3 = 0, 45, 90, 135 deg) found at those four
 * positions, each exactly once -- e.g. {2, 1, 3, 0} for a 90/45 over 135/0 polarizer array. */
POLCUE_API int polcue_fused_superpixel_u8(const uint8_t* mosaic, int B, int H, int W, const int* angle_at, const polcue_lut* lut,
                               uint8_t* planes, float* iun, float* xolp, float* normals, polcue_stream_t stream);

/* The same fused kernel fed by four separate B x H x W uint8 planes (the loader's pol00 / pol01 / pol10 / pol11 images,
 * manydepth/datasets/indoor_dataset.py:435-438) instead of a quadrant mosaic: XOLP (+ Iun) and the nine normal channels
 * at the planes' resolution in one launch -- get_xolp (:430-442) + get_normals (pre_encoders.py:99-113). */
POLCUE_API int polcue_fused_planes_u8(const uint8_t* i0, const uint8_t* i45, const uint8_t* i90, const uint8_t* i135,
                           int B, int H, int W, const polcue_lut* lut, float* iun, float* xolp, float* normals,
                           polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loader front end.  The Lanczos resize of the 8-bit polarizer images, bit-exact with Pillow:
 *   manydepth/datasets/indoor_dataset.py:77,115   Image.ANTIALIAS, transforms.Resize((height, width))
 *   manydepth/datasets/indoor_dataset.py:335-349  resize_pol(get_gray(..)) of pol00 / pol10 / pol01 / pol11
 *   manydepth/datasets/hammer_dataset.py:68-75    get_gray: 'L' image, optional FLIP_LEFT_RIGHT
 * Arithmetic: Pillow (6.2.1 pinned, environment.yml:14) src/libImaging/Resample.c -- float64 Lanczos-3 weights,
 * 22-bit fixed point, horizontal pass into an 8-bit intermediate, vertical pass.
 * A plan holds the weights of one (in_h, in_w) -> (out_h, out_w) geometry on the CURRENT device.
 * workspace: polcue_resize_workspace_bytes(plan, images) device bytes (the 8-bit intermediate, images x in_h x out_w, + 32 rows).
 * flip: NULL or one device byte per image (polcue_resize_lanczos_u8) / per sample (front end): non-zero mirrors the
 *       image left-right before the resize, as hammer_dataset.py:72-73 does.
 * ------------------------------------------------------------------------------------------- */
typedef struct polcue_resize_plan polcue_resize_plan;
POLCUE_API int polcue_resize_plan_create(int in_h, int in_w, int out_h, int out_w, polcue_resize_plan** out);
POLCUE_API int polcue_resize_plan_host_build(int in_h, int in_w, int out_h, int out_w, polcue_resize_plan** out); /* no device copy */
POLCUE_API void polcue_resize_plan_destroy(polcue_resize_plan* plan);
/* Host introspection: returns ksize of axis (0 horizontal, 1 vertical); bounds = out x (first, count), kk = out x ksize. */
POLCUE_API int polcue_resize_plan_coeffs(const polcue_resize_plan* plan, int axis, int* bounds, int* kk, size_t kk_capacity);
/* Tests only: 1 = run the horizontal pass in its byte-load form even where the dp4a form applies. */
POLCUE_API int polcue_debug_resize_force_bytes(int on);
/* Profiling aid: enable = 1 records CUDA events around the two passes of every resize call; a later call with non-null
 * pointers waits for the last resize and returns the device time of its horizontal and vertical pass. */
POLCUE_API int polcue_debug_resize_pass_times(int enable, float* ms_h, float* ms_v);
POLCUE_API size_t polcue_resize_workspace_bytes(const polcue_resize_plan* plan, int images);
/* src: images x in_h x in_w uint8 -> dst: images x out_h x out_w uint8 (two launches). */
POLCUE_API int polcue_resize_lanczos_u8(const polcue_resize_plan* plan, const uint8_t* src, int images, const uint8_t* flip,
                             uint8_t* workspace, uint8_t* dst, polcue_stream_t stream);
/* The polarization branch of __getitem__ + the encoder front end for a batch: the four full-resolution gray images
 * (each B x in_h x in_w; i0 = pol00, i45 = pol01, i90 = pol10, i135 = pol11, indoor_dataset.py:435-438) are resized
 * into planes (B x 4 x out_h x out_w, angle order, required), then get_xolp (:430-442) and get_normals
 * (pre_encoders.py:99-113) run on them as in polcue_fused_planes_u8.  workspace: 4 * B images.  Three launches.
 * xolp_norm (B x 2 x out_h x out_w, or NULL) additionally receives ShallowEncoder.normalizeInput(xolp, 'XOLP')
 * (pre_encoders.py:75-83) = (xolp - mean) / std with xolp_mean_std = 2 HOST floats {mean, std}
 * ({0.08693199701957657, 0.44430732785457433} in the reference), folded into the kernel's store. */
POLCUE_API int polcue_loader_front_end_u8(const polcue_resize_plan* plan, const uint8_t* i0, const uint8_t* i45, const uint8_t* i90,
                               const uint8_t* i135, int B, const uint8_t* flip, const polcue_lut* lut, uint8_t* workspace,
                               uint8_t* planes, float* iun, float* xolp, float* normals, const float* xolp_mean_std,
                               float* xolp_norm, polcue_stream_t stream);

/* The loader front end with HOST buffers (pinned or pageable): the four image stacks (each B x in_h x in_w) are copied
 * in by chunks of `chunk_samples` (<= 0: default), resized and reduced on the device, and planes (B x 4 x out_h x out_w,
 * or NULL), xolp, normals (or NULL) and xolp_norm (or NULL) are copied back, copies and kernels overlapping on three
 * streams.  h_flip: NULL or B host bytes.  Blocks until done.  The weights of the geometry are cached between calls. */
POLCUE_API int polcue_loader_front_end_u8_host(int in_h, int in_w, int out_h, int out_w, const uint8_t* h_i0, const uint8_t* h_i45,
                                    const uint8_t* h_i90, const uint8_t* h_i135, int B, const uint8_t* h_flip, const polcue_lut* lut,
                                    uint8_t* h_planes, float* h_xolp, float* h_normals, const float* xolp_mean_std,
                                    float* h_xolp_norm, int chunk_samples);

/* Same pipeline with HOST buffers (pinned or pageable): chunks of frames are copied in, processed
 * and copied out on three streams so H2D, the kernel and D2H overlap.  Blocks until done.
 * `chunk_frames` <= 0 picks a default.  This is the call bench.py's `e2e` figure times. */
POLCUE_API int polcue_fused_mosaic_u8_host(const uint8_t* h_mosaic, int B, int H, int W, const polcue_lut* lut,
                                float* h_iun, float* h_xolp, float* h_normals, int chunk_frames);
/* Pinned host memory for the callers of the *_host entry points: driver-allocated portable pinned memory (cudaHostAlloc),
 * zero-filled, allocated while the calling thread prefers the NUMA node of the GPU that will DMA into it when the platform
 * exposes one (/sys/bus/pci/devices/<bus id>/numa_node).  `device` < 0: the current device.
 * polcue_host_alloc = polcue_host_alloc_on(.., -1).  polcue_host_numa_node: the node found for `device`, or -1. */
POLCUE_API int polcue_host_alloc(void** ptr, size_t bytes);
POLCUE_API int polcue_host_alloc_on(void** ptr, size_t bytes, int device);
POLCUE_API int polcue_host_free(void* ptr);
POLCUE_API int polcue_host_numa_node(int device);

/* ---------------------------------------------------------------------------------------------
 * XOLP only.   polarisation/xolp.py:8-34   Iun_and_xolp(images[H,W,4], angles[4]) -> Iun, rho, phi
 * stack : B x H x W x 4 interleaved samples in angle order (the np.stack of indoor_dataset.py:439).
 * pinv  : NULL for the canonical angles (0,45,90,135 deg; closed form), else the 3x4 row-major
 *         pseudo-inverse of [1, cos 2a, sin 2a] (what np.linalg.lstsq applies), 12 HOST floats.
 * iun   : B x H x W or NULL;  xolp : B x 2 x H x W (rho, phi).
 * inf/nan scrub as xolp.py:26-29 (+inf -> 0, nan -> 0, -inf -> -FLT_MAX).
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_xolp_stack_u8(const uint8_t* stack, int B, int H, int W, const float* pinv,
                         float* iun, float* xolp, polcue_stream_t stream);
POLCUE_API int polcue_xolp_stack_f32(const float* stack, int B, int H, int W, const float* pinv,
                          float* iun, float* xolp, polcue_stream_t stream);
/* Loader path: four separate B x H x W uint8 planes (indoor_dataset.py:435-438), canonical angles. */
POLCUE_API int polcue_xolp_planes_u8(const uint8_t* i0, const uint8_t* i45, const uint8_t* i90, const uint8_t* i135,
                          int B, int H, int W, float* iun, float* xolp, polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Physics normals from XOLP.
 *   manydepth/networks/pre_encoders.py:99-113  ShallowNormalsEncoder.get_normals(x, n)
 * xolp: B x 2 x H x W float32 -> normals: B x 9 x H x W float32.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_normals_from_xolp_f32(const float* xolp, int B, int H, int W, const polcue_lut* lut,
                                 float* normals, polcue_stream_t stream);
/* The three fine-grained functions of manydepth/normals_vec.py (count = number of elements).
 *   rho_diffuse :11-22, rho_spec :25-50, calc_normals :53-60 (phi, theta: B x HW -> B x 3 x HW). */
POLCUE_API int polcue_rho_diffuse_f32(const float* rho, size_t count, const polcue_lut* lut, float* theta,
                           polcue_stream_t stream);
POLCUE_API int polcue_rho_spec_f32(const float* rho, size_t count, const polcue_lut* lut, float* theta1, float* theta2,
                        polcue_stream_t stream);
POLCUE_API int polcue_calc_normals_f32(const float* phi, const float* theta, int B, size_t hw, float* normals,
                            polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Masked direct-Stokes "channel" variant.
 *   ppp_code/physical_normals_channels.py:15-36  PolarisationImage_channel (s0 = I0 + I90, no nan scrub)
 *   ppp_code/physical_normals_channels.py:75-83  calc_normals_channel (H x W x 3, zero outside mask)
 * stack: H x W x 4 float32; mask: H x W uint8 (non-zero = inside); rho/phi/iun: H x W float32.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_stokes_channel_f32(const float* stack, const uint8_t* mask, int H, int W,
                              float* rho, float* phi, float* iun, polcue_stream_t stream);
POLCUE_API int polcue_calc_normals_channel_f32(const float* phi, const float* theta, const uint8_t* mask, size_t hw,
                                    float phi_offset, float* normals_hw3, polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Depth -> surface normals 3x3 stencil.
 *   kornia.geometry.depth.depth_to_normals (kornia 0.5.11, environment.yml:41), called at
 *   manydepth/trainer.py:1305-1306,1477,1484.
 * depth: B x 1 x H x W float32; K: B x 3 x 3 float32 (pixel units); normals: B x 3 x H x W float32.
 * Forward only (the GT branch); see DESIGN.md for the autograd note.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API int polcue_depth_to_normals_f32(const float* depth, const float* K, int B, int H, int W, float* normals,
                                polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Supervised normals loss, forward and backward (the consumer of the stencil in training).
 *   manydepth/trainer.py:1298-1309  Trainer.compute_supervised_normals_losses(depth_gt, depth_pred, intrinsics, mask)
 *     loss = sum((2 - cos(n_gt, n_pred)) * mask) / sum(mask),  n = depth_to_normals(depth, K)
 * depth_gt, depth_pred, mask: B x 1 x H x W float32; K: B x 3 x 3.
 * fwd: sums2 (device, 2 doubles) receives {sum((2-cos) m), sum(m)}; loss (device float, may be NULL) = their ratio.
 *      workspace: polcue_normals_loss_workspace_bytes() bytes (per-CTA partials; a second tiny launch folds them in a
 *      fixed order, so results are bitwise reproducible).
 * bwd: grad_pred (B x 1 x H x W) = d(loss)/d(depth_pred) * grad_out, with sums2 from the forward call and grad_out a
 *      DEVICE float scalar.  depth_gt and mask receive no gradient (they are data).
 * Both kernels evaluate the stencil in a cancellation-free form scaled by f_x f_y: |64 f_x f_y n|^2 must stay
 * finite in float32, i.e. depths up to ~1e7 in whatever unit (metres in the reference); the clamps of F.normalize and
 * cosine_similarity are reproduced for |n| < 1e-12 (zero-depth holes).
 * ------------------------------------------------------------------------------------------- */
POLCUE_API size_t polcue_normals_loss_workspace_bytes(void);
POLCUE_API int polcue_normals_loss_fwd_f32(const float* depth_gt, const float* depth_pred, const float* K, const float* mask,
                                int B, int H, int W, void* workspace, double* sums2, float* loss, polcue_stream_t stream);
POLCUE_API int polcue_normals_loss_bwd_f32(const float* depth_gt, const float* depth_pred, const float* K, const float* mask,
                                int B, int H, int W, const double* sums2, const float* grad_out, float* grad_pred,
                                polcue_stream_t stream);

/* The whole supervised block of Trainer.compute_losses for one scale (manydepth/trainer.py:1240-1251):
 *     mask = (gt >= min_d).float() * (gt <= max_d).float()
 *     supervised_depth_loss   = (|gt - pred| * mask).sum() / mask.sum()
 *     supervised_normals_loss = compute_supervised_normals_losses(gt, pred, K, mask)          (:1298-1309)
 * in the same forward kernel (the mask is derived from the staged GT tile, no mask tensor is read) and the same backward
 * kernel.  sums3 = {sum (2 - cos) m, sum m, sum |gt - pred| m} (device doubles); losses2 = {normals loss, depth loss} or
 * NULL.  Backward: grad_pred = grad_normals * d(normals loss) + grad_depth * sign(pred - gt) m / sum m (device scalars).
 * workspace: polcue_normals_loss_workspace_bytes(). */
POLCUE_API int polcue_supervised_losses_fwd_f32(const float* depth_gt, const float* depth_pred, const float* K, float min_d, float max_d,
                                     int B, int H, int W, void* workspace, double* sums3, float* losses2,
                                     polcue_stream_t stream);
POLCUE_API int polcue_supervised_losses_bwd_f32(const float* depth_gt, const float* depth_pred, const float* K, float min_d, float max_d,
                                     int B, int H, int W, const double* sums3, const float* grad_normals,
                                     const float* grad_depth, float* grad_pred, polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Depth error metrics.
 *   manydepth/layers.py:539-557 compute_depth_errors / :559-577 compute_depth_errors_numpy
 * sums8   : 8 doubles (device): count, n[t<1.25], n[t<1.25^2], n[t<1.25^3], S d^2, S dlog^2, S |d|/gt, S d^2/gt
 * metrics7: 7 floats (device) in the reference order abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3; or NULL
 * workspace: device scratch of polcue_depth_errors_workspace_bytes() bytes (contents irrelevant, but its
 *            first 8 bytes must be ZERO before the first use; the kernel re-zeroes them itself).
 * The sums are additive across ranks (NCCL all-reduce of 8 doubles).
 * ------------------------------------------------------------------------------------------- */
POLCUE_API size_t polcue_depth_errors_workspace_bytes(void);
POLCUE_API int polcue_depth_errors_f32(const float* gt, const float* pred, size_t count, void* workspace,
                            double* sums8, float* metrics7, polcue_stream_t stream);
/* Per-image masked evaluation of Trainer.compute_depth_losses_from_list (manydepth/trainer.py:1376-1428):
 * mask = gt > min_d && gt < max_d [&& inst == inst_id], pred clamped to [min_d, max_d].
 * inst: B x px uint8 instance-id map or NULL (then inst_id is ignored).
 * sums: B x 8 doubles; metrics: B x 7 floats or NULL.  One thread-block cluster per image, no workspace.
 * The cluster size (1..8 CTAs per image) follows the batch: the largest that still fits one wave on the device, fewer CTAs
 * per image for batches of several waves.  The pixel counts (sums[0..3]) are exact integers for every size; the float sums
 * follow the per-thread summation order, so they are bitwise reproducible from run to run for the same batch size, image
 * size and GPU model, and agree to float32 rounding (~1e-7 relative) between launch geometries.  The same holds for the
 * *_scaled, *_groups and eval_pass entry points below. */
POLCUE_API int polcue_depth_errors_images_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px,
                                   float min_d, float max_d, int inst_id, double* sums, float* metrics,
                                   polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Per-channel sum and sum of squares (float64) of a planar float32 tensor x[B, C, hw].
 *   polarisation/xolp_mean_and_std_dev.py:26-32 (mean / std of DoLP and AoLP over a set of frames; the constants
 *   of manydepth/networks/pre_encoders.py:79); also the output checksum of the sequence benchmark.
 * stats: C x 2 doubles (device): sum, sum of squares per channel; additive across frames and ranks.  Canonical order
 * (see polcue_fused_mosaic_stats_u8): float32 inside a tile of 1024 elements, float64 above; bitwise reproducible.
 * workspace: polcue_channel_stats_workspace_bytes(B, C, hw) bytes, first 8 bytes zero before the first use.
 * ------------------------------------------------------------------------------------------- */
POLCUE_API size_t polcue_channel_stats_workspace_bytes(int B, int C, size_t hw);
POLCUE_API int polcue_channel_stats_f32(const float* x, int B, int C, size_t hw, void* workspace, double* stats,
                             polcue_stream_t stream);

/* The evaluation loops in full (manydepth/trainer.py:1356-1430, manydepth/evaluation.py:215-288), one mask group per call:
 *   mask      = (gt > min_d) & (gt < max_d) [& inst_lo <= inst <= inst_hi]      (inst NULL: no material filter;
 *               evaluation.py's "objects" group is the id RANGE 20..160, every other group has inst_lo == inst_hi)
 *   clamp_first != 0: pred is clamped to [min_d, max_d] first, as both loops clamp the whole batch (trainer.py:1368-1370)
 *   median scaling (trainer.py:1413-1414, the configurations without depth supervision):
 *               pred *= np.median(gt[mask]) / np.median(pred[mask])
 *   pred clamped to [min_d, max_d], then the seven metrics per image.
 * polcue_masked_median_scale_f32: medians[B][2] = (median of gt[mask], median of pred[mask]) exactly as np.median gives them
 * for float32 arrays (radix select; an even count averages the two middle values in float32) and scale[B] (or NULL) = their
 * float32 ratio; an empty mask gives NaN.
 * polcue_depth_errors_images_scaled_f32: the per-image metrics with pred multiplied by pred_scale[image] (device floats,
 * NULL = no scaling) between the two clamps. */
POLCUE_API int polcue_masked_median_scale_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                   float max_d, int inst_lo, int inst_hi, int clamp_first, float* medians, float* scale,
                                   polcue_stream_t stream);
POLCUE_API int polcue_depth_errors_images_scaled_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px,
                                          float min_d, float max_d, int inst_lo, int inst_hi, int clamp_first,
                                          const float* pred_scale, double* sums, float* metrics, polcue_stream_t stream);

/* All mask groups of the evaluation loop in ONE launch (the reference makes 11-12 CPU passes, manydepth/trainer.py:920-980,
 * evaluation.py:169-213): group_ids[g] = instance id of group g (20, 40, ... 200, trainer.py:1389-1411) or -1 for
 * object == "all"; n_groups <= 16 HOST ints.  sums: B x n_groups x 8 doubles, metrics: B x n_groups x 7 floats or NULL.
 * One pass: each pixel's contributions are computed once and routed to "all" and to its material group. */
POLCUE_API int polcue_depth_errors_groups_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px,
                                   float min_d, float max_d, const int* group_ids, int n_groups, double* sums,
                                   float* metrics, polcue_stream_t stream);

/* One evaluation pass over a shard of the test split with no host work between the launches (BASELINE configs[4]):
 *   normals (B x 3 x H x W, or NULL to skip) = depth_to_normals(gt, K)                 manydepth/trainer.py:1484 (GT normals)
 *   sums / metrics = polcue_depth_errors_groups_f32(..)                                  trainer.py:1356-1430, evaluation.py:215-288
 *   mean_acc (1 + n_groups * 7 device doubles) = {B, sum over images of every group's seven per-image metrics}: the
 *   accumulators of the reference's mean over images (np.array(errors).mean(0), trainer.py:1426), additive across
 *   ranks -- one all-reduce of this vector and a division finish the evaluation.
 * Three launches; the stencil runs on an internal side stream forked from and joined back into `stream` (the two passes
 * only share their input and overlap); capturable in a CUDA graph. */
POLCUE_API int polcue_eval_pass_f32(const float* gt, const float* pred, const uint8_t* inst, const float* K, int B, int H, int W,
                         float min_d, float max_d, const int* group_ids, int n_groups, float* normals, double* sums,
                         float* metrics, double* mean_acc, polcue_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The one collective of the path, over NVLink peer memory.  The reference evaluates on one device and has no
 * collective; a sharded evaluation ends in a sum over ranks of {n_images, per-group sums of the per-image metrics}
 * (manydepth/trainer.py:1426-1428: np.array(errors).mean(0)) -- at most 1 + 16 x 7 doubles.  One process per GPU, all on
 * ONE node: every rank owns a small exchange block in its HBM, shared with the other ranks through CUDA IPC; an exchange
 * is one CTA per rank that stores its values into every rank's block through NVLink and adds what arrived in RANK ORDER
 * (all ranks get the identical, bitwise reproducible sum).  polcue_eval_pass_peer_f32 does it inside the last kernel of
 * the evaluation pass, so a sharded pass costs no extra launch and no library round trip.
 *   polcue_peer_create   allocates this rank's block on the CURRENT device; ipc_handle receives POLCUE_PEER_HANDLE_BYTES
 *                        bytes to hand to the other ranks (any transport: torch.distributed, MPI, a file).  world == 1
 *                        needs no handle and no connect.
 *   polcue_peer_connect  ipc_handles = the world handles in rank order; maps the other ranks' blocks.  All ranks must have
 *                        connected before the first exchange and must make the same sequence of exchanges; the calls of
 *                        one rank must be stream-ordered.  Capturable in a CUDA graph (the call counter lives on the device).
 *   polcue_peer_status   synchronising read of the number of exchanges made and of the first one whose wait for a peer
 *                        timed out (~4 s; its results are NaN), 0 = none.
 *   polcue_peer_destroy  after a barrier of the caller's: no peer may still be exchanging.
 * ------------------------------------------------------------------------------------------- */
#define POLCUE_PEER_HANDLE_BYTES 64
#define POLCUE_PEER_MAX_RANKS 16
#define POLCUE_PEER_MAX_VALUES 128
typedef struct polcue_peer polcue_peer;
POLCUE_API int polcue_peer_create(int world, int rank, polcue_peer** out, void* ipc_handle);
POLCUE_API int polcue_peer_connect(polcue_peer* peer, const void* ipc_handles);
POLCUE_API int polcue_peer_allreduce_f64(polcue_peer* peer, const double* in, int n, double* out, polcue_stream_t stream);
POLCUE_API int polcue_peer_status(polcue_peer* peer, unsigned long long* calls, unsigned long long* failed_call);
POLCUE_API int polcue_peer_destroy(polcue_peer* peer);
/* polcue_eval_pass_f32 whose last kernel also exchanges the accumulators: mean_acc receives this rank's own
 * {B, sums} as before, mean_acc_all (1 + n_groups * 7 device doubles) the sum over all ranks.  Three launches. */
POLCUE_API int polcue_eval_pass_peer_f32(const float* gt, const float* pred, const uint8_t* inst, const float* K, int B, int H, int W,
                              float min_d, float max_d, const int* group_ids, int n_groups, float* normals, double* sums,
                              float* metrics, double* mean_acc, polcue_peer* peer, double* mean_acc_all,
                              polcue_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POLCUE_H_ */
