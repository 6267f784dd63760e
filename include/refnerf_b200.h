/* refnerf_b200 -- C ABI of the B200-native Ref-NeRF per-ray rendering hot path.
 *
 * The reference (minfenli/refnerf-pl) is pure Python/PyTorch and has no FFI layer of its own
 * (SURVEY.md 8(b)); each entry point below names the reference Python function(s) it replaces
 * (path:line under the reference tree) and is what a ctypes binding in the reference would call
 * (INTEGRATION.md shows that binding).  Conventions:
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise;
 *   - no allocation, no global state besides a thread-local error string, work is enqueued on the
 *     caller's stream (`stream` is a cudaStream_t passed as void*), no host synchronisation;
 *   - return value 0 = success, negative = error (rn_last_error() gives the text).
 *   - N = rays, S = samples per ray (fenceposts S+1), R = N*S rows.
 */
#ifndef REFNERF_B200_H_
#define REFNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RN_API __attribute__((visibility("default")))
#else
#define RN_API
#endif

#define RN_OK 0
#define RN_ERR_ARG -1
#define RN_ERR_CUDA -2
#define RN_ERR_UNSUPPORTED -3

/* arithmetic of the MLP GEMM chains */
#define RN_PREC_FP32 0   /* fp32 FFMA GEMMs (exact-parity anchor)                     */
#define RN_PREC_BF16 1   /* tcgen05 kind::f16 BF16 operands, fp32 accumulate in TMEM  */
#define RN_PREC_BF16X3 2 /* split-BF16 (hi/lo, 3 MMAs): ~16-bit mantissa, parity mode */
#define RN_PREC_FP16 3   /* tcgen05 kind::f16 with FP16 (11-bit) weights, activations and gradient tiles (each dgrad
                            chain runs in its own power-of-two scale chosen on the device), fp32 accumulate;
                            fused chains only (gemm_impl = 0) */

/* number of parameter tensors of one NerfMLP in rn order (see rn_mlp_param_name) */
#define RN_MLP_NUM_PARAMS 46

RN_API const char* rn_last_error(void);
RN_API int rn_abi_version(void);

/* ---- K4a: interval resampling -------------------------------------------------------------
 * Replaces models.py:200-203 (logits), stepfun.sample_intervals/sample/invert_cdf/
 * integrate_weights (stepfun.py:134-258), math.sorted_interp (math.py:88-111) and the s_to_t ray
 * warp (coord.py:98).  sdist_in [N,s_in+1], weights_in [N,s_in], u [s_out] (the linspace grid of
 * stepfun.py:195-204), near/far [N].  Outputs sdist_out, tdist_out [N,s_out+1]; optional (may be
 * NULL) cw_out [N,s_in+1] and idx_out [N,s_out] (int32, idx = #{cw <= u} - 1).
 * The CDF is accumulated in fp64 and rounded per prefix, as torch.cumsum does for a float row on the CPU.
 * cw_in (optional, [N,s_in+1]): use this CDF instead of computing one from the weights (weights_in may then be NULL);
 * the search + interpolation stage is bit-exact against math.sorted_interp on the same CDF. */
RN_API int rn_resample(const float* sdist_in, const float* weights_in, const float* cw_in, const float* u, const float* near_, const float* far_,
                int64_t n_rays, int s_in, int s_out, float padding, float anneal, float dom_lo, float dom_hi,
                float* sdist_out, float* tdist_out, float* cw_out, int32_t* idx_out, void* stream);

/* ---- K4a': max_dilate_weights ------------------------------------------------------------------
 * Replaces stepfun.max_dilate_weights / max_dilate / weight_to_pdf / pdf_to_weight (stepfun.py:92-131) and, with
 * trim = 1, the [1:-1] trim of its caller (models.py:177-187).  t [N,s+1] sorted fenceposts, w [N,s] weights.
 * trim = 0: t_out [N,3s+1], w_out [N,3s];  trim = 1: t_out [N,3s-1], w_out [N,3s-2].
 * A warp per ray: 3-way merge by rank, range-max of the pdf over the window of dilated intervals containing each new
 * fencepost (sparse table in shared memory) instead of the reference's dense [3s+1, s] mask.  No gradient (the
 * resampler that consumes the result is detached, models.py:208-215). */
RN_API int rn_max_dilate_weights(const float* t, const float* w, int64_t n_rays, int s, float dilation, float dom_lo,
                          float dom_hi, int renormalize, int trim, float* t_out, float* w_out, void* stream);

/* ---- K4b: alpha compositing ---------------------------------------------------------------
 * Replaces render.compute_alpha_weights (render.py:132-149) and render.volumetric_rendering
 * (render.py:152-254, srgb_mapping='none'; the [N,3] tone map is applied by the host wrapper).
 * Per-sample inputs [N,S] / [N,S,3]; normals/normals_pred/roughness/tint may be NULL (extras off).
 * comp_out [N,16]: rgb(0:3) diffuse(3:6) specular(6:9) distance(9) acc(10) distance_mean(11).
 * extras_out [N,12] (NULL if unused): normals(0:3) normals_pred(3:6) tint(6:9) roughness(9).
 * pct_out [N,3] fp64 percentiles 5/50/95 (NULL if unused). */
RN_API int rn_composite_fwd(const float* density, const float* tdist, const float* dirs, const float* far_, const float* rgb,
                     const float* diffuse, const float* specular, const float* normals, const float* normals_pred,
                     const float* roughness, const float* tint, int64_t n_rays, int s, float bg, float* weights_out,
                     float* comp_out, float* extras_out, double* pct_out, void* stream);

/* Backward of rn_composite_fwd.  g_weights [N,S] (NULL = 0), g_comp [N,16] (distance_mean column
 * ignored), g_extras [N,12] (NULL = 0).  Outputs d_density [N,S], d_rgb/d_diffuse/d_specular [N,S,3]
 * and (NULL ok) d_normals_pred [N,S,3], d_roughness [N,S], d_tint [N,S,3]. */
RN_API int rn_composite_bwd(const float* density, const float* tdist, const float* dirs, const float* rgb, const float* diffuse,
                     const float* specular, const float* normals, const float* normals_pred, const float* roughness,
                     const float* tint, const float* weights, const float* comp, const float* g_weights,
                     const float* g_comp, const float* g_extras, int64_t n_rays, int s, float bg, float* d_density,
                     float* d_rgb, float* d_diffuse, float* d_specular, float* d_normals_pred, float* d_roughness,
                     float* d_tint, void* stream);

/* ---- interlevel / distortion losses ---------------------------------------------------------
 * stepfun.lossfun_outer (stepfun.py:67-89): t [N,s+1], w [N,s], t_env [N,se+1], w_env [N,se] ->
 * loss [N,s]; backward w.r.t. w_env only (t, w are detached by train_utils.py:151-162). */
RN_API int rn_lossfun_outer_fwd(const float* t, const float* w, const float* t_env, const float* w_env, int64_t n_rays, int s,
                         int se, float* loss_out, void* stream);
RN_API int rn_lossfun_outer_bwd(const float* t, const float* w, const float* t_env, const float* w_env, const float* g_loss,
                         int64_t n_rays, int s, int se, float* d_w_env, void* stream);
/* stepfun.lossfun_distortion (stepfun.py:261-272): loss [N]; backward w.r.t. w. */
RN_API int rn_distortion_fwd(const float* t, const float* w, int64_t n_rays, int s, float* loss_out, void* stream);
RN_API int rn_distortion_bwd(const float* t, const float* w, const float* g_loss, int64_t n_rays, int s, float* d_w,
                      void* stream);
/* train_utils.orientation_loss + predicted_normal_loss (train_utils.py:165-204), per ray, one pass:
 * ori[r] = sum_s w (min(0, n_t . (-v)))^2 with n_t = normals_pred (ori_target_is_pred) or the density-gradient
 * normals; pred[r] = sum_s w (1 - normals . normals_pred) (0 when normals == NULL).  The caller takes the means and
 * applies the loss multipliers.  weights [N,s]; normals, normals_pred [N,s,3]; viewdirs [N,3]; outputs [N].
 * Backward w.r.t. weights and normals_pred (the density-gradient normals are constants, models.py:603-609);
 * g_ori / g_pred [N] may be NULL (= 0). */
RN_API int rn_normal_losses_fwd(const float* weights, const float* normals, const float* normals_pred, const float* viewdirs,
                         int64_t n_rays, int s, int ori_target_is_pred, float* ori_out, float* pred_out, void* stream);
RN_API int rn_normal_losses_bwd(const float* weights, const float* normals, const float* normals_pred, const float* viewdirs,
                         const float* g_ori, const float* g_pred, int64_t n_rays, int s, int ori_target_is_pred,
                         float* d_weights, float* d_normals_pred, void* stream);

/* train_utils.compute_data_loss (train_utils.py:33-88) of one level, the loss epilogue of SURVEY 8(f) rank 2: ONE launch
 * instead of ~10 elementwise / reduction launches over [N,3] tensors.  rgb, gt [N,3]; lossmult [N] (the reference
 * broadcasts rays.lossmult over the channels) or NULL (= 1, config.disable_multiscale_loss).  sums_out[3] =
 * { sum lm (rgb-gt)^2,  sum lm term(rgb-gt),  sum lm } with term = r^2 (charb == 0, data_loss_type 'mse') or
 * sqrt(r^2 + charb_padding^2) ('charb'); fp64 accumulation in a fixed order (bit-reproducible).  The caller divides
 * (mse = s0 / s2, data loss = s1 / s2) and applies data_(coarse_)loss_mult.  Backward w.r.t. rgb given g_sums[0..1]. */
RN_API int rn_data_loss_fwd(const float* rgb, const float* gt, const float* lossmult, int64_t n_rays, int charb,
                     float charb_padding, float* sums_out, void* stream);
RN_API int rn_data_loss_bwd(const float* rgb, const float* gt, const float* lossmult, const float* g_sums, int64_t n_rays,
                     int charb, float charb_padding, float* d_rgb, void* stream);

/* ---- K1 / K2 unit-level entry points (the fused MLP uses the same device code) -------------
 * rn_encode: render.cast_rays (render.py:105-129, cone, full cov) + coord.lift_and_diagonalize
 * (coord.py:129-133, octahedron-1 basis) + coord.integrated_pos_enc (coord.py:107-126, degrees
 * 0..15) -> feat [R,96]. */
RN_API int rn_encode(const float* tdist, const float* origins, const float* dirs, const float* radii, int64_t n_rays, int s,
              float* feat_out, void* stream);
/* rn_ide: ref_utils.generate_ide_fn(5) (ref_utils.py:98-161): dirs [R,3], kappa_inv [R] -> [R,72]. */
RN_API int rn_ide(const float* dirs, const float* kappa_inv, int64_t n, float* out, void* stream);

/* ---- ray generation (the caller side of the path) ----------------------------
 * rn_pixels_to_rays: camera_utils.pixels_to_rays (camera_utils.py:502-614) for perspective cameras without lens
 * distortion: pixel centres through pixtocams[cam] [C,3,3] and camtoworlds[cam] [C,3,4] (fp32, row-major) ->
 * origins, directions (not normalised), viewdirs [n,3], radii [n] (mip-NeRF cone radii from the +x / +y
 * neighbour rays), imageplane [n,2].  pixtocam_ndc != NULL additionally applies camera_utils.convert_to_ndc
 * (camera_utils.py:31-97, near = 1) to origins / directions and takes the radii from the NDC origin offsets.
 * cam_idx may be NULL (single camera).  Arithmetic in float64 like the reference's numpy path, results fp32. */
RN_API int rn_pixels_to_rays(const int32_t* pix_x, const int32_t* pix_y, const int32_t* cam_idx, const float* pixtocams,
                      const float* camtoworlds, const float* pixtocam_ndc, int64_t n, float* origins, float* directions,
                      float* viewdirs, float* radii, float* imageplane, void* stream);

/* ---- K3: the NerfMLP (models.py:533-750) ---------------------------------------------------
 * Parameter order (RN_MLP_NUM_PARAMS pointers, weights [out,in] row-major as in PyTorch):
 *   spatial_net.{0..7}.{weight,bias}, raw_density, grad_pred, raw_roughness, raw_rgb_diffuse,
 *   raw_tint, bottleneck (each weight,bias), viewdir_mlp.{0..7}.{weight,bias}, rgb.{weight,bias}. */
RN_API const char* rn_mlp_param_name(int i);
RN_API int64_t rn_mlp_param_numel(int i);

typedef struct RnMlpConfig {
  int prec;                 /* RN_PREC_*                                                   */
  int srgb_mapping;         /* NerfMLP.srgb_mapping (models.py:712)                        */
  int srgb_normalization;   /* NerfMLP.srgb_mapping_normalization (models.py:718)          */
  float density_bias;       /* models.py:623                                               */
  float roughness_bias;     /* models.py:640                                               */
  float rgb_premultiplier;  /* models.py:699                                               */
  float rgb_bias;
  float rgb_padding;        /* models.py:729                                               */
  int chunk_rows;           /* rows processed per internal chunk (multiple of 128)         */
  int gemm_impl;            /* 0 = default (fused tcgen05 chains: bf16 / fp16 / split-bf16), 1 = SIMT GEMMs, 2 = per-layer tcgen05 */
  int deterministic_wgrad;  /* 1 = weight / bias gradients are reduced over the row-split CTAs in a fixed order (per-CTA partial
                               tiles + one reduction pass) instead of with floating-point atomics: bit-reproducible steps */
} RnMlpConfig;

/* bytes of the packed-weight blob / of the scratch workspace for the given chunk size.
 * training: 0 = eval forward, 1 = training forward + backward that recomputes the activations per chunk,
 * 2 = training forward + backward that share a "saved" region (rn_mlp_saved_bytes) instead of recomputing. */
RN_API size_t rn_mlp_packed_bytes(int prec);
RN_API size_t rn_mlp_workspace_bytes(const RnMlpConfig* cfg, int training);
/* bytes of the saved-activation region for n_rows = n_rays * s rows (chain inputs, every hidden activation and
 * the raw head outputs): written by rn_mlp_forward, read by rn_mlp_backward of the same call */
RN_API size_t rn_mlp_saved_bytes(const RnMlpConfig* cfg, int64_t n_rows);
/* fp32 parameters -> padded (hi/lo bf16 or fp32) GEMM operands incl. transposed copies for dgrad */
RN_API int rn_mlp_pack(const float* const* params, void* packed, int prec, void* stream);

typedef struct RnMlpOutputs {
  float* density;      /* [R]    */
  float* rgb;          /* [R,3]  */
  float* normals;      /* [R,3]  NULL in eval mode (models.py:603) */
  float* normals_pred; /* [R,3]  */
  float* grad_pred;    /* [R,3]  */
  float* tint;         /* [R,3]  */
  float* diffuse;      /* [R,3]  */
  float* specular;     /* [R,3]  */
  float* roughness;    /* [R]    */
} RnMlpOutputs;

/* Forward for one level: cast_rays + IPE + spatial net + heads (+ density-gradient normals when
 * out->normals != NULL) + reflect/IDE/n.v + view net + colour combine. */
RN_API int rn_mlp_forward(const RnMlpConfig* cfg, const void* packed, const float* tdist, const float* origins,
                   const float* dirs, const float* viewdirs, const float* radii, int64_t n_rays, int s,
                   const RnMlpOutputs* out, void* workspace, size_t workspace_bytes, void* saved, size_t saved_bytes,
                   void* stream);

/* Backward: upstream gradients in `g` (same shapes as RnMlpOutputs; NULL members = zero; normals
 * is a detached constant in the reference and is ignored).  Accumulates (+=) into
 * grads[RN_MLP_NUM_PARAMS] (fp32, shapes of the parameters).  `saved` = the region the training forward
 * filled (NULL: the activations are recomputed chunk by chunk). */
RN_API int rn_mlp_backward(const RnMlpConfig* cfg, const void* packed, const float* tdist, const float* origins,
                    const float* dirs, const float* viewdirs, const float* radii, int64_t n_rays, int s,
                    const RnMlpOutputs* g, float* const* grads, void* workspace, size_t workspace_bytes,
                    const void* saved, size_t saved_bytes, void* stream);

/* ---- GEMM building block, exposed for unit tests and the roofline bench ---------------------
 * C[M,N] (fp32, ldc) = act( A[M,K] * B[N,K]^T + bias ), A/B fp32 row-major, converted per `prec`
 * into scratch (scratch_bytes >= rn_gemm_scratch_bytes).  impl: 0 = tcgen05, 1 = SIMT.
 * wgrad variant: C[N,K] += sum_m dY[m,n] X[m,k]. */
RN_API size_t rn_gemm_scratch_bytes(int64_t m, int n, int k);
RN_API int rn_gemm_test(const float* a, const float* b, const float* bias, int64_t m, int n, int k, int relu, int prec,
                 int impl, float* c, void* scratch, size_t scratch_bytes, void* stream);
RN_API int rn_wgrad_test(const float* dy, const float* x, int64_t m, int n, int k, int prec, int impl, float* c, void* scratch,
                  size_t scratch_bytes, void* stream);
/* times `iters` back-to-back launches of one 256x256 layer GEMM over m rows; returns ms per launch */
RN_API int rn_gemm_bench(int64_t m, int prec, int impl, int iters, float* ms_out, void* scratch, size_t scratch_bytes,
                  void* stream);

/* fp16 mode only: how many times (per wgrad launch and thread) a gradient element saturated the fp16 range since the
 * last reset.  The dgrad chains run in one power-of-two scale per chain chosen from the chain's seed tile; an element
 * growing more than 16x through a chain is clamped to +-65504 by the saturating conversion, which would bias the weight
 * gradients silently.  0 = no clamping happened.  Synchronises the device; reset != 0 clears the counter.
 * (The split-bf16 and bf16 modes have fp32's exponent range and need no such check.) */
RN_API int64_t rn_fp16_saturation_count(int reset);

/* ---- instrumentation -------------------------------------------------------------------------
 * rn_launch_count: kernels launched by this library since load (process-wide).
 * rn_prof_enable(1) makes the GEMM launchers bracket every launch with CUDA events on the launching
 * stream; rn_prof_summary(cls, ...) synchronises the recorded events and returns, for kernel class
 * cls (0 = tcgen05 fwd/dgrad GEMM, 1 = tcgen05 wgrad, 2 = SIMT GEMMs, 3 = fused tcgen05 chains; time only: 4 = encode,
 * 5 / 6 = heads prologue forward / backward, 7 = IPE-gradient normals, 8 = colour combine, 9 = pack / unpack / conversion
 * glue, 10 = the ray kernels: resample, compositing, losses), the number of launches, their summed device time in ms and
 * the algorithmic FLOPs they carried, then clears the class. */
RN_API int64_t rn_launch_count(void);
RN_API int rn_prof_enable(int on);
RN_API int rn_prof_summary(int cls, int64_t* launches, double* total_ms, double* algo_flops);
/* the same plus the FLOPs the tensor pipe executed for those launches (padded shapes x MMAs per K step: 3 for the
 * split-bf16 forward / normals chains, 2 for its loss-backward chains, 1 otherwise) */
RN_API int rn_prof_summary2(int cls, int64_t* launches, double* total_ms, double* algo_flops, double* exec_flops);

#ifdef __cplusplus
}
#endif
#endif /* REFNERF_B200_H_ */
