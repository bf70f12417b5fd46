/*
 * simplenerf_b200 -- C ABI of the B200-native SimpleNeRF volumetric-rendering hot path.
 *
 * The reference (NagabhushanSN95/SimpleNeRF) is pure PyTorch and has no FFI; the "interface each
 * entry point replaces" is therefore the Python function of src/models/SimpleNeRF01.py cited
 * beside it.  The host-side mirror of that interface (same class contract as the reference's
 * models.SimpleNeRF01.SimpleNeRF) is simplenerf_b200/models/FusedSimpleNeRF01.py, which binds
 * these symbols through ctypes.
 *
 * Conventions
 *   - every function returns an int status (SNERF_OK == 0); nothing throws across the ABI;
 *     snerf_last_error() returns a thread-local message for the last non-zero status;
 *   - all pointers are DEVICE pointers (unless named host_*), 16-byte aligned, row-major, fp32;
 *   - the caller allocates every buffer including workspaces; the library never allocates device
 *     memory and never synchronises: kernels are enqueued on `stream` (a cudaStream_t);
 *   - "nullable" pointers switch the corresponding optional input/output off.
 */
#ifndef SIMPLENERF_B200_H
#define SIMPLENERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNERF_OK 0
#define SNERF_ERR_INVALID 1    /* bad argument (message says which)            */
#define SNERF_ERR_CUDA 2       /* a CUDA runtime call / launch failed            */
#define SNERF_ERR_UNSUPPORTED 3 /* shape outside what the kernels are built for  */

#define SNERF_ABI_VERSION 1

/* Shape of one reference `MLP` (ctor src/models/SimpleNeRF01.py:561-609). */
typedef struct snerf_mlp_desc {
    int32_t depth;        /* points_net_depth (8)                                              */
    int32_t width;        /* points_net_width (256)                                            */
    int32_t skip_layer;   /* skip concat happens after this trunk layer (4), :580              */
    int32_t pts_degree;   /* points_positional_encoding_degree (10)                            */
    int32_t trunk_degree; /* PE bands fed to the sigma trunk: 10, or 3 for points-augmentation
                             (points_sigma_positional_encoding_degree, :576-578)               */
    int32_t view_degree;  /* views_positional_encoding_degree (4); 0 when use_view_dirs=False  */
    int32_t view_width;   /* views_net_width (128); 0 when there is no view branch             */
    int32_t head_out;     /* pts_output_linear rows: 1 (sigma) or 4 (sigma+rgb, views-aug)     */
} snerf_mlp_desc;

/* Order of the fp32 parameter pointers handed to the MLP entry points (torch.nn.Linear layout:
 * weight [out,in] row-major, bias [out]).  Entries of absent layers are NULL.                  */
enum {
    SNERF_P_TRUNK_W0 = 0, /* pts_linears.i.weight at 2*i, .bias at 2*i+1, i < 8               */
    SNERF_P_HEAD_W = 16,  /* pts_output_linear                                                 */
    SNERF_P_HEAD_B = 17,
    SNERF_P_FEAT_W = 18,  /* feature_linear                                                    */
    SNERF_P_FEAT_B = 19,
    SNERF_P_VIEW_W = 20,  /* views_linears.0                                                   */
    SNERF_P_VIEW_B = 21,
    SNERF_P_RGB_W = 22,   /* views_output_linear                                               */
    SNERF_P_RGB_B = 23,
    SNERF_P_COUNT = 24
};

/* flags */
#define SNERF_FLAG_NDC 1u          /* z is NDC depth; use rays_d_ndc for delta, convert depth (:437-441, :456-460) */
#define SNERF_FLAG_WHITE_BKGD 2u   /* rgb += 1 - acc (:462-463)                                  */
#define SNERF_FLAG_LINDISP 4u      /* sample linearly in disparity (:288-289)                    */
#define SNERF_FLAG_SAVE_FOR_BWD 8u /* MLP forward keeps activations in the workspace             */
#define SNERF_FLAG_PRECISE 16u     /* fp32 CUDA-core MLP instead of the bf16 tcgen05 MLP         */
#define SNERF_FLAG_VIS_GRAD 32u    /* snerf_mlp_backward: add what snerf_visibility_backward left in the workspace */
#define SNERF_FLAG_VIS_HEAD 64u    /* tensor path: snerf_mlp_forward also keeps what the snerf_visibility_* calls read (the view
                                      layer's point part in bf16, the rays' own direction encodings); ignored by the precise path */

int snerf_abi_version(void);
const char* snerf_last_error(void);
/* 1 if the library was built with the tcgen05 (sm_100a) MLP kernels. */
int snerf_has_tensor_path(void);

/* ---- a3: stratified coarse sampling (get_z_vals_coarse, :272-302) --------------------------
 * t_vals [n_samples] = torch.linspace(0,1,n_samples) (computed by the host exactly as the
 * reference does); near/far [n_rays]; t_rand [n_rays,n_samples] nullable (NULL = no perturb).  */
int snerf_sample_coarse(const float* near, const float* far, const float* t_vals, const float* t_rand,
                        float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream);

/* ---- a11+a12: hierarchical resampling (get_z_vals_fine :304-315, sample_pdf :328-361) ------
 * z_coarse, weights_coarse [n_rays,s_coarse]; u [n_rays or 1, n_new] with row stride u_stride
 * (0 = one row broadcast: the deterministic linspace).  z_fine [n_rays, s_coarse+n_new] sorted.
 * Optional debug outputs (nullable): samples [n_rays,n_new] (unsorted, in u order),
 * cdf [n_rays,s_coarse-1], below/above int32 [n_rays,n_new].                                   */
int snerf_sample_fine(const float* z_coarse, const float* weights_coarse, const float* u, int u_stride,
                      float* z_fine, float* samples_dbg, float* cdf_dbg, int32_t* below_dbg, int32_t* above_dbg,
                      int n_rays, int s_coarse, int n_new, void* stream);

/* ---- a9+a10: alpha compositing (volume_rendering :430-483, convert_depth_from_ndc :485-502) -
 * sigma, z [n_rays,S]; rgb [n_rays,S,3]; rays_* [n_rays,3] (rays_d_ndc only with FLAG_NDC).
 * Per-ray outputs [n_rays] / [n_rays,3]; depth_ndc / depth_var_ndc only with FLAG_NDC.
 * Per-sample outputs alpha / visibility / weights [n_rays,S] are nullable.                     */
int snerf_composite_forward(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                            const float* rays_d, const float* rays_d_ndc, float* rgb_map, float* acc,
                            float* depth, float* depth_var, float* depth_ndc, float* depth_var_ndc,
                            float* alpha, float* visibility, float* weights, int n_rays, int n_samples,
                            uint32_t flags, void* stream);

/* Backward of the above.  Incoming gradients (all nullable): d_rgb_map [n,3], d_acc, d_depth,
 * d_depth_var, d_depth_ndc, d_depth_var_ndc [n], d_alpha / d_visibility / d_weights [n,S].
 * Outputs: d_sigma [n,S], d_rgb [n,S,3].                                                       */
int snerf_composite_backward(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                             const float* rays_d, const float* rays_d_ndc, const float* d_rgb_map,
                             const float* d_acc, const float* d_depth, const float* d_depth_var,
                             const float* d_depth_ndc, const float* d_depth_var_ndc, const float* d_alpha,
                             const float* d_visibility, const float* d_weights, float* d_sigma, float* d_rgb,
                             int n_rays, int n_samples, uint32_t flags, void* stream);

/* ---- a4-a8: point generation + positional encoding + MLP (run_network :363-428, MLP :560-715)
 * pts = rays_o + rays_d * z is never materialised.  view_dirs [n_rays,3] (nullable when the MLP
 * has no view branch).  sigma_noise [n_rays*S] nullable (already multiplied by raw_noise_std).
 * host_params: HOST array of SNERF_P_COUNT device pointers.  packed: device image made by
 * snerf_pack_weights (bf16 tensor path; ignored with FLAG_PRECISE).
 * Outputs sigma [n_rays*S], rgb [n_rays*S,3].                                                  */
size_t snerf_mlp_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, uint32_t flags);
size_t snerf_packed_weights_bytes(const snerf_mlp_desc* desc);
int snerf_pack_weights(const snerf_mlp_desc* desc, const float* const* host_params, void* packed, void* stream);

int snerf_mlp_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                      const float* rays_o, const float* rays_d, const float* view_dirs, const float* z,
                      const float* sigma_noise, float* sigma, float* rgb, void* workspace, size_t workspace_bytes,
                      int n_rays, int n_samples, uint32_t flags, void* stream);

/* ---- X1: evaluation with the samples kept on chip (run_network + volume_rendering in one pass, tensor path) ----------
 * Replaces  raw = run_network(...); volume_rendering(raw, z_vals, ...)  (src/models/SimpleNeRF01.py:153-160, :219-224, :430-483)
 * for calls that do not return the raw network outputs (retraw=False, :265-269): sigma / rgb of a sample live in registers
 * from the head epilogue of the MLP kernel to the compositing arithmetic; what leaves the SM is one 48-byte record per
 * 32 consecutive samples of a ray (transmittance product, weight / colour / depth sums and centred second moments of the
 * segment), which a warp-per-ray pass folds into the per-ray maps.  Needs n_samples % 32 == 0 and an MLP with a view branch
 * (SNERF_ERR_UNSUPPORTED otherwise: the caller keeps using snerf_mlp_forward + snerf_composite_forward).
 * pts_o / pts_d: the rays the points are generated from (NDC rays with SNERF_FLAG_NDC, :142); rays_o / rays_d: the camera rays
 * (only their z components are read, for the NDC depth conversion :495-501; may equal pts_* otherwise).
 * Outputs: rgb_map [n,3], acc, depth, depth_var [n], depth_ndc, depth_var_ndc [n] (NDC only), alpha [n,S] (nullable; :446, a
 * key of the reference's output dict), weights [n,S] (nullable; the coarse pass hands them to snerf_sample_fine).           */
size_t snerf_render_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, uint32_t flags);
int snerf_render_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                         const float* pts_o, const float* pts_d, const float* view_dirs, const float* z,
                         const float* rays_o, const float* rays_d, float* rgb_map, float* acc, float* depth,
                         float* depth_var, float* depth_ndc, float* depth_var_ndc, float* alpha, float* weights,
                         void* workspace, size_t workspace_bytes, int n_rays, int n_samples, uint32_t flags, void* stream);

/* ---- in-kernel random numbers (production draws in the kernels that consume them) ----------------
 * The reference draws t_rand (:299), u (:341) and the sigma noise (:670) on the CPU generator and copies them to the device.
 * The *_rng entry points draw them where they are consumed, from Philox4x32-10 keyed by `seed`, counter = (element / 4, `offset`):
 * element e of a draw is the same number in every kernel, and snerf_fill_random writes exactly those numbers to memory, so
 *   snerf_sample_coarse(t_rand = fill(uniform))        == snerf_sample_coarse_rng      bit for bit,
 *   snerf_sample_fine(u = fill(uniform), stride n_new)  == snerf_sample_fine_rng,
 *   snerf_mlp_forward(sigma_noise = fill(normal, std))  == snerf_mlp_forward_rng        (tensor path).
 * Use a fresh `offset` per draw.  Elements: ray * n_samples + k (coarse), ray * n_new + k (fine), point index (noise).        */
int snerf_fill_random(float* out, long long n, int normal, float scale, uint64_t seed, uint64_t offset, void* stream);
int snerf_sample_coarse_rng(const float* near, const float* far, const float* t_vals, uint64_t seed, uint64_t offset,
                            float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream);
int snerf_sample_fine_rng(const float* z_coarse, const float* weights_coarse, uint64_t seed, uint64_t offset, float* z_fine,
                          int n_rays, int s_coarse, int n_new, void* stream);
int snerf_mlp_forward_rng(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed, const float* rays_o,
                          const float* rays_d, const float* view_dirs, const float* z, float noise_std, uint64_t seed,
                          uint64_t offset, float* sigma, float* rgb, void* workspace, size_t workspace_bytes, int n_rays,
                          int n_samples, uint32_t flags, void* stream);

/* Backward: consumes the workspace left by a forward run with FLAG_SAVE_FOR_BWD (same desc,
 * shapes and flags).  d_sigma [P], d_rgb [P,3] are gradients w.r.t. the forward outputs.
 * host_grads: HOST array of SNERF_P_COUNT device pointers, same shapes as the parameters;
 * gradients are ACCUMULATED (+=) into them (caller zeroes).  No gradient flows to rays / z.   */
int snerf_mlp_backward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                       const float* rays_o, const float* rays_d, const float* view_dirs, const float* z,
                       const float* sigma, const float* rgb, const float* d_sigma, const float* d_rgb,
                       float* const* host_grads, void* workspace, size_t workspace_bytes, int n_rays,
                       int n_samples, uint32_t flags, void* stream);

/* ---- a14 / N4: secondary-view visibility head (predict_visibility=True) ------------------------------
 * With predict_visibility the reference's views_output_linear has a fourth row (src/models/SimpleNeRF01.py:596-608):
 *   visibility  = sigmoid(row 3 . relu(view layer))                                   (:710-713)
 *   visibility2 = the same with the view layer re-run per OTHER view on [feature | enc_hi | PE(dir2)], dir2 the unit
 *                 vector from that view's camera centre rays_o2[ray, v] to the sample point (:317-325, :646-649).
 * params[SNERF_P_RGB_W] / [SNERF_P_RGB_B] then point at [4, view_width] / [4].  Both calls work on the workspace that
 * snerf_mlp_forward(SNERF_FLAG_PRECISE | SNERF_FLAG_SAVE_FOR_BWD) of the same MLP and points left behind, plus their own
 * (snerf_visibility_workspace_bytes).  z is NDC depth with SNERF_FLAG_NDC (converted as :319-321).
 * Backward order within a step: snerf_visibility_backward FIRST (it adds the fourth-row and view-layer gradients to
 * `grads` and leaves d hv / d feature in the MLP workspace), then snerf_mlp_backward(flags | SNERF_FLAG_VIS_GRAD).
 * Tensor path (no SNERF_FLAG_PRECISE): the same calls and order on the workspace of snerf_mlp_forward(SNERF_FLAG_VIS_HEAD
 * [| SNERF_FLAG_SAVE_FOR_BWD]); pass the forward call's flags (plus SNERF_FLAG_NDC) to both visibility calls.  The view
 * layer's point part W_vf h (+ enc_hi part) is shared by all views and comes from the tcgen05 kernel (bf16); the per-view
 * direction part (27 x 128 per point and view), the ReLU, the fourth-row dot product and their gradients run on the CUDA
 * cores in fp32 (vis_tc.cu).  n_other <= 8 there.                                                                        */
size_t snerf_visibility_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, int n_other);
int snerf_visibility_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* mlp_workspace,
                             const float* rays_o, const float* rays_d, const float* z, const float* rays_o2,
                             float* visibility, float* visibility2, void* workspace, size_t workspace_bytes, int n_rays,
                             int n_samples, int n_other, uint32_t flags, void* stream);
int snerf_visibility_backward(const snerf_mlp_desc* desc, const float* const* host_params, void* mlp_workspace,
                              const float* rays_o, const float* rays_d, const float* z, const float* rays_o2,
                              const float* visibility, const float* visibility2, const float* d_visibility,
                              const float* d_visibility2, float* const* host_grads, void* workspace, size_t workspace_bytes,
                              int n_rays, int n_samples, int n_other, uint32_t flags, void* stream);

/* visibility2 map [n_rays, n_other] = sum_s weights[s] visibility2[s, v] / (acc + 1e-6)  (volume_rendering :479-482) and its
 * backward: d_visibility2 [n_rays, n_samples, n_other]; d_weights [n_rays, n_samples] and d_acc [n_rays] are added by the
 * caller to the incoming gradients of snerf_composite_backward.  n_other <= 8.                                          */
int snerf_visibility2_composite_forward(const float* weights, const float* acc, const float* visibility2, float* visibility2_map,
                                        int n_rays, int n_samples, int n_other, void* stream);
int snerf_visibility2_composite_backward(const float* weights, const float* acc, const float* visibility2, const float* visibility2_map,
                                         const float* d_visibility2_map, float* d_visibility2, float* d_weights, float* d_acc,
                                         int n_rays, int n_samples, int n_other, void* stream);

/* ---- (f) N1, one frame per call: rays of a camera pose and output post-processing -------------------
 * (DataPreprocessor01.py get_rays :351-368, get_ndc_rays :371-389, get_view_dirs :392-394, post_process_image
 * :1106-1109, post_process_depth :1112-1114; replaces the host numpy pass of create_test_data :807-895).
 * pose34 (rows of [R|t]) and kinv33 (inverse intrinsic) are HOST arrays; s_w = -1/(W/(2 fx)), s_h = -1/(H/(2 fy)),
 * two_near = 2 near, all computed by the caller in fp32 like the reference.  Rays of image rows [row0, row0+n_rows)
 * are written row-major to the [n_rows*W, 3] device outputs (the *_ndc pair only if ndc != 0).                     */
int snerf_generate_rays(const float* pose34, const float* kinv33, float s_w, float s_h, float near, float two_near,
                        int h, int w, int row0, int n_rows, int ndc, float* rays_o, float* rays_d, float* view_dirs,
                        float* rays_o_ndc, float* rays_d_ndc, void* stream);
/* image[n_pixels,3] = uint8(round_half_even(clip(rgb,0,1)*255)); every depth map (<= 4, host array of device pointers)
 * is clipped to >= 0 in place.                                                                                       */
int snerf_postprocess_frame(const float* rgb, uint8_t* image, float* const* depth_maps, int n_maps, long long n_pixels,
                            void* stream);

/* ---- (f) N2, optimizer step of the process-per-GPU trainer ------------------------------------
 * Adam update of n_tensors (<= SNERF_ADAM_MAX_TENSORS) fp32 tensors in ONE launch; replaces the per-tensor work of
 * torch.optim.Adam as the reference uses it (src/Trainer01.py:516; no weight decay, no amsgrad):
 *   m = m + (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 * params / grads / exp_avg / exp_avg_sq are HOST arrays of device pointers (16-byte aligned), numel a host array.   */
#define SNERF_ADAM_MAX_TENSORS 64
int snerf_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                    const long long* numel, int n_tensors, float lr, float beta1, float beta2, float eps, int step,
                    void* stream);

/* ---- (f) N3, the masked per-ray losses of the training step ---------------------------------------
 * One "stream" = one masked mean-squared error of the reference's loss modules: MSE01/02/03.compute_mse
 * (src/loss_functions/MSE01.py:53-67: pred[mask] vs target[mask], mean over channels then over rays) and
 * SparseDepthMSE01/02/03.compute_depth_loss (SparseDepthMSE01.py:58-71, one channel).  The weighted sum is
 * LossComputer.compute_losses (LossComputer01.py:33-52).  The table is a HOST array; its pointers are device pointers. */
#define SNERF_LOSS_MAX_STREAMS 8
#define SNERF_LOSS_SQUARED 0          /* mean over channels and masked rays of (pred - target)^2                     */
#define SNERF_LOSS_ABSOLUTE 1         /* ... of |pred - target|        (VisibilityLoss01.compute_mae, :70-74)        */
#define SNERF_LOSS_PRIOR_SHORTFALL 2  /* mean over masked rays of sum_c target_c (1 - pred_c)
                                         (VisibilityPriorLoss01.compute_consistency_loss, :64-80; target = prior)   */
typedef struct snerf_loss_stream {
    const float* pred;    /* [n_rays, channels]                                       */
    const float* target;  /* [n_rays, channels]                                       */
    const uint8_t* mask;  /* [n_rays] bool; nullable = every ray                      */
    float* grad;          /* [n_rays, channels], written by the backward call only    */
    int32_t channels;     /* 3 (rgb), 1 (depth), nf-1 or the samples per ray; 1..1024 */
    float weight;         /* loss weight (LossComputer.get_loss_weight)               */
    int32_t kind;         /* SNERF_LOSS_SQUARED / _ABSOLUTE / _PRIOR_SHORTFALL        */
} snerf_loss_stream;
size_t snerf_ray_losses_workspace_bytes(void);
/* values[n_streams + 1]: the mean of every stream (0 when its mask is empty), then the weighted total;
 * counts[n_streams]: masked-in rays per stream.  The workspace must be zeroed once after allocation and belongs to
 * one stream at a time (block partials and a ticket counter live in it between the blocks of one launch).           */
int snerf_ray_losses_forward(const snerf_loss_stream* streams, int n_streams, int n_rays, float* values,
                             int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
/* grad_values[n_streams + 1] (device): incoming gradient of `values`; stream s receives
 * (grad_values[s] + grad_values[n_streams] * weight_s) * 2 (pred - target) / (count_s * channels) on masked rays, else 0
 * (sign(pred - target) / (count_s * channels) and -target / count_s for the other two kinds).                           */
int snerf_ray_losses_backward(const snerf_loss_stream* streams, int n_streams, int n_rays, const int32_t* counts,
                              const float* grad_values, void* stream);

/* Per-ray loss maps of the same streams (validation: src/Trainer01.py:195-196 passes return_loss_maps, :252-259 saves them):
 * stream s writes [n_rays] floats through its `grad` pointer (used as the OUTPUT here): the ray's error averaged over its
 * channels -- the `loss_maps` entry of compute_mse (MSE01.py:56, :63-66) -- and 0 on masked-out rays.                     */
int snerf_ray_loss_maps(const snerf_loss_stream* streams, int n_streams, int n_rays, void* stream);

/* Patch-reprojection depth losses: compute_loss_nerf of PointsAugmentationDepthLoss02 / ViewsAugmentationDepthLoss02 /
 * CoarseFineConsistencyLoss02 (src/loss_functions/PointsAugmentationDepthLoss02.py:98-176, identical in the three
 * modules; reprojection: src/utils/CommonUtils01.py:45-72).  One call compares ONE main depth with up to 4 other
 * depths.  Host struct; every pointer is a device pointer.
 *   proj[v]    = intrinsics[0] @ diag(1,-1,-1) @ poses[v,:3,:3]^T  (row-major 3x3),  origins[v] = poses[v,:3,3],
 *   closest[v] = index of the second smallest camera distance from view v (:126-130).
 * Default behaviour is the reference's: through its in-place masking on detach() aliases (:204-206) only the term that
 * pulls the MAIN depth towards the other depth survives; SNERF_REPROJ_SYMMETRIC adds the other term.              */
#define SNERF_REPROJ_MAX_OTHERS 4
#define SNERF_REPROJ_SYMMETRIC 1u
typedef struct snerf_reproj_args {
    const float* depth_main;                            /* [n_rays]                                   */
    const float* depth_other[SNERF_REPROJ_MAX_OTHERS];  /* [n_rays] each                              */
    float* grad_main;                                   /* backward output [n_rays]                   */
    float* grad_other[SNERF_REPROJ_MAX_OTHERS];         /* backward outputs, nullable                 */
    float weight[SNERF_REPROJ_MAX_OTHERS];              /* loss weights                               */
    int32_t n_others;
    const float* rays_o;                                /* [n_rays,3]                                 */
    const float* rays_d;                                /* [n_rays,3]                                 */
    const int32_t* pixel_id;                            /* [n_rays,3] = (view, x, y)                  */
    const uint8_t* mask_nerf;                           /* [n_rays] bool; nullable = every ray        */
    const float* images;                                /* [n_views, height, width, 3]                */
    const float* proj;                                  /* [n_views, 9]                               */
    const float* origins;                               /* [n_views, 3]                               */
    const int32_t* closest;                             /* [n_views]                                  */
    int32_t n_views, height, width;
    int32_t half_patch;                                 /* patch_size // 2 (square patches, <= 2)     */
    float rmse_threshold;
    uint32_t flags;
} snerf_reproj_args;
/* codes[n_others, n_rays]: bit 0 = the main model is the more accurate one on this ray, bit 1 = the other model is;
 * values[n_others + 1]: loss per pair, then the weighted total; counts[1]: rays with mask_nerf.
 * Workspace: snerf_ray_losses_workspace_bytes(), zeroed once.                                                      */
int snerf_reprojection_losses_forward(const snerf_reproj_args* args, int n_rays, uint8_t* codes, float* values,
                                      int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
int snerf_reprojection_losses_backward(const snerf_reproj_args* args, int n_rays, const uint8_t* codes,
                                       const int32_t* counts, const float* grad_values, void* stream);

/* ---- (f) N3, batch assembly: the per-ray tables of one training batch in one launch ---------------
 * Replaces the `-1 * ones` + `t[mask] = table[indices[mask]]` pairs of load_nerf_cached_batch
 * (src/data_preprocessors/DataPreprocessor01.py:572-620) and load_sparse_depth_cached_batch (:655-700):
 *   dst[i, :] = (mask == NULL || mask[i]) ? src[indices[i], :] : fill.   Rows are row_bytes wide (a multiple of 4),
 * fill_bits is the 32-bit pattern of the fill value (-1.0f or int32 -1).  Host array of tables, device pointers.
 * indices[i] of a masked-in row must be a valid source row (not checked, like every device-side index here).            */
#define SNERF_GATHER_MAX_TABLES 24
typedef struct snerf_gather_table {
    const void* src;      /* [n_source_rows, row_bytes]              */
    void* dst;            /* [n_rows, row_bytes]                     */
    const uint8_t* mask;  /* [n_rows] bool; nullable = every row     */
    int32_t row_bytes;
    uint32_t fill_bits;
} snerf_gather_table;
int snerf_gather_rows(const snerf_gather_table* tables, int n_tables, const int64_t* indices, int n_rows, void* stream);

/* Measurement hook of snerf_mlp_backward on the tensor path (two-launch form): `event` is a cudaEvent_t that the next calls
 * record on their launch stream BETWEEN the dgrad chain kernel and the wgrad kernel, so that a caller timing the call with
 * its own events can split the two (bench.py's per-kernel roofline lines).  NULL (the default) disables it.  Process-global.  */
void snerf_set_backward_split_event(void* event);

/* Self-test of the tcgen05 GEMM building blocks against a CUDA-core GEMM (used by tests).
 * Returns SNERF_OK and writes the max abs error of each mode to host_max_err[4].               */
int snerf_tensor_selftest(float* host_max_err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLENERF_B200_H */
