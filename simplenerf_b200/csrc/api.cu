// C-ABI glue: error reporting, descriptor validation and dispatch of the MLP entry points to the
// tensor (tcgen05, bf16) or the precise (CUDA-core, fp32) implementation.
#include <stdarg.h>

#include "common.cuh"

namespace snerf {

std::string& last_error() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int validate_desc(const snerf_mlp_desc* d) {
    SNERF_REQUIRE(d != nullptr, "mlp desc is null");
    if (d->depth != 8 || d->width != 256 || d->skip_layer != 4)
        return fail(SNERF_ERR_UNSUPPORTED, "mlp desc: only the reference trunk (depth 8, width 256, skip after layer 4) is built; got %d/%d/%d",
                    d->depth, d->width, d->skip_layer);
    SNERF_REQUIRE(d->pts_degree >= 1 && d->pts_degree <= 10, "mlp desc: pts_degree %d outside [1,10]", d->pts_degree);
    SNERF_REQUIRE(d->trunk_degree >= 0 && d->trunk_degree <= d->pts_degree, "mlp desc: trunk_degree %d outside [0,pts_degree]", d->trunk_degree);
    SNERF_REQUIRE(d->view_degree >= 0 && d->view_degree <= 4, "mlp desc: view_degree %d outside [0,4]", d->view_degree);
    if (d->view_width != 0 && d->view_width != 128)
        return fail(SNERF_ERR_UNSUPPORTED, "mlp desc: view_width must be 0 or 128, got %d", d->view_width);
    SNERF_REQUIRE((d->view_width > 0 && d->head_out == 1) || (d->view_width == 0 && d->head_out == 4),
                  "mlp desc: head_out %d inconsistent with view_width %d", d->head_out, d->view_width);
    SNERF_REQUIRE(d->view_width > 0 || d->view_degree == 0, "mlp desc: view_degree without a view branch");
    return SNERF_OK;
}

static int check_params(const snerf_mlp_desc& d, const void* const* p, const char* what) {
    SNERF_REQUIRE(p != nullptr, "%s: null pointer table", what);
    for (int i = 0; i < 2 * d.depth; ++i) SNERF_REQUIRE(p[i] != nullptr, "%s: trunk entry %d is null", what, i);
    SNERF_REQUIRE(p[SNERF_P_HEAD_W] && p[SNERF_P_HEAD_B], "%s: pts_output_linear is null", what);
    if (d.view_width > 0)
        for (int i = SNERF_P_FEAT_W; i <= SNERF_P_RGB_B; ++i) SNERF_REQUIRE(p[i] != nullptr, "%s: view-branch entry %d is null", what, i);
    return SNERF_OK;
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_abi_version(void) { return SNERF_ABI_VERSION; }
extern "C" const char* snerf_last_error(void) { return last_error().c_str(); }

extern "C" size_t snerf_mlp_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, uint32_t flags) {
    if (validate_desc(desc) != SNERF_OK || n_rays < 0 || n_samples < 1) return 0;
    const MlpDims m(*desc);
    const size_t b = (flags & SNERF_FLAG_PRECISE) ? simt_workspace_bytes(m, *desc, n_rays, n_samples, flags)
                                                  : tc_workspace_bytes(m, *desc, n_rays, n_samples, flags);
    return b < 256 ? 256 : b;
}

extern "C" size_t snerf_packed_weights_bytes(const snerf_mlp_desc* desc) {
    if (validate_desc(desc) != SNERF_OK) return 0;
    return tc_packed_bytes(*desc);
}

extern "C" int snerf_pack_weights(const snerf_mlp_desc* desc, const float* const* host_params, void* packed, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_pack_weights");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(packed != nullptr, "snerf_pack_weights: packed is null");
    return tc_pack(*desc, host_params, packed, (cudaStream_t)stream);
}

extern "C" int snerf_mlp_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                                 const float* rays_o, const float* rays_d, const float* view_dirs, const float* z,
                                 const float* sigma_noise, float* sigma, float* rgb, void* workspace,
                                 size_t workspace_bytes, int n_rays, int n_samples, uint32_t flags, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_mlp_forward");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_mlp_forward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(rays_o && rays_d && z && sigma && rgb && workspace, "snerf_mlp_forward: null pointer");
    SNERF_REQUIRE(desc->view_degree == 0 || view_dirs != nullptr, "snerf_mlp_forward: view_dirs required by this MLP");
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_mlp_forward: bad sizes");
    SNERF_REQUIRE((long long)n_rays * n_samples < (1LL << 31), "snerf_mlp_forward: more than 2^31 points in one call");
    if (flags & SNERF_FLAG_PRECISE)
        return simt_forward(*desc, host_params, rays_o, rays_d, view_dirs, z, sigma_noise, sigma, rgb, workspace,
                            workspace_bytes, n_rays, n_samples, flags, (cudaStream_t)stream);
    SNERF_REQUIRE(packed != nullptr, "snerf_mlp_forward: the tensor path needs packed weights (snerf_pack_weights)");
    return tc_forward(*desc, host_params, packed, rays_o, rays_d, view_dirs, z, sigma_noise, sigma, rgb, workspace,
                      workspace_bytes, n_rays, n_samples, flags, (cudaStream_t)stream);
}

extern "C" int snerf_mlp_forward_rng(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                                     const float* rays_o, const float* rays_d, const float* view_dirs, const float* z,
                                     float noise_std, uint64_t seed, uint64_t offset, float* sigma, float* rgb, void* workspace,
                                     size_t workspace_bytes, int n_rays, int n_samples, uint32_t flags, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_mlp_forward_rng");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_mlp_forward_rng: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    if (flags & SNERF_FLAG_PRECISE)
        return fail(SNERF_ERR_UNSUPPORTED, "snerf_mlp_forward_rng: tensor path only (precise path: snerf_fill_random + snerf_mlp_forward draw the same numbers)");
    SNERF_REQUIRE(rays_o && rays_d && z && sigma && rgb && workspace && packed, "snerf_mlp_forward_rng: null pointer");
    SNERF_REQUIRE(desc->view_degree == 0 || view_dirs != nullptr, "snerf_mlp_forward_rng: view_dirs required by this MLP");
    SNERF_REQUIRE((long long)n_rays * n_samples < (1LL << 31), "snerf_mlp_forward_rng: more than 2^31 points in one call");
    const unsigned long long key[2] = {seed, offset};
    return tc_forward(*desc, host_params, packed, rays_o, rays_d, view_dirs, z, nullptr, sigma, rgb, workspace, workspace_bytes,
                      n_rays, n_samples, flags, (cudaStream_t)stream, key, noise_std);
}

extern "C" int snerf_mlp_backward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                                  const float* rays_o, const float* rays_d, const float* view_dirs, const float* z,
                                  const float* sigma, const float* rgb, const float* d_sigma, const float* d_rgb,
                                  float* const* host_grads, void* workspace, size_t workspace_bytes, int n_rays,
                                  int n_samples, uint32_t flags, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_mlp_backward(params)");
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_grads, "snerf_mlp_backward(grads)");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_mlp_backward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(sigma && rgb && d_sigma && d_rgb && workspace, "snerf_mlp_backward: null pointer");
    SNERF_REQUIRE(flags & SNERF_FLAG_SAVE_FOR_BWD, "snerf_mlp_backward: forward must have run with SNERF_FLAG_SAVE_FOR_BWD");
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_mlp_backward: bad sizes");
    if (flags & SNERF_FLAG_PRECISE)
        return simt_backward(*desc, host_params, sigma, rgb, d_sigma, d_rgb, host_grads, workspace, workspace_bytes,
                             n_rays, n_samples, flags, (cudaStream_t)stream);
    SNERF_REQUIRE(packed != nullptr && rays_o && rays_d && z, "snerf_mlp_backward: the tensor path needs packed weights and the rays");
    return tc_backward(*desc, host_params, packed, rays_o, rays_d, view_dirs, z, sigma, rgb, d_sigma, d_rgb, host_grads,
                       workspace, workspace_bytes, n_rays, n_samples, flags, (cudaStream_t)stream);
}

extern "C" size_t snerf_render_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, uint32_t flags) {
    if (validate_desc(desc) != SNERF_OK || n_rays < 0 || n_samples < 1 || (flags & SNERF_FLAG_PRECISE)) return 0;
    return tc_render_workspace_bytes(MlpDims(*desc), *desc, n_rays, n_samples);
}

extern "C" int snerf_render_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* packed,
                                    const float* pts_o, const float* pts_d, const float* view_dirs, const float* z,
                                    const float* rays_o, const float* rays_d, float* rgb_map, float* acc, float* depth,
                                    float* depth_var, float* depth_ndc, float* depth_var_ndc, float* alpha, float* weights,
                                    void* workspace, size_t workspace_bytes, int n_rays, int n_samples, uint32_t flags, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_render_forward");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_render_forward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    if (flags & (SNERF_FLAG_PRECISE | SNERF_FLAG_SAVE_FOR_BWD))
        return fail(SNERF_ERR_UNSUPPORTED, "snerf_render_forward: evaluation on the tensor path only (training keeps sigma / rgb for the backward pass)");
    const bool ndc = (flags & SNERF_FLAG_NDC) != 0;
    SNERF_REQUIRE(packed && pts_o && pts_d && z && workspace, "snerf_render_forward: null input");
    SNERF_REQUIRE(desc->view_degree == 0 || view_dirs != nullptr, "snerf_render_forward: view_dirs required by this MLP");
    SNERF_REQUIRE(rgb_map && acc && depth && depth_var, "snerf_render_forward: null per-ray output");
    SNERF_REQUIRE(!ndc || (rays_o && rays_d && depth_ndc && depth_var_ndc), "snerf_render_forward: NDC mode needs rays_o, rays_d, depth_ndc, depth_var_ndc");
    SNERF_REQUIRE((long long)n_rays * n_samples < (1LL << 31), "snerf_render_forward: more than 2^31 points in one call");
    FusedComposite fc{};
    fc.rays_o = rays_o; fc.rays_d = rays_d; fc.rgb_map = rgb_map; fc.acc = acc; fc.depth = depth; fc.depth_var = depth_var;
    fc.depth_ndc = depth_ndc; fc.depth_var_ndc = depth_var_ndc; fc.alpha = alpha; fc.weights = weights;
    fc.ndc = ndc; fc.white = (flags & SNERF_FLAG_WHITE_BKGD) != 0;
    return tc_render_forward(*desc, host_params, packed, pts_o, pts_d, view_dirs, z, fc, workspace, workspace_bytes, n_rays, n_samples,
                             (cudaStream_t)stream);
}

extern "C" size_t snerf_visibility_workspace_bytes(const snerf_mlp_desc* desc, int n_rays, int n_samples, int n_other) {
    if (validate_desc(desc) != SNERF_OK || n_rays < 0 || n_samples < 1 || n_other < 0 || desc->view_width <= 0) return 0;
    const size_t b = simt_visibility_workspace_bytes(MlpDims(*desc), n_rays, n_samples, n_other);
    return b < 256 ? 256 : b;
}

extern "C" int snerf_visibility_forward(const snerf_mlp_desc* desc, const float* const* host_params, const void* mlp_workspace,
                                        const float* rays_o, const float* rays_d, const float* z, const float* rays_o2,
                                        float* visibility, float* visibility2, void* workspace, size_t workspace_bytes,
                                        int n_rays, int n_samples, int n_other, uint32_t flags, void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_visibility_forward");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_other >= 0, "snerf_visibility_forward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(desc->view_width > 0 && desc->view_degree > 0, "snerf_visibility_forward: the MLP has no view branch");
    SNERF_REQUIRE(mlp_workspace && rays_o && rays_d && z && visibility && workspace, "snerf_visibility_forward: null pointer");
    SNERF_REQUIRE(n_other == 0 || (rays_o2 && visibility2), "snerf_visibility_forward: rays_o2 / visibility2 needed for %d other views", n_other);
    if (!(flags & SNERF_FLAG_PRECISE))      // tensor path: everything it needs sits in the MLP workspace (SNERF_FLAG_VIS_HEAD)
        return tc_visibility_forward(*desc, host_params, mlp_workspace, rays_o, rays_d, z, rays_o2, visibility, visibility2, n_rays,
                                     n_samples, n_other, flags, (cudaStream_t)stream);
    return simt_visibility_forward(*desc, host_params, mlp_workspace, rays_o, rays_d, z, rays_o2, visibility, visibility2, workspace,
                                   workspace_bytes, n_rays, n_samples, n_other, flags, (cudaStream_t)stream);
}

extern "C" int snerf_visibility_backward(const snerf_mlp_desc* desc, const float* const* host_params, void* mlp_workspace,
                                         const float* rays_o, const float* rays_d, const float* z, const float* rays_o2,
                                         const float* visibility, const float* visibility2, const float* d_visibility,
                                         const float* d_visibility2, float* const* host_grads, void* workspace,
                                         size_t workspace_bytes, int n_rays, int n_samples, int n_other, uint32_t flags,
                                         void* stream) {
    int rc = validate_desc(desc);
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_params, "snerf_visibility_backward(params)");
    if (rc != SNERF_OK) return rc;
    rc = check_params(*desc, (const void* const*)host_grads, "snerf_visibility_backward(grads)");
    if (rc != SNERF_OK) return rc;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_other >= 0, "snerf_visibility_backward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(flags & SNERF_FLAG_SAVE_FOR_BWD, "snerf_visibility_backward: the forward must have run with SNERF_FLAG_SAVE_FOR_BWD");
    SNERF_REQUIRE(desc->view_width > 0 && desc->view_degree > 0, "snerf_visibility_backward: the MLP has no view branch");
    SNERF_REQUIRE(mlp_workspace && rays_o && rays_d && z && visibility && workspace, "snerf_visibility_backward: null pointer");
    SNERF_REQUIRE(n_other == 0 || (rays_o2 && visibility2), "snerf_visibility_backward: rays_o2 / visibility2 needed");
    if (!(flags & SNERF_FLAG_PRECISE))
        return tc_visibility_backward(*desc, host_params, mlp_workspace, rays_o, rays_d, z, rays_o2, visibility, visibility2, d_visibility,
                                      d_visibility2, host_grads, n_rays, n_samples, n_other, flags, (cudaStream_t)stream);
    return simt_visibility_backward(*desc, host_params, mlp_workspace, rays_o, rays_d, z, rays_o2, visibility, visibility2, d_visibility,
                                    d_visibility2, host_grads, workspace, workspace_bytes, n_rays, n_samples, n_other, flags,
                                    (cudaStream_t)stream);
}

extern "C" int snerf_tensor_selftest(float* host_max_err, void* stream) {
    SNERF_REQUIRE(host_max_err != nullptr, "snerf_tensor_selftest: null output");
    return tc_selftest(host_max_err, (cudaStream_t)stream);
}
