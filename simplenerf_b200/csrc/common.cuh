// Shared helpers for the simplenerf_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/simplenerf_b200.h"

namespace snerf {

// thread-local error text returned by snerf_last_error()
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define SNERF_REQUIRE(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) return ::snerf::fail(SNERF_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define SNERF_CUDA_OK(expr)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return ::snerf::fail(SNERF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                                 __FILE__, __LINE__);                                               \
    } while (0)

#define SNERF_LAUNCH_OK(name)                                                                     \
    do {                                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess)                                                                   \
            return ::snerf::fail(SNERF_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Encoded width of a positional encoding with `degree` bands on 3 inputs
__host__ __device__ inline int pe_dim(int degree) { return 3 * (1 + 2 * degree); }

// Derived shapes of one MLP (mirrors oracle.MlpSpec / reference MLP ctor :561-609)
struct MlpDims {
    int depth, width, skip_layer;
    int enc;        // full points encoding width (63)
    int trunk_in;   // columns of the encoding fed to the trunk (63 or 21)
    int enc_hi;     // enc - trunk_in: columns appended to the view-layer input (0 or 42)
    int venc;       // view-dir encoding width (27 or 0)
    int view_width; // 128 or 0
    int view_in;    // width + enc_hi + venc (283 / 325) or 0
    int head_out;   // 1 or 4
    bool has_view;
    __host__ explicit MlpDims(const snerf_mlp_desc& d)
        : depth(d.depth), width(d.width), skip_layer(d.skip_layer), enc(pe_dim(d.pts_degree)),
          trunk_in(pe_dim(d.trunk_degree)), enc_hi(pe_dim(d.pts_degree) - pe_dim(d.trunk_degree)),
          venc(d.view_degree > 0 ? pe_dim(d.view_degree) : 0), view_width(d.view_width),
          view_in(d.view_width > 0 ? d.width + (pe_dim(d.pts_degree) - pe_dim(d.trunk_degree)) +
                                        (d.view_degree > 0 ? pe_dim(d.view_degree) : 0)
                                  : 0),
          head_out(d.head_out), has_view(d.view_width > 0) {}
    __host__ int trunk_fan_in(int layer) const {
        if (layer == 0) return trunk_in;
        return width + ((layer - 1) == skip_layer ? trunk_in : 0);
    }
};

int validate_desc(const snerf_mlp_desc* desc);

// ---- entry points implemented per source file -------------------------------------------------
// mlp_simt.cu : fp32 CUDA-core MLP (precise path)
size_t simt_workspace_bytes(const MlpDims& m, const snerf_mlp_desc& d, int n_rays, int n_samples, uint32_t flags);
int simt_forward(const snerf_mlp_desc& d, const float* const* prm, const float* rays_o, const float* rays_d,
                 const float* view_dirs, const float* z, const float* noise, float* sigma, float* rgb, void* ws,
                 size_t ws_bytes, int n_rays, int n_samples, uint32_t flags, cudaStream_t st);
int simt_backward(const snerf_mlp_desc& d, const float* const* prm, const float* sigma, const float* rgb,
                  const float* d_sigma, const float* d_rgb, float* const* grads, void* ws, size_t ws_bytes,
                  int n_rays, int n_samples, uint32_t flags, cudaStream_t st);

size_t simt_visibility_workspace_bytes(const MlpDims& m, int n_rays, int n_samples, int n_other);
int simt_visibility_forward(const snerf_mlp_desc& d, const float* const* prm, const void* mlp_ws, const float* rays_o,
                            const float* rays_d, const float* z, const float* rays_o2, float* visibility, float* visibility2,
                            void* vis_ws, size_t vis_ws_bytes, int n_rays, int n_samples, int n_other, uint32_t flags,
                            cudaStream_t st);
int simt_visibility_backward(const snerf_mlp_desc& d, const float* const* prm, void* mlp_ws, const float* rays_o,
                             const float* rays_d, const float* z, const float* rays_o2, const float* visibility,
                             const float* visibility2, const float* d_visibility, const float* d_visibility2,
                             float* const* grads, void* vis_ws, size_t vis_ws_bytes, int n_rays, int n_samples, int n_other,
                             uint32_t flags, cudaStream_t st);

// mlp_tc.cu : bf16 tcgen05 MLP (tensor path)
size_t tc_workspace_bytes(const MlpDims& m, const snerf_mlp_desc& d, int n_rays, int n_samples, uint32_t flags);
size_t tc_packed_bytes(const snerf_mlp_desc& d);
int tc_pack(const snerf_mlp_desc& d, const float* const* prm, void* packed, cudaStream_t st);
struct FusedRun {                     // fused evaluation: where the forward kernel leaves the per-run records (mlp_tc.cu)
    float *seg, *alpha, *wloc;
    const float *cam_o, *cam_d;
};
int tc_forward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* rays_o,
               const float* rays_d, const float* view_dirs, const float* z, const float* noise, float* sigma,
               float* rgb, void* ws, size_t ws_bytes, int n_rays, int n_samples, uint32_t flags, cudaStream_t st,
               const unsigned long long* rng_seed_offset = nullptr, float noise_std = 0.f, const FusedRun* fused = nullptr);
int tc_backward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* rays_o,
                const float* rays_d, const float* view_dirs, const float* z, const float* sigma, const float* rgb,
                const float* d_sigma, const float* d_rgb, float* const* grads, void* ws, size_t ws_bytes,
                int n_rays, int n_samples, uint32_t flags, cudaStream_t st);
int tc_selftest(float* host_max_err, cudaStream_t st);
// vis_tc.cu : visibility head on the workspace of a tensor-path forward with SNERF_FLAG_VIS_HEAD
int tc_visibility_forward(const snerf_mlp_desc& d, const float* const* prm, const void* mlp_ws, const float* rays_o,
                          const float* rays_d, const float* z, const float* rays_o2, float* visibility, float* visibility2,
                          int n_rays, int n_samples, int n_other, uint32_t flags, cudaStream_t st);
int tc_visibility_backward(const snerf_mlp_desc& d, const float* const* prm, void* mlp_ws, const float* rays_o,
                           const float* rays_d, const float* z, const float* rays_o2, const float* visibility,
                           const float* visibility2, const float* d_visibility, const float* d_visibility2,
                           float* const* grads, int n_rays, int n_samples, int n_other, uint32_t flags, cudaStream_t st);
// fused evaluation (row X1): MLP forward with the compositing arithmetic in its head epilogue + the per-ray fold
struct FusedComposite {
    const float *rays_o, *rays_d;     // camera rays (NDC depth conversion)
    float *rgb_map, *acc, *depth, *depth_var, *depth_ndc, *depth_var_ndc, *alpha, *weights;
    bool ndc, white;
};
size_t tc_render_workspace_bytes(const MlpDims& m, const snerf_mlp_desc& d, int n_rays, int n_samples);
int tc_render_forward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* pts_o,
                      const float* pts_d, const float* view_dirs, const float* z, const FusedComposite& fc, void* ws,
                      size_t ws_bytes, int n_rays, int n_samples, cudaStream_t st);

}  // namespace snerf
