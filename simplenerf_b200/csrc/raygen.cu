// Next row N1: per-frame ray construction and output post-processing on the device, so that a
// frame is one call: pose -> rays -> render -> uint8 image / clipped depth maps, without the H*W-ray host numpy pass and
// the all-keys .cpu().numpy() round trip of the reference (src/Tester01.py:57-66).
//
// Reference behaviour (src/data_preprocessors/DataPreprocessor01.py): get_rays :351-368, get_ndc_rays :371-389,
// get_view_dirs :392-394, post_process_image :1106-1109, post_process_depth :1112-1114.  The arithmetic follows numpy's
// fp32 operation order with explicitly rounded operations (no FMA contraction); one thread per pixel, coalesced 12-byte
// rows (the kernels are a few MB of traffic per frame and far from any roof: they exist to remove host work).
#include "common.cuh"

namespace snerf {

struct RayGenArgs {
    float pose[12];      // rows of [R | t]
    float kinv[9];       // inverse intrinsic, row major
    float s_w, s_h;      // -1 / (W / (2 fx)),  -1 / (H / (2 fy))      (:379-380)
    float near, two_near;
    int h, w, row0, n_rows, ndc;
    float *rays_o, *rays_d, *view_dirs, *rays_o_ndc, *rays_d_ndc;
};

__global__ void __launch_bounds__(256) raygen_kernel(const RayGenArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)a.n_rows * a.w;
    if (i >= n) return;
    const float x = (float)(int)(i % a.w), y = (float)(a.row0 + (int)(i / a.w));                    // :353-356
    float dir[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)                                                                         // :362  K^-1 [x y 1]^T
        dir[r] = __fadd_rn(__fadd_rn(__fmul_rn(a.kinv[3 * r], x), __fmul_rn(a.kinv[3 * r + 1], y)), a.kinv[3 * r + 2]);
    dir[1] = -dir[1];                                                                                    // :363
    dir[2] = -dir[2];
    float d[3], o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {                                                                        // :365, :367
        d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], a.pose[4 * r]), __fmul_rn(dir[1], a.pose[4 * r + 1])), __fmul_rn(dir[2], a.pose[4 * r + 2]));
        o[r] = a.pose[4 * r + 3];
    }
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));   // :393
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        a.rays_o[i * 3 + r] = o[r];
        a.rays_d[i * 3 + r] = d[r];
        a.view_dirs[i * 3 + r] = __fdiv_rn(d[r], nrm);
    }
    if (a.ndc) {
        const float t = __fdiv_rn(-__fadd_rn(a.near, o[2]), d[2]);                                       // :375
        float on[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) on[r] = __fadd_rn(o[r], __fmul_rn(t, d[r]));                         // :376
        a.rays_o_ndc[i * 3 + 0] = __fdiv_rn(__fmul_rn(a.s_w, on[0]), on[2]);                             // :379
        a.rays_o_ndc[i * 3 + 1] = __fdiv_rn(__fmul_rn(a.s_h, on[1]), on[2]);
        a.rays_o_ndc[i * 3 + 2] = __fadd_rn(1.f, __fdiv_rn(a.two_near, on[2]));
        a.rays_d_ndc[i * 3 + 0] = __fmul_rn(a.s_w, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(on[0], on[2])));   // :383
        a.rays_d_ndc[i * 3 + 1] = __fmul_rn(a.s_h, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(on[1], on[2])));
        a.rays_d_ndc[i * 3 + 2] = __fdiv_rn(-a.two_near, on[2]);
    }
}

// image = uint8(round_half_even(clip(rgb, 0, 1) * 255));  depth maps: clip(x, 0, inf) in place
struct PostArgs {
    const float* rgb;
    uint8_t* image;
    float* maps[4];
    int n_maps;
    long long n_pixels;
};

__global__ void __launch_bounds__(256) postprocess_kernel(const PostArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_pixels) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = fminf(fmaxf(a.rgb[i * 3 + c], 0.f), 1.f);                                        // numpy.clip propagates NaN; a NaN colour maps to 0 here
        a.image[i * 3 + c] = (uint8_t)__float2int_rn(__fmul_rn(v, 255.f));                               // numpy.round: half to even
    }
    for (int m = 0; m < a.n_maps; ++m) a.maps[m][i] = fmaxf(a.maps[m][i], 0.f);
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_generate_rays(const float* pose34, const float* kinv33, float s_w, float s_h, float near, float two_near,
                                   int h, int w, int row0, int n_rows, int ndc, float* rays_o, float* rays_d, float* view_dirs,
                                   float* rays_o_ndc, float* rays_d_ndc, void* stream) {
    SNERF_REQUIRE(pose34 && kinv33, "snerf_generate_rays: null camera");
    SNERF_REQUIRE(h >= 1 && w >= 1 && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= h, "snerf_generate_rays: bad row band (%d + %d of %d)", row0, n_rows, h);
    if (n_rows == 0) return SNERF_OK;
    SNERF_REQUIRE(rays_o && rays_d && view_dirs && (!ndc || (rays_o_ndc && rays_d_ndc)), "snerf_generate_rays: null output");
    RayGenArgs a{};
    for (int i = 0; i < 12; ++i) a.pose[i] = pose34[i];
    for (int i = 0; i < 9; ++i) a.kinv[i] = kinv33[i];
    a.s_w = s_w; a.s_h = s_h; a.near = near; a.two_near = two_near;
    a.h = h; a.w = w; a.row0 = row0; a.n_rows = n_rows; a.ndc = ndc;
    a.rays_o = rays_o; a.rays_d = rays_d; a.view_dirs = view_dirs; a.rays_o_ndc = rays_o_ndc; a.rays_d_ndc = rays_d_ndc;
    const long long n = (long long)n_rows * w;
    raygen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    SNERF_LAUNCH_OK("raygen_kernel");
    return SNERF_OK;
}

extern "C" int snerf_postprocess_frame(const float* rgb, uint8_t* image, float* const* depth_maps, int n_maps, long long n_pixels,
                                       void* stream) {
    SNERF_REQUIRE(n_pixels >= 0 && n_maps >= 0 && n_maps <= 4, "snerf_postprocess_frame: bad sizes");
    if (n_pixels == 0) return SNERF_OK;
    SNERF_REQUIRE(rgb && image && (n_maps == 0 || depth_maps), "snerf_postprocess_frame: null pointer");
    PostArgs a{};
    a.rgb = rgb; a.image = image; a.n_maps = n_maps; a.n_pixels = n_pixels;
    for (int m = 0; m < n_maps; ++m) {
        SNERF_REQUIRE(depth_maps[m] != nullptr, "snerf_postprocess_frame: map %d is null", m);
        a.maps[m] = depth_maps[m];
    }
    postprocess_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    SNERF_LAUNCH_OK("postprocess_kernel");
    return SNERF_OK;
}
