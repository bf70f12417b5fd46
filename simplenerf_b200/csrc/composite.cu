// Alpha compositing, forward and backward: one warp per ray, shuffle scans, fp32.
//
// Reference behaviour (src/models/SimpleNeRF01.py): volume_rendering :430-483 and
// convert_depth_from_ndc :485-502.  Lane l of the warp owns samples l, l+32, l+64, ... so that every
// global access of sigma / z / per-sample outputs is a coalesced 128-byte row; the transmittance is an
// exclusive product scan (5 shuffle steps per 32-sample chunk plus a running carry).
#include "common.cuh"

namespace snerf {

constexpr int kCompWarps = 8;

struct CompositeArgs {
    const float *sigma, *rgb, *z, *rays_o, *rays_d, *rays_d_ndc;
    // forward outputs (nullable per-sample maps)
    float *rgb_map, *acc, *depth, *depth_var, *depth_ndc, *depth_var_ndc, *alpha, *vis, *weights;
    // backward inputs (all nullable) and outputs
    const float *g_rgb_map, *g_acc, *g_depth, *g_depth_var, *g_depth_ndc, *g_depth_var_ndc, *g_alpha, *g_vis, *g_weights;
    float *d_sigma, *d_rgb;
    int n_rays, s;
    bool ndc, white;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Per-ray constants and the per-sample quantities every pass needs.
template <int ITEMS>
struct RayState {
    float sig[ITEMS], zz[ITEMS], zm[ITEMS];   // sigma, z (ndc or metric), metric z
    float delta[ITEMS], alpha[ITEMS], trans[ITEMS], w[ITEMS];
    float acc, dsum, dsum_ndc;   // sum w, sum w*zm, sum w*z_ndc
};

template <int ITEMS>
__device__ __forceinline__ void load_and_scan(const CompositeArgs& a, int ray, int lane, RayState<ITEMS>& r) {
    const int s = a.s;
    const float* sig = a.sigma + (size_t)ray * s;
    const float* z = a.z + (size_t)ray * s;
    const float* dvec = (a.ndc ? a.rays_d_ndc : a.rays_d) + (size_t)ray * 3;
    const float dn = sqrtf(dvec[0] * dvec[0] + dvec[1] * dvec[1] + dvec[2] * dvec[2]);          // :436 / :441
    const float tail = a.ndc ? 1.f : 1e10f;                                                     // :433 / :438
    float oz = 0.f, dz = 1.f, tn = 0.f;
    if (a.ndc) {
        oz = a.rays_o[(size_t)ray * 3 + 2];
        dz = a.rays_d[(size_t)ray * 3 + 2];
        tn = -(1.f + oz) / dz;                                                                  // :498
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int k = i * kWarp + lane;
        r.sig[i] = 0.f;
        r.zz[i] = 0.f;
        float znext = 0.f;
        if (k < s) {
            r.sig[i] = sig[k];
            r.zz[i] = z[k];
            znext = (k + 1 < s) ? z[k + 1] : tail;
        }
        r.delta[i] = (znext - r.zz[i]) * dn;                                                    // :435-436
        r.alpha[i] = (k < s) ? 1.f - expf(-r.sig[i] * r.delta[i]) : 0.f;                        // :446
        if (a.ndc) {
            const float guard = (r.zz[i] == 1.f) ? 1e-3f : 0.f;                                 // :499
            r.zm[i] = (oz + tn * dz) / dz * (1.f / (1.f - r.zz[i] + guard) - 1.f) + tn;         // :501
        } else {
            r.zm[i] = r.zz[i];
        }
    }
    // exclusive product scan of (1 - alpha + 1e-10)                                            // :447
    float carry = 1.f;
    r.acc = r.dsum = r.dsum_ndc = 0.f;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int k = i * kWarp + lane;
        const float f = (k < s) ? (1.f - r.alpha[i]) + 1e-10f : 1.f;
        float incl = f;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
            const float v = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl *= v;
        }
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.f;
        r.trans[i] = carry * excl;
        carry *= __shfl_sync(kFull, incl, kWarp - 1);
        r.w[i] = r.alpha[i] * r.trans[i];                                                       // :448
        r.acc += r.w[i];
        r.dsum += r.w[i] * r.zm[i];
        r.dsum_ndc += r.w[i] * r.zz[i];
    }
    r.acc = warp_sum(r.acc);                                                                    // :451
    r.dsum = warp_sum(r.dsum);
    r.dsum_ndc = warp_sum(r.dsum_ndc);
}

template <int ITEMS>
__global__ void __launch_bounds__(kCompWarps* kWarp) composite_fwd_kernel(const CompositeArgs a) {
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int ray = blockIdx.x * kCompWarps + warp;
    if (ray >= a.n_rays) return;
    const int s = a.s;
    RayState<ITEMS> r;
    load_and_scan<ITEMS>(a, ray, lane, r);

    const float* rgb = a.rgb + (size_t)ray * s * 3;
    float cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int k = i * kWarp + lane;
        if (k < s) {
            cr += r.w[i] * rgb[k * 3 + 0];                                                      // :449
            cg += r.w[i] * rgb[k * 3 + 1];
            cb += r.w[i] * rgb[k * 3 + 2];
            const size_t o = (size_t)ray * s + k;
            if (a.alpha) a.alpha[o] = r.alpha[i];
            if (a.vis) a.vis[o] = r.trans[i];
            if (a.weights) a.weights[o] = r.w[i];
        }
    }
    cr = warp_sum(cr);
    cg = warp_sum(cg);
    cb = warp_sum(cb);
    const float inv = 1.f / (r.acc + 1e-6f);
    const float depth = r.dsum * inv;                                                           // :453 / :459
    const float depth_ndc = r.dsum_ndc * inv;                                                   // :456
    float var = 0.f, var_ndc = 0.f;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const float e = r.zm[i] - depth, en = r.zz[i] - depth_ndc;
        var += r.w[i] * e * e;                                                                  // :454 / :460
        var_ndc += r.w[i] * en * en;                                                            // :457
    }
    var = warp_sum(var);
    var_ndc = warp_sum(var_ndc);
    if (lane == 0) {
        const float bg = a.white ? 1.f - r.acc : 0.f;                                           // :463
        a.rgb_map[(size_t)ray * 3 + 0] = cr + bg;
        a.rgb_map[(size_t)ray * 3 + 1] = cg + bg;
        a.rgb_map[(size_t)ray * 3 + 2] = cb + bg;
        a.acc[ray] = r.acc;
        a.depth[ray] = depth;
        a.depth_var[ray] = var;
        if (a.ndc) {
            a.depth_ndc[ray] = depth_ndc;
            a.depth_var_ndc[ray] = var_ndc;
        }
    }
}

// Backward.  With f_k = 1 - alpha_k + 1e-10, T_k = prod_{j<k} f_j, w_k = alpha_k T_k:
//   g_k   = dL/dw_k  (collected from every per-ray map plus d_weights)
//   G_k   = g_k alpha_k + dL/dT_k
//   dL/dalpha_k = g_k T_k + d_alpha_k - (sum_{j>k} G_j T_j) / f_k
//   dL/dsigma_k = dL/dalpha_k * delta_k * (1 - alpha_k)
template <int ITEMS>
__global__ void __launch_bounds__(kCompWarps* kWarp) composite_bwd_kernel(const CompositeArgs a) {
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int ray = blockIdx.x * kCompWarps + warp;
    if (ray >= a.n_rays) return;
    const int s = a.s;
    RayState<ITEMS> r;
    load_and_scan<ITEMS>(a, ray, lane, r);

    const float inv = 1.f / (r.acc + 1e-6f);
    const float depth = r.dsum * inv, depth_ndc = r.dsum_ndc * inv;
    float gr = 0.f, gg = 0.f, gb = 0.f;
    if (a.g_rgb_map) {
        gr = a.g_rgb_map[(size_t)ray * 3 + 0];
        gg = a.g_rgb_map[(size_t)ray * 3 + 1];
        gb = a.g_rgb_map[(size_t)ray * 3 + 2];
    }
    float g_const = a.g_acc ? a.g_acc[ray] : 0.f;
    if (a.white) g_const -= gr + gg + gb;
    const float gd = a.g_depth ? a.g_depth[ray] : 0.f;
    const float gdn = (a.ndc && a.g_depth_ndc) ? a.g_depth_ndc[ray] : 0.f;
    const float gv = a.g_depth_var ? a.g_depth_var[ray] : 0.f;
    const float gvn = (a.ndc && a.g_depth_var_ndc) ? a.g_depth_var_ndc[ray] : 0.f;
    // sum_j w_j (z_j - depth) = D - depth * acc  (tiny, but kept exact)
    const float resid = r.dsum - depth * r.acc, resid_ndc = r.dsum_ndc - depth_ndc * r.acc;

    const float* rgb = a.rgb + (size_t)ray * s * 3;
    float g[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int k = i * kWarp + lane;
        g[i] = 0.f;
        if (k < s) {
            const size_t o = (size_t)ray * s + k;
            const float c0 = rgb[k * 3 + 0], c1 = rgb[k * 3 + 1], c2 = rgb[k * 3 + 2];
            const float e = r.zm[i] - depth, en = r.zz[i] - depth_ndc;
            float gi = gr * c0 + gg * c1 + gb * c2 + g_const;
            gi += gd * e * inv + gdn * en * inv;
            gi += gv * (e * e - 2.f * e * inv * resid) + gvn * (en * en - 2.f * en * inv * resid_ndc);
            if (a.g_weights) gi += a.g_weights[o];
            g[i] = gi;
            a.d_rgb[o * 3 + 0] = r.w[i] * gr;
            a.d_rgb[o * 3 + 1] = r.w[i] * gg;
            a.d_rgb[o * 3 + 2] = r.w[i] * gb;
        }
    }
    // reverse exclusive scan of G_k T_k
    float carry = 0.f;
#pragma unroll
    for (int i = ITEMS - 1; i >= 0; --i) {
        const int k = i * kWarp + lane;
        const size_t o = (size_t)ray * s + k;
        float gt = 0.f;
        if (k < s) {
            float big_g = g[i] * r.alpha[i];
            if (a.g_vis) big_g += a.g_vis[o];
            gt = big_g * r.trans[i];
        }
        float incl = gt;
#pragma unroll
        for (int off = 1; off < kWarp; off <<= 1) {
            const float v = __shfl_down_sync(kFull, incl, off);
            if (lane + off < kWarp) incl += v;
        }
        const float suffix = carry + (incl - gt);   // sum over j > k
        carry += __shfl_sync(kFull, incl, 0);
        if (k < s) {
            const float f = (1.f - r.alpha[i]) + 1e-10f;
            float d_alpha = g[i] * r.trans[i] - suffix / f;
            if (a.g_alpha) d_alpha += a.g_alpha[o];
            a.d_sigma[o] = d_alpha * r.delta[i] * (1.f - r.alpha[i]);
        }
    }
}

template <int ITEMS>
static int launch_composite(const CompositeArgs& a, bool backward, cudaStream_t st) {
    const int blocks = ceil_div(a.n_rays, kCompWarps);
    if (backward)
        composite_bwd_kernel<ITEMS><<<blocks, kCompWarps * kWarp, 0, st>>>(a);
    else
        composite_fwd_kernel<ITEMS><<<blocks, kCompWarps * kWarp, 0, st>>>(a);
    SNERF_LAUNCH_OK(backward ? "composite_bwd_kernel" : "composite_fwd_kernel");
    return SNERF_OK;
}

static int dispatch_composite(const CompositeArgs& a, bool backward, cudaStream_t st) {
    if (a.n_rays == 0) return SNERF_OK;
    const int items = ceil_div(a.s, kWarp);
    switch (items) {
        case 1: return launch_composite<1>(a, backward, st);
        case 2: return launch_composite<2>(a, backward, st);
        case 3: return launch_composite<3>(a, backward, st);
        case 4: return launch_composite<4>(a, backward, st);
        case 5: case 6: return launch_composite<6>(a, backward, st);
        case 7: case 8: return launch_composite<8>(a, backward, st);
        default: break;
    }
    if (items <= 16) return launch_composite<16>(a, backward, st);
    return fail(SNERF_ERR_UNSUPPORTED, "composite: %d samples per ray > 512", a.s);
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_composite_forward(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                                       const float* rays_d, const float* rays_d_ndc, float* rgb_map, float* acc,
                                       float* depth, float* depth_var, float* depth_ndc, float* depth_var_ndc,
                                       float* alpha, float* visibility, float* weights, int n_rays, int n_samples,
                                       uint32_t flags, void* stream) {
    const bool ndc = (flags & SNERF_FLAG_NDC) != 0;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_composite_forward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(sigma && rgb && z && rays_d, "snerf_composite_forward: null input");
    SNERF_REQUIRE(rgb_map && acc && depth && depth_var, "snerf_composite_forward: null per-ray output");
    SNERF_REQUIRE(!ndc || (rays_o && rays_d_ndc && depth_ndc && depth_var_ndc),
                  "snerf_composite_forward: NDC mode needs rays_o, rays_d_ndc, depth_ndc, depth_var_ndc");
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_composite_forward: bad sizes");
    CompositeArgs a{};
    a.sigma = sigma; a.rgb = rgb; a.z = z; a.rays_o = rays_o; a.rays_d = rays_d; a.rays_d_ndc = rays_d_ndc;
    a.rgb_map = rgb_map; a.acc = acc; a.depth = depth; a.depth_var = depth_var; a.depth_ndc = depth_ndc;
    a.depth_var_ndc = depth_var_ndc; a.alpha = alpha; a.vis = visibility; a.weights = weights;
    a.n_rays = n_rays; a.s = n_samples; a.ndc = ndc; a.white = (flags & SNERF_FLAG_WHITE_BKGD) != 0;
    return dispatch_composite(a, false, (cudaStream_t)stream);
}

extern "C" int snerf_composite_backward(const float* sigma, const float* rgb, const float* z, const float* rays_o,
                                        const float* rays_d, const float* rays_d_ndc, const float* d_rgb_map,
                                        const float* d_acc, const float* d_depth, const float* d_depth_var,
                                        const float* d_depth_ndc, const float* d_depth_var_ndc, const float* d_alpha,
                                        const float* d_visibility, const float* d_weights, float* d_sigma,
                                        float* d_rgb, int n_rays, int n_samples, uint32_t flags, void* stream) {
    const bool ndc = (flags & SNERF_FLAG_NDC) != 0;
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_composite_backward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(sigma && rgb && z && rays_d && d_sigma && d_rgb, "snerf_composite_backward: null pointer");
    SNERF_REQUIRE(!ndc || (rays_o && rays_d_ndc), "snerf_composite_backward: NDC mode needs rays_o and rays_d_ndc");
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_composite_backward: bad sizes");
    CompositeArgs a{};
    a.sigma = sigma; a.rgb = rgb; a.z = z; a.rays_o = rays_o; a.rays_d = rays_d; a.rays_d_ndc = rays_d_ndc;
    a.g_rgb_map = d_rgb_map; a.g_acc = d_acc; a.g_depth = d_depth; a.g_depth_var = d_depth_var;
    a.g_depth_ndc = d_depth_ndc; a.g_depth_var_ndc = d_depth_var_ndc; a.g_alpha = d_alpha; a.g_vis = d_visibility;
    a.g_weights = d_weights; a.d_sigma = d_sigma; a.d_rgb = d_rgb;
    a.n_rays = n_rays; a.s = n_samples; a.ndc = ndc; a.white = (flags & SNERF_FLAG_WHITE_BKGD) != 0;
    return dispatch_composite(a, true, (cudaStream_t)stream);
}
