// Alpha compositing, forward and backward: one warp per ray, fp32, HBM-bound.
//
// Reference behaviour (src/models/SimpleNeRF01.py): volume_rendering :430-483 and
// convert_depth_from_ndc :485-502.
//
// Layout of the work: lane l of the warp owns C = S/32 CONSECUTIVE samples (l*C .. l*C+C-1), so sigma, z and rgb are
// read with 8/16-byte vector loads that are contiguous across the warp (fully coalesced 128-byte lines, no strided
// rgb access), the transmittance needs one lane-local product plus ONE 5-step shuffle scan per ray, and every per-sample
// output is a vector store.  Rays whose sample count is not a multiple of 32 take the strided fallback (lane l owns
// samples l, l+32, ...).
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

// sum over the G lanes (32 or 16) that share a ray
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- vector access of C consecutive floats (C even) -------------------------------------------------------
template <int C>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[C]) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (int i = 0; i < C / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < C / 2; ++i) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(p) + i);
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
        }
    }
}
template <int C>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[C]) {
    if constexpr (C % 4 == 0) {
#pragma unroll
        for (int i = 0; i < C / 4; ++i)
            reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < C / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
    }
}

// Per-ray quantities shared by forward and backward.  CONTIG: lane l of the ray's G-lane group owns samples l*C + i
// (G = 16 puts two rays in a warp: 64-sample rays are issue-bound, and the per-ray setup, scans and reductions are then paid
// once per two rays); else (G = 32) lane owns samples lane + 32*i.  `lane` below is the lane WITHIN the group.
template <int C, bool CONTIG, int G = kWarp>
struct RayState {
    float sig[C], zz[C], zm[C];      // sigma, z (ndc or metric), metric z
    float delta[C], alpha[C], trans[C], w[C];
    float acc, dsum, dsum_ndc;       // sum w, sum w*zm, sum w*z_ndc

    __device__ __forceinline__ int sample(int lane, int i) const { return CONTIG ? lane * C + i : i * kWarp + lane; }

    __device__ __forceinline__ void load_and_scan(const CompositeArgs& a, int ray, int lane) {
        const int s = a.s;
        const float* sg = a.sigma + (size_t)ray * s;
        const float* z = a.z + (size_t)ray * s;
        const float* dvec = (a.ndc ? a.rays_d_ndc : a.rays_d) + (size_t)ray * 3;
        const float dn = sqrtf(dvec[0] * dvec[0] + dvec[1] * dvec[1] + dvec[2] * dvec[2]);          // :436 / :441
        const float tail = a.ndc ? 1.f : 1e10f;                                                     // :433 / :438
        float k0 = 0.f, tn = 0.f;
        if (a.ndc) {
            const float oz = a.rays_o[(size_t)ray * 3 + 2], dz = a.rays_d[(size_t)ray * 3 + 2];
            tn = -(1.f + oz) / dz;                                                                  // :498
            k0 = (oz + tn * dz) / dz;                                                               // per-ray factor of :501
        }
        float znext[C];
        if constexpr (CONTIG) {
            load_vec<C>(sg + lane * C, sig);
            load_vec<C>(z + lane * C, zz);
            const float up = __shfl_down_sync(kFull, zz[0], 1);      // first depth of the next lane
#pragma unroll
            for (int i = 0; i < C - 1; ++i) znext[i] = zz[i + 1];
            znext[C - 1] = lane == G - 1 ? tail : up;
        } else {
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int k = i * kWarp + lane;
                sig[i] = k < s ? sg[k] : 0.f;
                zz[i] = k < s ? z[k] : 0.f;
                znext[i] = (k + 1 < s) ? z[k + 1] : tail;
            }
        }
        float fprod = 1.f;   // product of this lane's (1 - alpha + 1e-10)   (contiguous ownership)
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const bool live = CONTIG || (i * kWarp + lane < s);
            delta[i] = (znext[i] - zz[i]) * dn;                                                     // :435-436
            alpha[i] = live ? 1.f - __expf(-sig[i] * delta[i]) : 0.f;                               // :446
            if (a.ndc) {
                const float guard = (zz[i] == 1.f) ? 1e-3f : 0.f;                                   // :499
                zm[i] = k0 * (__frcp_rn(1.f - zz[i] + guard) - 1.f) + tn;                           // :501
            } else {
                zm[i] = zz[i];
            }
            fprod *= (1.f - alpha[i]) + 1e-10f;
        }
        acc = dsum = dsum_ndc = 0.f;
        if constexpr (CONTIG) {
            // exclusive product scan over lanes, then walk the lane's own samples                  // :447
            float incl = fprod;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const float v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl *= v;
            }
            float t = __shfl_up_sync(kFull, incl, 1);
            if (lane == 0) t = 1.f;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                trans[i] = t;
                w[i] = alpha[i] * t;                                                                // :448
                t *= (1.f - alpha[i]) + 1e-10f;
                acc += w[i];
                dsum += w[i] * zm[i];
                dsum_ndc += w[i] * zz[i];
            }
        } else {
            float carry = 1.f;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const int k = i * kWarp + lane;
                const float f = (k < s) ? (1.f - alpha[i]) + 1e-10f : 1.f;
                float incl = f;
#pragma unroll
                for (int o = 1; o < kWarp; o <<= 1) {
                    const float v = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl *= v;
                }
                float excl = __shfl_up_sync(kFull, incl, 1);
                if (lane == 0) excl = 1.f;
                trans[i] = carry * excl;
                carry *= __shfl_sync(kFull, incl, kWarp - 1);
                w[i] = alpha[i] * trans[i];
                acc += w[i];
                dsum += w[i] * zm[i];
                dsum_ndc += w[i] * zz[i];
            }
        }
        acc = group_sum<G>(acc);                                                                    // :451
        dsum = group_sum<G>(dsum);
        dsum_ndc = group_sum<G>(dsum_ndc);
    }
};

template <int C, bool CONTIG, int G>
__global__ void __launch_bounds__(kCompWarps* kWarp) composite_fwd_kernel(const CompositeArgs a) {
    static_assert(G == kWarp || (CONTIG && G == 16), "two rays per warp only with lane-contiguous ownership");
    const int warp = threadIdx.x / kWarp, lane = (threadIdx.x % kWarp) % G;
    const int ray_raw = (blockIdx.x * kCompWarps + warp) * (kWarp / G) + (threadIdx.x % kWarp) / G;
    if (G == kWarp && ray_raw >= a.n_rays) return;
    const bool valid = ray_raw < a.n_rays;               // G = 16: the odd last ray's partner half works on a copy and stores nothing
    const int ray = valid ? ray_raw : a.n_rays - 1;
    const int s = a.s;
    // the colour row is requested before the depth / density rows are consumed: one memory round trip per ray, not two
    const float* rgb = a.rgb + (size_t)ray * s * 3;
    float c[CONTIG ? 3 * C : 1];
    if constexpr (CONTIG) load_vec<3 * C>(rgb + lane * 3 * C, c);
    RayState<C, CONTIG, G> r;
    r.load_and_scan(a, ray, lane);

    float cr = 0.f, cg = 0.f, cb = 0.f;
    if constexpr (CONTIG) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            cr += r.w[i] * c[3 * i];                                                                // :449
            cg += r.w[i] * c[3 * i + 1];
            cb += r.w[i] * c[3 * i + 2];
        }
        const size_t o = (size_t)ray * s + lane * C;
        if (a.alpha && valid) store_vec<C>(a.alpha + o, r.alpha);
        if (a.vis && valid) store_vec<C>(a.vis + o, r.trans);
        if (a.weights && valid) store_vec<C>(a.weights + o, r.w);
    } else {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const int k = i * kWarp + lane;
            if (k < s) {
                cr += r.w[i] * rgb[k * 3 + 0];
                cg += r.w[i] * rgb[k * 3 + 1];
                cb += r.w[i] * rgb[k * 3 + 2];
                const size_t o = (size_t)ray * s + k;
                if (a.alpha) a.alpha[o] = r.alpha[i];
                if (a.vis) a.vis[o] = r.trans[i];
                if (a.weights) a.weights[o] = r.w[i];
            }
        }
    }
    cr = group_sum<G>(cr);
    cg = group_sum<G>(cg);
    cb = group_sum<G>(cb);
    const float inv = 1.f / (r.acc + 1e-6f);
    const float depth = r.dsum * inv;                                                               // :453 / :459
    const float depth_ndc = r.dsum_ndc * inv;                                                       // :456
    float var = 0.f, var_ndc = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const float e = r.zm[i] - depth, en = r.zz[i] - depth_ndc;
        var += r.w[i] * e * e;                                                                      // :454 / :460
        var_ndc += r.w[i] * en * en;                                                                // :457
    }
    var = group_sum<G>(var);
    var_ndc = group_sum<G>(var_ndc);
    if (lane == 0 && valid) {
        const float bg = a.white ? 1.f - r.acc : 0.f;                                               // :463
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
template <int C, bool CONTIG, int G>
__global__ void __launch_bounds__(kCompWarps* kWarp, (C >= 8 ? 3 : 0)) composite_bwd_kernel(const CompositeArgs a) {
    static_assert(G == kWarp || (CONTIG && G == 16), "two rays per warp only with lane-contiguous ownership");
    const int warp = threadIdx.x / kWarp, lane = (threadIdx.x % kWarp) % G;
    const int ray_raw = (blockIdx.x * kCompWarps + warp) * (kWarp / G) + (threadIdx.x % kWarp) / G;
    if (G == kWarp && ray_raw >= a.n_rays) return;
    const bool valid = ray_raw < a.n_rays;
    const int ray = valid ? ray_raw : a.n_rays - 1;
    const int s = a.s;
    const float* rgb = a.rgb + (size_t)ray * s * 3;
    float c[CONTIG ? 3 * C : 1];
    if constexpr (CONTIG) load_vec<3 * C>(rgb + lane * 3 * C, c);     // in flight during the scan
    RayState<C, CONTIG, G> r;
    r.load_and_scan(a, ray, lane);

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
    const bool per_sample_grads = a.g_weights || a.g_vis || a.g_alpha;

    float g[C], gt[C];
    float lane_gt = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const int k = r.sample(lane, i);
        g[i] = 0.f;
        gt[i] = 0.f;
        if (CONTIG || k < s) {
            const size_t o = (size_t)ray * s + k;
            float c0, c1, c2;
            if constexpr (CONTIG) { c0 = c[3 * i]; c1 = c[3 * i + 1]; c2 = c[3 * i + 2]; }
            else { c0 = rgb[k * 3 + 0]; c1 = rgb[k * 3 + 1]; c2 = rgb[k * 3 + 2]; }
            const float e = r.zm[i] - depth, en = r.zz[i] - depth_ndc;
            float gi = gr * c0 + gg * c1 + gb * c2 + g_const;
            gi += gd * e * inv + gdn * en * inv;
            gi += gv * (e * e - 2.f * e * inv * resid) + gvn * (en * en - 2.f * en * inv * resid_ndc);
            float big_g;
            if (per_sample_grads) {
                if (a.g_weights) gi += a.g_weights[o];
                big_g = gi * r.alpha[i];
                if (a.g_vis) big_g += a.g_vis[o];
            } else {
                big_g = gi * r.alpha[i];
            }
            g[i] = gi;
            gt[i] = big_g * r.trans[i];
            lane_gt += gt[i];
        }
    }
    // d rgb = w * d rgb_map
    if constexpr (CONTIG && (C > 4)) {
        // 3C floats per lane are 72 / 96 bytes apart: a direct vector store touches 32 sectors per request and fills a
        // quarter / half of each (ncu, S=192: 30 sectors per store request, 4x the useful L2 write traffic).  The row goes
        // through shared memory and leaves as whole 512-byte warp stores.
        __shared__ __align__(16) float s_rows[kCompWarps][3 * C * kWarp];
        float* row = s_rows[warp];
        float dr[3 * C];
#pragma unroll
        for (int i = 0; i < C; ++i) { dr[3 * i] = r.w[i] * gr; dr[3 * i + 1] = r.w[i] * gg; dr[3 * i + 2] = r.w[i] * gb; }
        store_vec<3 * C>(row + lane * 3 * C, dr);     // 8/16-byte shared stores: at most two lanes per bank
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(a.d_rgb + (size_t)ray * s * 3);
#pragma unroll
        for (int i = 0; i < (3 * C * kWarp / 4 + kWarp - 1) / kWarp; ++i) {
            const int q = i * kWarp + lane;
            if (q < 3 * C * kWarp / 4) dst[q] = reinterpret_cast<const float4*>(row)[q];
        }
        __syncwarp();
    } else if constexpr (CONTIG) {
        float dr[3 * C];
#pragma unroll
        for (int i = 0; i < C; ++i) { dr[3 * i] = r.w[i] * gr; dr[3 * i + 1] = r.w[i] * gg; dr[3 * i + 2] = r.w[i] * gb; }
        if (valid) store_vec<3 * C>(a.d_rgb + ((size_t)ray * s + lane * C) * 3, dr);
    } else {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const int k = i * kWarp + lane;
            if (k < s) {
                const size_t o = (size_t)ray * s + k;
                a.d_rgb[o * 3 + 0] = r.w[i] * gr;
                a.d_rgb[o * 3 + 1] = r.w[i] * gg;
                a.d_rgb[o * 3 + 2] = r.w[i] * gb;
            }
        }
    }
    // suffix sums  sum_{j>k} G_j T_j
    float ds[C];
    if constexpr (CONTIG) {
        float incl = lane_gt;    // inclusive suffix over the lanes of the ray's group
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
            const float v = __shfl_down_sync(kFull, incl, off);
            if (lane + off < G) incl += v;
        }
        float suffix = incl - lane_gt;           // everything owned by later lanes
#pragma unroll
        for (int i = C - 1; i >= 0; --i) {
            const float f = (1.f - r.alpha[i]) + 1e-10f;
            float d_alpha = g[i] * r.trans[i] - suffix * __frcp_rn(f);
            if (a.g_alpha) d_alpha += a.g_alpha[(size_t)ray * s + lane * C + i];
            ds[i] = d_alpha * r.delta[i] * (1.f - r.alpha[i]);
            suffix += gt[i];
        }
        if constexpr (C > 4) {       // as above: C floats per lane -> coalesced float2 stores through the (now free) row
            __shared__ __align__(16) float s_sig[kCompWarps][C * kWarp];
            float* row = s_sig[warp];
            store_vec<C>(row + lane * C, ds);
            __syncwarp();
            float2* dst = reinterpret_cast<float2*>(a.d_sigma + (size_t)ray * s);
#pragma unroll
            for (int i = 0; i < C / 2; ++i) dst[i * kWarp + lane] = reinterpret_cast<const float2*>(row)[i * kWarp + lane];
        } else {
            if (valid) store_vec<C>(a.d_sigma + (size_t)ray * s + lane * C, ds);
        }
    } else {
        float carry = 0.f;
#pragma unroll
        for (int i = C - 1; i >= 0; --i) {
            const int k = i * kWarp + lane;
            float incl = gt[i];
#pragma unroll
            for (int off = 1; off < kWarp; off <<= 1) {
                const float v = __shfl_down_sync(kFull, incl, off);
                if (lane + off < kWarp) incl += v;
            }
            const float suffix = carry + (incl - gt[i]);
            carry += __shfl_sync(kFull, incl, 0);
            if (k < s) {
                const size_t o = (size_t)ray * s + k;
                const float f = (1.f - r.alpha[i]) + 1e-10f;
                float d_alpha = g[i] * r.trans[i] - suffix / f;
                if (a.g_alpha) d_alpha += a.g_alpha[o];
                a.d_sigma[o] = d_alpha * r.delta[i] * (1.f - r.alpha[i]);
            }
        }
    }
}

// Backward for 256 samples per ray: TWO warps per ray, four consecutive samples per lane.  With one warp per ray a lane owns
// eight samples and the kernel needs 108 registers (capped at 80: 0.62 of the HBM bandwidth); four per lane is the shape of the
// 128-sample kernel (0.72).  The two warps meet three times through shared memory: the transmittance that reaches the second
// half, the per-ray sums, and the suffix sum of the second half.  Arithmetic and order of operations inside a warp are those
// of RayState / composite_bwd_kernel above.
__global__ void __launch_bounds__(kCompWarps* kWarp) composite_bwd_pair_kernel(const CompositeArgs a) {
    constexpr int C = 4;
    __shared__ float s_x[kCompWarps / 2][2][8];
    const int warp = threadIdx.x / kWarp, l32 = threadIdx.x % kWarp, half = warp & 1, pair = warp >> 1;
    const int ray_raw = blockIdx.x * (kCompWarps / 2) + pair;
    const bool valid = ray_raw < a.n_rays;
    const int ray = valid ? ray_raw : a.n_rays - 1;            // both warps of a pair stay in step through the barriers
    const int s = a.s, lane = half * kWarp + l32;              // lane within the ray: samples lane * 4 .. + 3
    auto meet = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); };
    float* mine = s_x[pair][half];
    const float* other = s_x[pair][half ^ 1];

    const float* rgb = a.rgb + (size_t)ray * s * 3;
    float c[3 * C];
    load_vec<3 * C>(rgb + lane * 3 * C, c);
    const float* sg = a.sigma + (size_t)ray * s;
    const float* z = a.z + (size_t)ray * s;
    const float* dvec = (a.ndc ? a.rays_d_ndc : a.rays_d) + (size_t)ray * 3;
    const float dn = sqrtf(dvec[0] * dvec[0] + dvec[1] * dvec[1] + dvec[2] * dvec[2]);
    const float tail = a.ndc ? 1.f : 1e10f;
    float k0 = 0.f, tn = 0.f;
    if (a.ndc) {
        const float oz = a.rays_o[(size_t)ray * 3 + 2], dz = a.rays_d[(size_t)ray * 3 + 2];
        tn = -(1.f + oz) / dz;
        k0 = (oz + tn * dz) / dz;
    }
    float sig[C], zz[C], zm[C], delta[C], alpha[C], trans[C], w[C];
    load_vec<C>(sg + lane * C, sig);
    load_vec<C>(z + lane * C, zz);
    const float up = lane == 63 ? tail : __ldg(z + lane * C + C);       // first depth of the next lane (the other warp's for lane 31)
    float fprod = 1.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const float zn = i == C - 1 ? up : zz[i + 1];
        delta[i] = (zn - zz[i]) * dn;
        alpha[i] = 1.f - __expf(-sig[i] * delta[i]);
        if (a.ndc) {
            const float guard = (zz[i] == 1.f) ? 1e-3f : 0.f;
            zm[i] = k0 * (__frcp_rn(1.f - zz[i] + guard) - 1.f) + tn;
        } else {
            zm[i] = zz[i];
        }
        fprod *= (1.f - alpha[i]) + 1e-10f;
    }
    float incl = fprod;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
        const float v = __shfl_up_sync(kFull, incl, o);
        if (l32 >= o) incl *= v;
    }
    float t = __shfl_up_sync(kFull, incl, 1);
    if (l32 == 0) t = 1.f;
    if (l32 == kWarp - 1) mine[0] = incl;                      // product of this half
    meet();
    if (half == 1) t *= other[0];                              // transmittance that reaches the second half
    float acc = 0.f, dsum = 0.f, dsum_ndc = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        trans[i] = t;
        w[i] = alpha[i] * t;
        t *= (1.f - alpha[i]) + 1e-10f;
        acc += w[i];
        dsum += w[i] * zm[i];
        dsum_ndc += w[i] * zz[i];
    }
    acc = group_sum<kWarp>(acc);
    dsum = group_sum<kWarp>(dsum);
    dsum_ndc = group_sum<kWarp>(dsum_ndc);
    if (l32 == 0) { mine[1] = acc; mine[2] = dsum; mine[3] = dsum_ndc; }
    meet();
    // first half + second half, in that order in both warps: the two agree bit for bit
    acc = half == 0 ? acc + other[1] : other[1] + acc;
    dsum = half == 0 ? dsum + other[2] : other[2] + dsum;
    dsum_ndc = half == 0 ? dsum_ndc + other[3] : other[3] + dsum_ndc;

    const float inv = 1.f / (acc + 1e-6f);
    const float depth = dsum * inv, depth_ndc = dsum_ndc * inv;
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
    const float resid = dsum - depth * acc, resid_ndc = dsum_ndc - depth_ndc * acc;
    float g[C], gt[C], lane_gt = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const size_t o = (size_t)ray * s + lane * C + i;
        const float e = zm[i] - depth, en = zz[i] - depth_ndc;
        float gi = gr * c[3 * i] + gg * c[3 * i + 1] + gb * c[3 * i + 2] + g_const;
        gi += gd * e * inv + gdn * en * inv;
        gi += gv * (e * e - 2.f * e * inv * resid) + gvn * (en * en - 2.f * en * inv * resid_ndc);
        if (a.g_weights) gi += a.g_weights[o];
        float big_g = gi * alpha[i];
        if (a.g_vis) big_g += a.g_vis[o];
        g[i] = gi;
        gt[i] = big_g * trans[i];
        lane_gt += gt[i];
    }
    float dr[3 * C];
#pragma unroll
    for (int i = 0; i < C; ++i) { dr[3 * i] = w[i] * gr; dr[3 * i + 1] = w[i] * gg; dr[3 * i + 2] = w[i] * gb; }
    if (valid) store_vec<3 * C>(a.d_rgb + ((size_t)ray * s + lane * C) * 3, dr);
    float sincl = lane_gt;                                     // inclusive suffix over the lanes of this half
#pragma unroll
    for (int off = 1; off < kWarp; off <<= 1) {
        const float v = __shfl_down_sync(kFull, sincl, off);
        if (l32 + off < kWarp) sincl += v;
    }
    if (l32 == 0) mine[4] = sincl;                             // everything owned by this half
    meet();
    float suffix = sincl - lane_gt;
    if (half == 0) suffix += other[4];
    float ds[C];
#pragma unroll
    for (int i = C - 1; i >= 0; --i) {
        const float f = (1.f - alpha[i]) + 1e-10f;
        float d_alpha = g[i] * trans[i] - suffix * __frcp_rn(f);
        if (a.g_alpha) d_alpha += a.g_alpha[(size_t)ray * s + lane * C + i];
        ds[i] = d_alpha * delta[i] * (1.f - alpha[i]);
        suffix += gt[i];
    }
    if (valid) store_vec<C>(a.d_sigma + (size_t)ray * s + lane * C, ds);
}

template <int C, bool CONTIG, int G = kWarp>
static int launch_composite(const CompositeArgs& a, bool backward, cudaStream_t st) {
    const int blocks = ceil_div(a.n_rays, kCompWarps * (kWarp / G));
    if (backward)
        composite_bwd_kernel<C, CONTIG, G><<<blocks, kCompWarps * kWarp, 0, st>>>(a);
    else
        composite_fwd_kernel<C, CONTIG, G><<<blocks, kCompWarps * kWarp, 0, st>>>(a);
    SNERF_LAUNCH_OK(backward ? "composite_bwd_kernel" : "composite_fwd_kernel");
    return SNERF_OK;
}

static int dispatch_composite(const CompositeArgs& a, bool backward, cudaStream_t st) {
    if (a.n_rays == 0) return SNERF_OK;
    switch (a.s) {   // vectorised, lane-contiguous kernels for the sample counts the model uses
        case 64: return launch_composite<4, true, 16>(a, backward, st);    // two rays per warp, float4 rows
        case 128: return launch_composite<4, true>(a, backward, st);
        case 192: return launch_composite<6, true>(a, backward, st);
        case 256:
            if (backward) {
                composite_bwd_pair_kernel<<<ceil_div(a.n_rays, kCompWarps / 2), kCompWarps * kWarp, 0, st>>>(a);
                SNERF_LAUNCH_OK("composite_bwd_pair_kernel");
                return SNERF_OK;
            }
            return launch_composite<8, true>(a, backward, st);
        default: break;
    }
    const int items = ceil_div(a.s, kWarp);
    if (items <= 2) return launch_composite<2, false>(a, backward, st);
    if (items <= 4) return launch_composite<4, false>(a, backward, st);
    if (items <= 8) return launch_composite<8, false>(a, backward, st);
    if (items <= 16) return launch_composite<16, false>(a, backward, st);
    return fail(SNERF_ERR_UNSUPPORTED, "composite: %d samples per ray > 512", a.s);
}

// ------------------------------------------------------------------------------------------------
// Row a14 / N4: visibility2 = sum_s w_s vis2[s, v] / (acc + 1e-6) per ray and other view (volume_rendering :479-482).
// One warp per ray.  Backward: d vis2 = G w / (acc + 1e-6); d w = sum_v G_v vis2_v / (acc + 1e-6);
// d acc = - sum_v G_v visibility2_v / (acc + 1e-6)  (acc = sum w: both are handed to the compositing backward).
// ------------------------------------------------------------------------------------------------
constexpr int kVis2MaxViews = 8;

__global__ void __launch_bounds__(kCompWarps* kWarp) vis2_composite_fwd_kernel(const float* __restrict__ weights, const float* __restrict__ acc,
                                                                              const float* __restrict__ vis2, float* __restrict__ out,
                                                                              int n_rays, int s, int nv) {
    const int ray = blockIdx.x * kCompWarps + threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    if (ray >= n_rays) return;
    float sum[kVis2MaxViews] = {};
    for (int k = lane; k < s; k += kWarp) {
        const float w = weights[(size_t)ray * s + k];
        for (int v = 0; v < nv; ++v) sum[v] += w * vis2[((size_t)ray * s + k) * nv + v];
    }
    const float inv = 1.f / (acc[ray] + 1e-6f);
    for (int v = 0; v < nv; ++v) {
        const float t = group_sum<kWarp>(sum[v]);
        if (lane == 0) out[(size_t)ray * nv + v] = t * inv;
    }
}

__global__ void __launch_bounds__(kCompWarps* kWarp) vis2_composite_bwd_kernel(const float* __restrict__ weights, const float* __restrict__ acc,
                                                                              const float* __restrict__ vis2, const float* __restrict__ vis2_map,
                                                                              const float* __restrict__ g_map, float* __restrict__ d_vis2,
                                                                              float* __restrict__ d_weights, float* __restrict__ d_acc,
                                                                              int n_rays, int s, int nv) {
    const int ray = blockIdx.x * kCompWarps + threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    if (ray >= n_rays) return;
    const float inv = 1.f / (acc[ray] + 1e-6f);
    float g[kVis2MaxViews], dacc = 0.f;
    for (int v = 0; v < nv; ++v) {
        g[v] = g_map[(size_t)ray * nv + v] * inv;
        dacc -= g[v] * vis2_map[(size_t)ray * nv + v];
    }
    if (lane == 0) d_acc[ray] = dacc;
    for (int k = lane; k < s; k += kWarp) {
        const size_t o = (size_t)ray * s + k;
        const float w = weights[o];
        float dw = 0.f;
        for (int v = 0; v < nv; ++v) {
            dw += g[v] * vis2[o * nv + v];
            d_vis2[o * nv + v] = g[v] * w;
        }
        d_weights[o] = dw;
    }
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
    CompositeArgs a{};
    a.sigma = sigma; a.rgb = rgb; a.z = z; a.rays_o = rays_o; a.rays_d = rays_d; a.rays_d_ndc = rays_d_ndc;
    a.g_rgb_map = d_rgb_map; a.g_acc = d_acc; a.g_depth = d_depth; a.g_depth_var = d_depth_var;
    a.g_depth_ndc = d_depth_ndc; a.g_depth_var_ndc = d_depth_var_ndc; a.g_alpha = d_alpha; a.g_vis = d_visibility;
    a.g_weights = d_weights; a.d_sigma = d_sigma; a.d_rgb = d_rgb;
    a.n_rays = n_rays; a.s = n_samples; a.ndc = ndc; a.white = (flags & SNERF_FLAG_WHITE_BKGD) != 0;
    return dispatch_composite(a, true, (cudaStream_t)stream);
}

extern "C" int snerf_visibility2_composite_forward(const float* weights, const float* acc, const float* visibility2, float* visibility2_map,
                                                   int n_rays, int n_samples, int n_other, void* stream) {
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_other >= 1 && n_other <= kVis2MaxViews,
                  "snerf_visibility2_composite_forward: bad sizes (%d rays, %d samples, %d other views, max %d)", n_rays, n_samples, n_other,
                  kVis2MaxViews);
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(weights && acc && visibility2 && visibility2_map, "snerf_visibility2_composite_forward: null pointer");
    vis2_composite_fwd_kernel<<<ceil_div(n_rays, kCompWarps), kCompWarps * kWarp, 0, (cudaStream_t)stream>>>(weights, acc, visibility2, visibility2_map,
                                                                                                         n_rays, n_samples, n_other);
    SNERF_LAUNCH_OK("vis2_composite_fwd_kernel");
    return SNERF_OK;
}

extern "C" int snerf_visibility2_composite_backward(const float* weights, const float* acc, const float* visibility2, const float* visibility2_map,
                                                    const float* d_visibility2_map, float* d_visibility2, float* d_weights, float* d_acc,
                                                    int n_rays, int n_samples, int n_other, void* stream) {
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1 && n_other >= 1 && n_other <= kVis2MaxViews, "snerf_visibility2_composite_backward: bad sizes");
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(weights && acc && visibility2 && visibility2_map && d_visibility2_map && d_visibility2 && d_weights && d_acc,
                  "snerf_visibility2_composite_backward: null pointer");
    vis2_composite_bwd_kernel<<<ceil_div(n_rays, kCompWarps), kCompWarps * kWarp, 0, (cudaStream_t)stream>>>(
        weights, acc, visibility2, visibility2_map, d_visibility2_map, d_visibility2, d_weights, d_acc, n_rays, n_samples, n_other);
    SNERF_LAUNCH_OK("vis2_composite_bwd_kernel");
    return SNERF_OK;
}
