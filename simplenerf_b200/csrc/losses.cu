// Masked per-ray losses of the training step in two launches (next-row N3).
//
// Reference behaviour (src/loss_functions):
//   MSE01/02/03.compute_mse            MSE01.py:53-67   pred[mask], target[mask]; mean over channels, mean over rays
//   SparseDepthMSE01/02/03.compute_depth_loss  SparseDepthMSE01.py:58-71   the same on one channel
//   LossComputer.compute_losses        LossComputer01.py:33-52   total = sum_k weight_k * loss_k
// Each of these is ~10 eager kernels plus a boolean-mask gather (a device synchronisation) per stream and again that
// many in autograd; the shipped configuration has eight streams (rgb and depth of the coarse, fine, points-augmented
// and views-augmented renders).  Here a "stream" is (prediction [N,C], target [N,C], mask [N], weight):
//   forward : loss_s = sum_{masked rays, channels} (pred - target)^2 / (count_s * C)   (0 when nothing is masked in),
//             total = sum_s weight_s * loss_s; fixed summation order (per-block partials, reduced by the last block
//             to finish), so the values are reproducible run to run;
//   backward: grad_s[i, c] = coeff_s * 2 (pred - target) / (count_s * C) on masked rays, 0 elsewhere, with
//             coeff_s = d loss_s + d total * weight_s read from the incoming gradient vector on the device.
// HBM-bound and tiny (36 bytes per ray and rgb stream): one grid-stride pass each.
#include "common.cuh"

namespace snerf {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 148;

struct LossTable {
    const float* pred[SNERF_LOSS_MAX_STREAMS];
    const float* target[SNERF_LOSS_MAX_STREAMS];
    const uint8_t* mask[SNERF_LOSS_MAX_STREAMS];
    float* grad[SNERF_LOSS_MAX_STREAMS];
    float weight[SNERF_LOSS_MAX_STREAMS];
    int channels[SNERF_LOSS_MAX_STREAMS];
    int kind[SNERF_LOSS_MAX_STREAMS];
    int n_streams;
};

// workspace: [blocks][streams] partial sums, [blocks][streams] partial counts, one ticket counter
struct LossWorkspace {
    float sums[kLossMaxBlocks][SNERF_LOSS_MAX_STREAMS];
    int counts[kLossMaxBlocks][SNERF_LOSS_MAX_STREAMS];
    unsigned int ticket;
};

// per-ray contribution of one stream: squared error (MSE01.py:56-57), absolute error (VisibilityLoss01.py:70-72) or the
// prior-weighted shortfall sum_c target_c (1 - pred_c) (VisibilityPriorLoss01.py:76-78)
__device__ __forceinline__ float ray_term(const float* __restrict__ p, const float* __restrict__ t, int i, int c, int kind) {
    float e = 0.f;
    for (int k = 0; k < c; ++k) {
        const float pv = p[(size_t)i * c + k], tv = t[(size_t)i * c + k];
        const float d = pv - tv;
        e += kind == SNERF_LOSS_SQUARED ? d * d : kind == SNERF_LOSS_ABSOLUTE ? fabsf(d) : tv * (1.f - pv);
    }
    return e;
}

__global__ void __launch_bounds__(kLossThreads) ray_losses_fwd_kernel(const __grid_constant__ LossTable t, int n_rays,
                                                                     float* __restrict__ values, int* __restrict__ counts,
                                                                     LossWorkspace* __restrict__ ws) {
    __shared__ float s_sum[kLossThreads / kWarp][SNERF_LOSS_MAX_STREAMS];
    __shared__ int s_cnt[kLossThreads / kWarp][SNERF_LOSS_MAX_STREAMS];
    __shared__ bool s_last;
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    for (int s = 0; s < t.n_streams; ++s) {
        float sum = 0.f;
        int cnt = 0;
        const uint8_t* m = t.mask[s];
        for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n_rays; i += gridDim.x * kLossThreads) {
            if (m == nullptr || m[i]) {
                sum += ray_term(t.pred[s], t.target[s], i, t.channels[s], t.kind[s]);
                cnt += 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(kFull, sum, o);
            cnt += __shfl_xor_sync(kFull, cnt, o);
        }
        if (lane == 0) { s_sum[warp][s] = sum; s_cnt[warp][s] = cnt; }
    }
    __syncthreads();
    if (threadIdx.x < t.n_streams) {
        float sum = 0.f;
        int cnt = 0;
        for (int w = 0; w < kLossThreads / kWarp; ++w) { sum += s_sum[w][threadIdx.x]; cnt += s_cnt[w][threadIdx.x]; }
        ws->sums[blockIdx.x][threadIdx.x] = sum;
        ws->counts[blockIdx.x][threadIdx.x] = cnt;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < t.n_streams) {                 // the last block: partials in block order
        float sum = 0.f;
        int cnt = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) {
            sum += __ldcg(&ws->sums[b][threadIdx.x]);
            cnt += __ldcg(&ws->counts[b][threadIdx.x]);
        }
        // mean over channels and rays (MSE01.py:57-58); the prior-weighted sum is a mean over rays only (VisibilityPriorLoss01.py:78-79)
        const float per_ray = t.kind[threadIdx.x] == SNERF_LOSS_PRIOR_SHORTFALL ? 1.f : (float)t.channels[threadIdx.x];
        const float v = cnt > 0 ? sum / ((float)cnt * per_ray) : 0.f;
        values[threadIdx.x] = v;
        counts[threadIdx.x] = cnt;
        s_sum[0][threadIdx.x] = v * t.weight[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int s = 0; s < t.n_streams; ++s) total += s_sum[0][s];                             // LossComputer01.py:48
        values[t.n_streams] = total;
        ws->ticket = 0;                               // ready for the next call
    }
}

__global__ void __launch_bounds__(kLossThreads) ray_losses_bwd_kernel(const __grid_constant__ LossTable t, int n_rays,
                                                                     const int* __restrict__ counts,
                                                                     const float* __restrict__ g_values) {
    const float g_total = g_values[t.n_streams];
    for (int s = 0; s < t.n_streams; ++s) {
        const int c = t.channels[s], cnt = counts[s];
        const float coeff = g_values[s] + g_total * t.weight[s];
        const int kind = t.kind[s];
        const float per_ray = kind == SNERF_LOSS_PRIOR_SHORTFALL ? 1.f : (float)c;
        const float scale = cnt > 0 ? coeff / ((float)cnt * per_ray) : 0.f;
        const uint8_t* m = t.mask[s];
        const float* p = t.pred[s];
        const float* tg = t.target[s];
        float* g = t.grad[s];
        for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n_rays; i += gridDim.x * kLossThreads) {
            const bool on = m == nullptr || m[i];
            for (int k = 0; k < c; ++k) {
                const size_t o = (size_t)i * c + k;
                const float d = p[o] - tg[o];
                const float dv = kind == SNERF_LOSS_SQUARED ? 2.f * d
                               : kind == SNERF_LOSS_ABSOLUTE ? (d > 0.f ? 1.f : d < 0.f ? -1.f : 0.f)     // torch.abs: sign, 0 at 0
                                                            : -tg[o];
                g[o] = on ? scale * dv : 0.f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Patch-reprojection depth losses: PointsAugmentationDepthLoss02 / ViewsAugmentationDepthLoss02 /
// CoarseFineConsistencyLoss02 .compute_loss_nerf (src/loss_functions/PointsAugmentationDepthLoss02.py:98-176; the
// three modules share the function line for line).  For every ray of the NeRF mask the 3D points at the depths of two
// models are projected into the nearest other view (CommonUtils01.reproject :45-72), the 5x5 rgb patches around the
// two projections are compared with the patch around the ray's own pixel, and the depth of the model with the larger
// patch RMSE is pulled towards the other one.  All three losses compare the main coarse depth ("main") with another
// model's depth ("other" k), so one warp per ray gathers the source patch and the main model's patch once.
//
// Reference quirk reproduced on purpose: compute_depth_mse (:196-212) zeroes `pred_depth[~mask]` and `gt_depth[~mask]`
// IN PLACE, and gt_depth is `depth.detach()`, which shares storage with the other call's pred_depth.  The first call
// (depth1 against depth2 under mask2) therefore zeroes both depths everywhere outside mask2; mask1 and mask2 are
// mutually exclusive, so the second call sees 0 - 0 on every ray: its value and its gradient are identically zero.
// The loss the reference optimises is mean_{nerf rays}(mask2 (d_main - d_other)^2) with a gradient for the MAIN depth
// only (tests/golden/losses.npz, generated from the unmodified modules, pins this).  SNERF_REPROJ_SYMMETRIC adds the
// evidently intended second term.
// ------------------------------------------------------------------------------------------------
constexpr int kReprojWarps = 8;

struct ReprojTable {
    const float* depth_main;
    const float* depth_other[SNERF_REPROJ_MAX_OTHERS];
    float* grad_main;
    float* grad_other[SNERF_REPROJ_MAX_OTHERS];
    float weight[SNERF_REPROJ_MAX_OTHERS];
    int n_others;
    const float *rays_o, *rays_d;
    const int32_t* pixel_id;    // [n,3] = (view, x, y)
    const uint8_t* mask_nerf;   // nullable
    const float* images;        // [views, h, w, 3]
    const float* proj;          // [views, 9]  K[0] diag(1,-1,-1) R_v^T
    const float* origins;       // [views, 3]
    const int32_t* closest;     // [views]
    int n_views, h, w, hp;
    float threshold;
    bool symmetric;
};

// rounded pixel position of `point` in view b; false if it is not a finite position
__device__ __forceinline__ bool project(const ReprojTable& t, int b, float px, float py, float pz, int& x, int& y) {
    const float* m = t.proj + 9 * b;
    const float q0 = px - t.origins[3 * b], q1 = py - t.origins[3 * b + 1], q2 = pz - t.origins[3 * b + 2];
    const float p0 = fmaf(m[2], q2, fmaf(m[1], q1, m[0] * q0));
    const float p1 = fmaf(m[5], q2, fmaf(m[4], q1, m[3] * q0));
    const float p2 = fmaf(m[8], q2, fmaf(m[7], q1, m[6] * q0));
    const float u = rintf(__fdiv_rn(p0, p2)), v = rintf(__fdiv_rn(p1, p2));                        // :70-71, .round()
    if (!(fabsf(u) < 1e9f && fabsf(v) < 1e9f)) return false;
    x = (int)u;
    y = (int)v;
    return true;
}

__global__ void __launch_bounds__(kReprojWarps* kWarp) reproj_fwd_kernel(const __grid_constant__ ReprojTable t, int n_rays,
                                                                         uint8_t* __restrict__ codes, float* __restrict__ values,
                                                                         int* __restrict__ counts, LossWorkspace* __restrict__ ws) {
    __shared__ float s_sum[kReprojWarps][SNERF_REPROJ_MAX_OTHERS];
    __shared__ int s_cnt[kReprojWarps];
    __shared__ bool s_last;
    const int lane = threadIdx.x % kWarp, warp = threadIdx.x / kWarp;
    const int side = 2 * t.hp + 1, npix = side * side;
    float sums[SNERF_REPROJ_MAX_OTHERS] = {};
    int n_nerf = 0;
    for (int ray = blockIdx.x * kReprojWarps + warp; ray < n_rays; ray += gridDim.x * kReprojWarps) {
        if (t.mask_nerf != nullptr && !t.mask_nerf[ray]) {
            if (lane < t.n_others) codes[(size_t)lane * n_rays + ray] = 0;
            continue;
        }
        n_nerf += 1;
        const int va_id = t.pixel_id[3 * ray], xa = t.pixel_id[3 * ray + 1], ya = t.pixel_id[3 * ray + 2];
        const bool in_views = va_id >= 0 && va_id < t.n_views;
        const int b = in_views ? t.closest[va_id] : 0;
        const auto inside = [&](int x, int y) { return x >= t.hp && x < t.w - t.hp && y >= t.hp && y < t.h - t.hp; };   // :147-149
        const bool valid_a = in_views && inside(xa, ya);
        const float ox = t.rays_o[3 * ray], oy = t.rays_o[3 * ray + 1], oz = t.rays_o[3 * ray + 2];
        const float dx = t.rays_d[3 * ray], dy = t.rays_d[3 * ray + 1], dz = t.rays_d[3 * ray + 2];
        // patch pixel of this lane
        const int pdy = lane / side - t.hp, pdx = lane % side - t.hp;
        const bool has_pix = lane < npix;   // (patches up to 5x5 fit one warp; larger sides are refused by the host call)
        float pa[3] = {0.f, 0.f, 0.f};
        if (valid_a && has_pix) {
            const float* src = t.images + (((size_t)va_id * t.h + (ya + pdy)) * t.w + (xa + pdx)) * 3;
            pa[0] = src[0]; pa[1] = src[1]; pa[2] = src[2];
        }
        // patch RMSE of the point at `depth` seen from view b (:136-165, :180); valid = the projection lies inside the margin
        const auto patch_rmse = [&](float depth, bool& valid) {
            const float X = __fadd_rn(ox, __fmul_rn(dx, depth)), Y = __fadd_rn(oy, __fmul_rn(dy, depth)),
                        Z = __fadd_rn(oz, __fmul_rn(dz, depth));
            int x = 0, y = 0;
            valid = project(t, b, X, Y, Z, x, y) && inside(x, y);
            float e = 0.f;
            if (valid && has_pix) {
                const float* src = t.images + (((size_t)b * t.h + (y + pdy)) * t.w + (x + pdx)) * 3;
                const float e0 = pa[0] - src[0], e1 = pa[1] - src[1], e2 = pa[2] - src[2];
                e = e0 * e0 + e1 * e1 + e2 * e2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(kFull, e, o);
            return sqrtf(e / (float)(npix * 3));
        };
        const float d_main = t.depth_main[ray];
        bool v1 = false;
        float rmse1 = 0.f;
        if (valid_a) rmse1 = patch_rmse(d_main, v1);
        for (int k = 0; k < t.n_others; ++k) {
            const float d_other = t.depth_other[k][ray];
            bool v2 = false;
            float rmse2 = 0.f;
            if (valid_a) rmse2 = patch_rmse(d_other, v2);
            const bool m1 = valid_a && v1 && ((rmse1 < rmse2) || !v2) && (rmse1 < t.threshold);      // :167
            const bool m2 = valid_a && v2 && ((rmse2 < rmse1) || !v1) && (rmse2 < t.threshold);      // :169
            if (lane == 0) codes[(size_t)k * n_rays + ray] = (uint8_t)((m1 ? 1 : 0) | (m2 ? 2 : 0));
            const float diff = d_main - d_other;
            if (m2 || (t.symmetric && m1)) sums[k] += diff * diff;                                   // :172-175
        }
    }
    if (lane == 0) {
        for (int k = 0; k < SNERF_REPROJ_MAX_OTHERS; ++k) s_sum[warp][k] = sums[k];
        s_cnt[warp] = n_nerf;
    }
    __syncthreads();
    if (threadIdx.x <= t.n_others) {     // thread k < n_others: sum of stream k; thread n_others: the ray count
        float sum = 0.f;
        int cnt = 0;
        for (int w = 0; w < kReprojWarps; ++w) {
            if (threadIdx.x < t.n_others) sum += s_sum[w][threadIdx.x]; else cnt += s_cnt[w];
        }
        ws->sums[blockIdx.x][threadIdx.x] = sum;
        ws->counts[blockIdx.x][threadIdx.x] = cnt;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ float s_val[SNERF_REPROJ_MAX_OTHERS];
    __shared__ int s_n;
    if (threadIdx.x == t.n_others) {
        int cnt = 0;
        for (unsigned blk = 0; blk < gridDim.x; ++blk) cnt += __ldcg(&ws->counts[blk][t.n_others]);
        s_n = cnt;
        counts[0] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < t.n_others) {
        float sum = 0.f;
        for (unsigned blk = 0; blk < gridDim.x; ++blk) sum += __ldcg(&ws->sums[blk][threadIdx.x]);
        const float v = s_n > 0 ? sum / (float)s_n : 0.f;                                            // :209
        values[threadIdx.x] = v;
        s_val[threadIdx.x] = v * t.weight[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int k = 0; k < t.n_others; ++k) total += s_val[k];
        values[t.n_others] = total;
        ws->ticket = 0;
    }
}

__global__ void __launch_bounds__(kLossThreads) reproj_bwd_kernel(const __grid_constant__ ReprojTable t, int n_rays,
                                                                  const uint8_t* __restrict__ codes, const int* __restrict__ counts,
                                                                  const float* __restrict__ g_values) {
    const int n_nerf = counts[0];
    const float inv = n_nerf > 0 ? 2.f / (float)n_nerf : 0.f;
    const float g_total = g_values[t.n_others];
    for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n_rays; i += gridDim.x * kLossThreads) {
        const float d_main = t.depth_main[i];
        float g_main = 0.f;
        for (int k = 0; k < t.n_others; ++k) {
            const uint8_t code = codes[(size_t)k * n_rays + i];
            const float coeff = (g_values[k] + g_total * t.weight[k]) * inv;
            const float diff = d_main - t.depth_other[k][i];
            if (code & 2) g_main += coeff * diff;
            float g_other = 0.f;
            if (t.symmetric && (code & 1)) g_other = -coeff * diff;
            if (t.grad_other[k] != nullptr) t.grad_other[k][i] = g_other;
        }
        t.grad_main[i] = g_main;
    }
}

static int fill_table(LossTable& t, const snerf_loss_stream* streams, int n_streams, bool need_grad, const char* who) {
    SNERF_REQUIRE(n_streams >= 1 && n_streams <= SNERF_LOSS_MAX_STREAMS, "%s: %d streams (1..%d)", who, n_streams,
                  SNERF_LOSS_MAX_STREAMS);
    SNERF_REQUIRE(streams != nullptr, "%s: null stream table", who);
    t.n_streams = n_streams;
    for (int s = 0; s < n_streams; ++s) {
        SNERF_REQUIRE(streams[s].pred && streams[s].target, "%s: stream %d has a null pointer", who, s);
        SNERF_REQUIRE(!need_grad || streams[s].grad, "%s: stream %d has no gradient buffer", who, s);
        SNERF_REQUIRE(streams[s].channels >= 1 && streams[s].channels <= 1024, "%s: stream %d has %d channels (1..1024)", who, s,
                      streams[s].channels);
        SNERF_REQUIRE(streams[s].kind >= SNERF_LOSS_SQUARED && streams[s].kind <= SNERF_LOSS_PRIOR_SHORTFALL, "%s: stream %d has kind %d",
                      who, s, streams[s].kind);
        t.pred[s] = streams[s].pred; t.target[s] = streams[s].target; t.mask[s] = streams[s].mask;
        t.grad[s] = streams[s].grad; t.weight[s] = streams[s].weight; t.channels[s] = streams[s].channels;
        t.kind[s] = streams[s].kind;
    }
    return SNERF_OK;
}

}  // namespace snerf

using namespace snerf;

extern "C" size_t snerf_ray_losses_workspace_bytes(void) { return sizeof(LossWorkspace); }

extern "C" int snerf_ray_losses_forward(const snerf_loss_stream* streams, int n_streams, int n_rays, float* values,
                                        int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    SNERF_REQUIRE(n_rays >= 0, "snerf_ray_losses_forward: bad ray count %d", n_rays);
    LossTable t{};
    if (int rc = fill_table(t, streams, n_streams, false, "snerf_ray_losses_forward")) return rc;
    SNERF_REQUIRE(values && counts && workspace, "snerf_ray_losses_forward: null output / workspace");
    SNERF_REQUIRE(workspace_bytes >= sizeof(LossWorkspace), "snerf_ray_losses_forward: workspace of %zu bytes < %zu",
                  workspace_bytes, sizeof(LossWorkspace));
    const int blocks = max(1, min(kLossMaxBlocks, ceil_div(n_rays, kLossThreads)));
    ray_losses_fwd_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(t, n_rays, values, counts,
                                                                             static_cast<LossWorkspace*>(workspace));
    SNERF_LAUNCH_OK("ray_losses_fwd_kernel");
    return SNERF_OK;
}

// per-ray loss maps (validation images, src/Trainer01.py:195-196, :252-259): maps[s][i] = the ray's term divided by its channel
// count -- what compute_mse / compute_depth_loss return as `loss_maps` before averaging over rays (MSE01.py:56, :63-66)
__global__ void __launch_bounds__(kLossThreads) ray_loss_maps_kernel(const __grid_constant__ LossTable t, int n_rays) {
    for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n_rays; i += gridDim.x * kLossThreads)
        for (int s = 0; s < t.n_streams; ++s) {
            const bool in = t.mask[s] == nullptr || t.mask[s][i];
            const float per_ray = t.kind[s] == SNERF_LOSS_PRIOR_SHORTFALL ? 1.f : (float)t.channels[s];
            t.grad[s][i] = in ? ray_term(t.pred[s], t.target[s], i, t.channels[s], t.kind[s]) / per_ray : 0.f;
        }
}

extern "C" int snerf_ray_loss_maps(const snerf_loss_stream* streams, int n_streams, int n_rays, void* stream) {
    SNERF_REQUIRE(n_rays >= 0, "snerf_ray_loss_maps: bad ray count %d", n_rays);
    LossTable t{};
    if (int rc = fill_table(t, streams, n_streams, true, "snerf_ray_loss_maps")) return rc;
    if (n_rays == 0) return SNERF_OK;
    const int blocks = min(4 * kLossMaxBlocks, ceil_div(n_rays, kLossThreads));
    ray_loss_maps_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(t, n_rays);
    SNERF_LAUNCH_OK("ray_loss_maps_kernel");
    return SNERF_OK;
}

extern "C" int snerf_ray_losses_backward(const snerf_loss_stream* streams, int n_streams, int n_rays, const int32_t* counts,
                                         const float* grad_values, void* stream) {
    SNERF_REQUIRE(n_rays >= 0, "snerf_ray_losses_backward: bad ray count %d", n_rays);
    LossTable t{};
    if (int rc = fill_table(t, streams, n_streams, true, "snerf_ray_losses_backward")) return rc;
    SNERF_REQUIRE(counts && grad_values, "snerf_ray_losses_backward: null counts / incoming gradient");
    if (n_rays == 0) return SNERF_OK;
    const int blocks = min(4 * kLossMaxBlocks, ceil_div(n_rays, kLossThreads));
    ray_losses_bwd_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(t, n_rays, counts, grad_values);
    SNERF_LAUNCH_OK("ray_losses_bwd_kernel");
    return SNERF_OK;
}

static int fill_reproj(ReprojTable& t, const snerf_reproj_args* a, bool backward, const char* who) {
    SNERF_REQUIRE(a != nullptr, "%s: null arguments", who);
    SNERF_REQUIRE(a->n_others >= 1 && a->n_others <= SNERF_REPROJ_MAX_OTHERS, "%s: %d other depths (1..%d)", who, a->n_others,
                  SNERF_REPROJ_MAX_OTHERS);
    SNERF_REQUIRE(a->depth_main != nullptr, "%s: null main depth", who);
    t.depth_main = a->depth_main;
    t.grad_main = a->grad_main;
    t.n_others = a->n_others;
    for (int k = 0; k < a->n_others; ++k) {
        SNERF_REQUIRE(a->depth_other[k] != nullptr, "%s: other depth %d is null", who, k);
        t.depth_other[k] = a->depth_other[k];
        t.grad_other[k] = a->grad_other[k];
        t.weight[k] = a->weight[k];
    }
    SNERF_REQUIRE(!backward || a->grad_main, "%s: no gradient buffer for the main depth", who);
    if (!backward) {
        SNERF_REQUIRE(a->rays_o && a->rays_d && a->pixel_id && a->images && a->proj && a->origins && a->closest,
                      "%s: null ray / view table", who);
        SNERF_REQUIRE(a->n_views >= 2 && a->height >= 1 && a->width >= 1, "%s: bad view table (%d views of %dx%d)", who, a->n_views,
                      a->height, a->width);
        if (a->half_patch < 0 || a->half_patch > 2)
            return fail(SNERF_ERR_UNSUPPORTED, "%s: patches of side %d (1, 3 or 5 are built)", who, 2 * a->half_patch + 1);
    }
    t.rays_o = a->rays_o; t.rays_d = a->rays_d; t.pixel_id = a->pixel_id; t.mask_nerf = a->mask_nerf; t.images = a->images;
    t.proj = a->proj; t.origins = a->origins; t.closest = a->closest;
    t.n_views = a->n_views; t.h = a->height; t.w = a->width; t.hp = a->half_patch; t.threshold = a->rmse_threshold;
    t.symmetric = (a->flags & SNERF_REPROJ_SYMMETRIC) != 0;
    return SNERF_OK;
}

extern "C" int snerf_reprojection_losses_forward(const snerf_reproj_args* args, int n_rays, uint8_t* codes, float* values,
                                                 int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    SNERF_REQUIRE(n_rays >= 0, "snerf_reprojection_losses_forward: bad ray count %d", n_rays);
    ReprojTable t{};
    if (int rc = fill_reproj(t, args, false, "snerf_reprojection_losses_forward")) return rc;
    SNERF_REQUIRE(codes && values && counts && workspace, "snerf_reprojection_losses_forward: null output / workspace");
    SNERF_REQUIRE(workspace_bytes >= sizeof(LossWorkspace), "snerf_reprojection_losses_forward: workspace of %zu bytes < %zu",
                  workspace_bytes, sizeof(LossWorkspace));
    const int blocks = max(1, min(kLossMaxBlocks, ceil_div(n_rays, kReprojWarps)));
    reproj_fwd_kernel<<<blocks, kReprojWarps * kWarp, 0, (cudaStream_t)stream>>>(t, n_rays, codes, values, counts,
                                                                                static_cast<LossWorkspace*>(workspace));
    SNERF_LAUNCH_OK("reproj_fwd_kernel");
    return SNERF_OK;
}

extern "C" int snerf_reprojection_losses_backward(const snerf_reproj_args* args, int n_rays, const uint8_t* codes,
                                                  const int32_t* counts, const float* grad_values, void* stream) {
    SNERF_REQUIRE(n_rays >= 0, "snerf_reprojection_losses_backward: bad ray count %d", n_rays);
    ReprojTable t{};
    if (int rc = fill_reproj(t, args, true, "snerf_reprojection_losses_backward")) return rc;
    SNERF_REQUIRE(codes && counts && grad_values, "snerf_reprojection_losses_backward: null codes / counts / incoming gradient");
    if (n_rays == 0) return SNERF_OK;
    const int blocks = min(4 * kLossMaxBlocks, ceil_div(n_rays, kLossThreads));
    reproj_bwd_kernel<<<blocks, kLossThreads, 0, (cudaStream_t)stream>>>(t, n_rays, codes, counts, grad_values);
    SNERF_LAUNCH_OK("reproj_bwd_kernel");
    return SNERF_OK;
}
