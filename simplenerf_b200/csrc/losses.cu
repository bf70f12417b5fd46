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
    int n_streams;
};

// workspace: [blocks][streams] partial sums, [blocks][streams] partial counts, one ticket counter
struct LossWorkspace {
    float sums[kLossMaxBlocks][SNERF_LOSS_MAX_STREAMS];
    int counts[kLossMaxBlocks][SNERF_LOSS_MAX_STREAMS];
    unsigned int ticket;
};

__device__ __forceinline__ float sq_err(const float* __restrict__ p, const float* __restrict__ t, int i, int c) {
    float e = 0.f;
    for (int k = 0; k < c; ++k) {
        const float d = p[(size_t)i * c + k] - t[(size_t)i * c + k];
        e += d * d;
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
                sum += sq_err(t.pred[s], t.target[s], i, t.channels[s]);
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
        const float v = cnt > 0 ? sum / ((float)cnt * (float)t.channels[threadIdx.x]) : 0.f;   // MSE01.py:58
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
        const float scale = cnt > 0 ? coeff * 2.f / ((float)cnt * (float)c) : 0.f;
        const uint8_t* m = t.mask[s];
        const float* p = t.pred[s];
        const float* tg = t.target[s];
        float* g = t.grad[s];
        for (int i = blockIdx.x * kLossThreads + threadIdx.x; i < n_rays; i += gridDim.x * kLossThreads) {
            const bool on = m == nullptr || m[i];
            for (int k = 0; k < c; ++k) {
                const size_t o = (size_t)i * c + k;
                g[o] = on ? scale * (p[o] - tg[o]) : 0.f;
            }
        }
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
        SNERF_REQUIRE(streams[s].channels >= 1 && streams[s].channels <= 4, "%s: stream %d has %d channels (1..4)", who, s,
                      streams[s].channels);
        t.pred[s] = streams[s].pred; t.target[s] = streams[s].target; t.mask[s] = streams[s].mask;
        t.grad[s] = streams[s].grad; t.weight[s] = streams[s].weight; t.channels[s] = streams[s].channels;
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
