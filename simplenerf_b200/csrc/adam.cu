// One-launch Adam over a table of tensors (next-row N2: the optimizer step of the trainer, src/Trainer01.py:516).
// torch.optim.Adam(fused=True) spends ~80 us per launch on this model's 57 small tensors (multi_tensor_apply chunking);
// the update itself is 63 MB of traffic.  Here every block takes one 8192-element chunk of one tensor; the chunk -> tensor
// map is a prefix table in the kernel parameters.
#include <math.h>

#include "common.cuh"

namespace snerf {

constexpr int kAdamChunk = 2048;     // 2 K elements per block: ~1100 blocks for the 2.27 M parameters (8 K left the loads latency-bound)
constexpr int kAdamThreads = 256;

struct AdamTable {
    float* p[SNERF_ADAM_MAX_TENSORS];
    const float* g[SNERF_ADAM_MAX_TENSORS];
    float* m[SNERF_ADAM_MAX_TENSORS];
    float* v[SNERF_ADAM_MAX_TENSORS];
    int n[SNERF_ADAM_MAX_TENSORS];
    int chunk0[SNERF_ADAM_MAX_TENSORS + 1];
    int n_tensors;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float one_minus_b1, float b2, float one_minus_b2,
                                         float step_size, float inv_sqrt_bc2, float eps) {
    m = m + (g - m) * one_minus_b1;                       // exp_avg.lerp_(grad, 1 - beta1)
    v = b2 * v + one_minus_b2 * g * g;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamTable t, float one_minus_b1, float b2,
                                                                  float one_minus_b2, float step_size, float inv_sqrt_bc2, float eps) {
    int lo = 0, hi = t.n_tensors;                          // last tensor whose first chunk is <= blockIdx.x
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t.chunk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const int base = ((int)blockIdx.x - t.chunk0[lo]) * kAdamChunk;
    const int n = min(kAdamChunk, t.n[lo] - base);
    float* p = t.p[lo] + base;
    const float* g = t.g[lo] + base;
    float* m = t.m[lo] + base;
    float* v = t.v[lo] + base;
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += kAdamThreads) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam_one(pp.x, gg.x, mm.x, vv.x, one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps);
        adam_one(pp.y, gg.y, mm.y, vv.y, one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps);
        adam_one(pp.z, gg.z, mm.z, vv.z, one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps);
        adam_one(pp.w, gg.w, mm.w, vv.w, one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += kAdamThreads) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_one(pp, g[i], mm, vv, one_minus_b1, b2, one_minus_b2, step_size, inv_sqrt_bc2, eps);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                               const long long* numel, int n_tensors, float lr, float beta1, float beta2, float eps, int step,
                               void* stream) {
    SNERF_REQUIRE(n_tensors >= 0 && n_tensors <= SNERF_ADAM_MAX_TENSORS, "snerf_adam_step: %d tensors (max %d per call)", n_tensors,
                  SNERF_ADAM_MAX_TENSORS);
    SNERF_REQUIRE(step >= 1, "snerf_adam_step: step counts from 1");
    if (n_tensors == 0) return SNERF_OK;
    SNERF_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel, "snerf_adam_step: null table");
    AdamTable t{};
    int chunks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        SNERF_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "snerf_adam_step: tensor %d has a null pointer", i);
        SNERF_REQUIRE(((reinterpret_cast<uintptr_t>(params[i]) | reinterpret_cast<uintptr_t>(grads[i]) |
                        reinterpret_cast<uintptr_t>(exp_avg[i]) | reinterpret_cast<uintptr_t>(exp_avg_sq[i])) & 15) == 0,
                      "snerf_adam_step: tensor %d is not 16-byte aligned", i);
        SNERF_REQUIRE(numel[i] >= 0 && numel[i] < (1LL << 31), "snerf_adam_step: tensor %d has a bad size", i);
        t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i];
        t.n[i] = (int)numel[i];
        t.chunk0[i] = chunks;
        chunks += (int)((numel[i] + kAdamChunk - 1) / kAdamChunk);
    }
    t.chunk0[n_tensors] = chunks;
    t.n_tensors = n_tensors;
    if (chunks == 0) return SNERF_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_multi_kernel<<<chunks, kAdamThreads, 0, (cudaStream_t)stream>>>(t, 1.f - beta1, beta2, 1.f - beta2, (float)((double)lr / bc1),
                                                                         (float)(1.0 / sqrt(bc2)), eps);
    SNERF_LAUNCH_OK("adam_multi_kernel");
    return SNERF_OK;
}
