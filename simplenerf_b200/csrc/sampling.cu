// Ray-depth sampling kernels: stratified coarse sampling and the inverse-CDF resampler.
//
// Reference behaviour (src/models/SimpleNeRF01.py):
//   get_z_vals_coarse :272-302, get_z_vals_fine :304-315, sample_pdf :328-361.
// Both kernels reproduce the reference's fp32 arithmetic op by op (explicit _rn intrinsics stop
// nvcc from contracting a*b+c into FMA, which the eager CPU reference never does).
#include "common.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------------------
// coarse: one thread per (ray, sample)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_depth(float near, float far, float t, bool lindisp) {
    const float one_minus_t = __fsub_rn(1.f, t);
    if (!lindisp) return __fadd_rn(__fmul_rn(near, one_minus_t), __fmul_rn(far, t));           // :287
    const float a = __fmul_rn(__fdiv_rn(1.f, near), one_minus_t);
    const float b = __fmul_rn(__fdiv_rn(1.f, far), t);
    return __fdiv_rn(1.f, __fadd_rn(a, b));                                                    // :289
}

__global__ void __launch_bounds__(256) sample_coarse_kernel(const float* __restrict__ near,
                                                            const float* __restrict__ far,
                                                            const float* __restrict__ t_vals,
                                                            const float* __restrict__ t_rand,
                                                            float* __restrict__ z_out, int n_rays, int s,
                                                            bool lindisp) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_rays * s) return;
    const int ray = (int)(gid / s), k = (int)(gid % s);
    const float nr = near[ray], fr = far[ray];
    const float zk = lerp_depth(nr, fr, t_vals[k], lindisp);
    if (t_rand == nullptr) {
        z_out[gid] = zk;
        return;
    }
    // lower = [z0, mids], upper = [mids, z_last]   (:295-297)
    float lo = zk, hi = zk;
    if (k > 0) lo = __fmul_rn(.5f, __fadd_rn(zk, lerp_depth(nr, fr, t_vals[k - 1], lindisp)));
    if (k < s - 1) hi = __fmul_rn(.5f, __fadd_rn(lerp_depth(nr, fr, t_vals[k + 1], lindisp), zk));
    z_out[gid] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t_rand[gid]));                      // :301
}

// ------------------------------------------------------------------------------------------------
// fine: one warp per ray.  smem per warp: cdf[nb] | bins[nb] | sort[npad]
// ------------------------------------------------------------------------------------------------
constexpr int kFineWarps = 4;

__global__ void __launch_bounds__(kFineWarps* kWarp)
    sample_fine_kernel(const float* __restrict__ z_coarse, const float* __restrict__ w_coarse,
                       const float* __restrict__ u, int u_stride, float* __restrict__ z_fine,
                       float* __restrict__ samples_dbg, float* __restrict__ cdf_dbg, int* __restrict__ below_dbg,
                       int* __restrict__ above_dbg, int n_rays, int sc, int n_new, int npad) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int ray = blockIdx.x * kFineWarps + warp;
    if (ray >= n_rays) return;
    const int nb = sc - 1;   // number of bins (mid points) == cdf entries
    const int nw = sc - 2;   // number of interior weights
    float* cdf = smem + (size_t)warp * (2 * nb + npad);
    float* bins = cdf + nb;
    float* sorted = bins + nb;
    const float* zc = z_coarse + (size_t)ray * sc;
    const float* wc = w_coarse + (size_t)ray * sc;

    // bins = .5 * (z[1:] + z[:-1])  (:310) ; also seed the sort buffer with the coarse depths
    for (int i = lane; i < sc; i += kWarp) {
        const float zi = zc[i];
        sorted[i] = zi;
        if (i < nb) bins[i] = __fmul_rn(.5f, __fadd_rn(zc[i + 1], zi));
    }
    for (int i = sc + n_new + lane; i < npad; i += kWarp) sorted[i] = __int_as_float(0x7f800000);

    // weights + 1e-5, normaliser.  torch's CPU sum is an fp32 cascade whose order depends on the host
    // vector width; the correctly rounded sum (fp64 accumulate) is the closest host-independent match.
    const int ipl = ceil_div(nw, kWarp);   // contiguous items per lane
    const int first = lane * ipl;
    double lane_sum = 0.0;
    for (int j = 0; j < ipl; ++j) {
        const int i = first + j;
        if (i < nw) lane_sum += (double)__fadd_rn(wc[i + 1], 1e-5f);                           // :331
    }
    double total = lane_sum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
    const float norm = (float)total;                                                            // :332

    // cdf = [0, cumsum(pdf)]: torch's CPU cumsum accumulates in fp64 and rounds every prefix to fp32
    // (SURVEY.md H2); a fp64 warp scan reproduces that.
    double lane_pdf = 0.0;
    for (int j = 0; j < ipl; ++j) {
        const int i = first + j;
        if (i < nw) lane_pdf += (double)__fdiv_rn(__fadd_rn(wc[i + 1], 1e-5f), norm);
    }
    double incl = lane_pdf;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
        const double v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    double run = incl - lane_pdf;   // exclusive prefix of the lanes before this one
    if (lane == 0) cdf[0] = 0.f;                                                                // :334
    for (int j = 0; j < ipl; ++j) {
        const int i = first + j;
        if (i < nw) {
            run += (double)__fdiv_rn(__fadd_rn(wc[i + 1], 1e-5f), norm);
            cdf[i + 1] = (float)run;                                                            // :333
        }
    }
    __syncwarp();
    if (cdf_dbg != nullptr)
        for (int i = lane; i < nb; i += kWarp) cdf_dbg[(size_t)ray * nb + i] = cdf[i];

    // invert the cdf (:345-359)
    const float* urow = u + (size_t)ray * u_stride;
    for (int s = lane; s < n_new; s += kWarp) {
        const float us = urow[s];
        int lo = 0, hi = nb;   // searchsorted(right=True): first index with cdf[idx] > u
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= us) lo = mid + 1; else hi = mid;
        }
        const int below = max(lo - 1, 0);                                                       // :346
        const int above = min(lo, nb - 1);                                                      // :347
        const float c0 = cdf[below], c1 = cdf[above];
        float denom = __fsub_rn(c1, c0);                                                        // :356
        if (denom < 1e-5f) denom = 1.f;                                                         // :357
        const float t = __fdiv_rn(__fsub_rn(us, c0), denom);                                    // :358
        const float b0 = bins[below], b1 = bins[above];
        const float smp = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));                       // :359
        sorted[sc + s] = smp;
        if (samples_dbg != nullptr) samples_dbg[(size_t)ray * n_new + s] = smp;
        if (below_dbg != nullptr) below_dbg[(size_t)ray * n_new + s] = below;
        if (above_dbg != nullptr) above_dbg[(size_t)ray * n_new + s] = above;
    }
    __syncwarp();

    // sort(cat(z_coarse, samples)) (:314): in-smem bitonic network over npad (power of two) slots
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < npad / 2; i += kWarp) {
                const int a = 2 * j * (i / j) + (i % j);
                const int b = a + j;
                const float va = sorted[a], vb = sorted[b];
                const bool up = (a & k) == 0;
                if ((va > vb) == up) {
                    sorted[a] = vb;
                    sorted[b] = va;
                }
            }
            __syncwarp();
        }
    }
    const int tot = sc + n_new;
    for (int i = lane; i < tot; i += kWarp) z_fine[(size_t)ray * tot + i] = sorted[i];
}

}  // namespace snerf

using namespace snerf;

extern "C" int snerf_sample_coarse(const float* near, const float* far, const float* t_vals, const float* t_rand,
                                   float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream) {
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "snerf_sample_coarse: bad sizes (%d rays, %d samples)", n_rays, n_samples);
    if (n_rays == 0) return SNERF_OK;   // empty batch: nothing to launch (pointers may be null)
    SNERF_REQUIRE(near && far && t_vals && z_out, "snerf_sample_coarse: null pointer");
    const long long total = (long long)n_rays * n_samples;
    const int blocks = (int)((total + 255) / 256);
    sample_coarse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, t_vals, t_rand, z_out, n_rays, n_samples,
                                                                   (flags & SNERF_FLAG_LINDISP) != 0);
    SNERF_LAUNCH_OK("sample_coarse_kernel");
    return SNERF_OK;
}

extern "C" int snerf_sample_fine(const float* z_coarse, const float* weights_coarse, const float* u, int u_stride,
                                 float* z_fine, float* samples_dbg, float* cdf_dbg, int32_t* below_dbg,
                                 int32_t* above_dbg, int n_rays, int s_coarse, int n_new, void* stream) {
    SNERF_REQUIRE(n_rays >= 0 && s_coarse >= 3 && n_new >= 1, "snerf_sample_fine: bad sizes (%d rays, %d coarse, %d new)",
                  n_rays, s_coarse, n_new);
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(z_coarse && weights_coarse && u && z_fine, "snerf_sample_fine: null pointer");
    SNERF_REQUIRE(u_stride == 0 || u_stride >= n_new, "snerf_sample_fine: u_stride %d < n_new %d", u_stride, n_new);
    if (s_coarse + n_new > 1024) return fail(SNERF_ERR_UNSUPPORTED, "snerf_sample_fine: %d + %d samples > 1024", s_coarse, n_new);
    if (n_rays == 0) return SNERF_OK;
    int npad = 2;
    while (npad < s_coarse + n_new) npad <<= 1;
    const size_t smem = (size_t)kFineWarps * (2 * (s_coarse - 1) + npad) * sizeof(float);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        SNERF_CUDA_OK(cudaFuncSetAttribute(sample_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    sample_fine_kernel<<<ceil_div(n_rays, kFineWarps), kFineWarps * kWarp, smem, (cudaStream_t)stream>>>(
        z_coarse, weights_coarse, u, u_stride, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg, n_rays, s_coarse,
        n_new, npad);
    SNERF_LAUNCH_OK("sample_fine_kernel");
    return SNERF_OK;
}
