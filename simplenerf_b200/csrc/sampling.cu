// Ray-depth sampling kernels: stratified coarse sampling and the inverse-CDF resampler (HBM-bound).
//
// Reference behaviour (src/models/SimpleNeRF01.py):
//   get_z_vals_coarse :272-302, get_z_vals_fine :304-315, sample_pdf :328-361.
// Both kernels reproduce the reference's fp32 arithmetic op by op (explicit _rn intrinsics stop
// nvcc from contracting a*b+c into FMA, which the eager CPU reference never does).
#include "common.cuh"
#include "rng.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------------------
// coarse: one thread per 4 consecutive samples (vector loads / stores)
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
                                                            bool lindisp, const RngKey rng, const bool use_rng) {
    // generic path: one thread per sample
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_rays * s) return;
    const int ray = (int)(gid / s), k = (int)(gid % s);
    const float nr = near[ray], fr = far[ray];
    const float zk = lerp_depth(nr, fr, t_vals[k], lindisp);
    if (t_rand == nullptr && !use_rng) {
        z_out[gid] = zk;
        return;
    }
    const float tr = use_rng ? rng_pick(rng_uniform4(rng, (unsigned long long)gid >> 2), (unsigned long long)gid) : t_rand[gid];
    // lower = [z0, mids], upper = [mids, z_last]   (:295-297)
    float lo = zk, hi = zk;
    if (k > 0) lo = __fmul_rn(.5f, __fadd_rn(zk, lerp_depth(nr, fr, t_vals[k - 1], lindisp)));
    if (k < s - 1) hi = __fmul_rn(.5f, __fadd_rn(lerp_depth(nr, fr, t_vals[k + 1], lindisp), zk));
    z_out[gid] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), tr));                               // :301
}

// s % 4 == 0: thread handles samples 4q..4q+3 of one ray; q4 = s / 4 threads per ray.  The kernel is issue-bound (ncu: 66 %
// issue slots at 0.67 of the HBM bandwidth), so the model's own shape (64 samples, linear in depth) is a compile-time
// specialisation: no runtime division for the ray index, no disparity branch in the six interpolations.
template <int S_FIXED, bool LINDISP_RT>
__global__ void __launch_bounds__(256) sample_coarse_vec4_kernel(const float* __restrict__ near,
                                                                 const float* __restrict__ far,
                                                                 const float* __restrict__ t_vals,
                                                                 const float* __restrict__ t_rand,
                                                                 float* __restrict__ z_out, int n_rays, int s_rt, int q4_rt,
                                                                 bool lindisp_rt, const RngKey rng, const bool use_rng) {
    const int s = S_FIXED > 0 ? S_FIXED : s_rt;
    const int q4 = S_FIXED > 0 ? S_FIXED / 4 : q4_rt;
    const bool lindisp = LINDISP_RT ? lindisp_rt : false;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned ray = gid / (unsigned)q4;
    if (ray >= (unsigned)n_rays) return;
    const int k0 = (int)(gid - ray * q4) * 4;
    const float nr = near[ray], fr = far[ray];
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(t_vals + k0));
    float zc[6];   // depths k0-1 .. k0+4
    zc[1] = lerp_depth(nr, fr, t4.x, lindisp);
    zc[2] = lerp_depth(nr, fr, t4.y, lindisp);
    zc[3] = lerp_depth(nr, fr, t4.z, lindisp);
    zc[4] = lerp_depth(nr, fr, t4.w, lindisp);
    float out[4] = {zc[1], zc[2], zc[3], zc[4]};
    if (t_rand != nullptr || use_rng) {
        zc[0] = k0 > 0 ? lerp_depth(nr, fr, t_vals[k0 - 1], lindisp) : zc[1];
        zc[5] = k0 + 4 < s ? lerp_depth(nr, fr, t_vals[k0 + 4], lindisp) : zc[4];
        // in-kernel draw: element ray * s + k0 is a multiple of four, i.e. one Philox block per thread
        const float4 r4 = use_rng ? rng_uniform4(rng, ((unsigned long long)ray * s + k0) >> 2)
                                  : __ldg(reinterpret_cast<const float4*>(t_rand + (size_t)ray * s + k0));
        const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + i;
            const float zk = zc[i + 1];
            const float lo = k > 0 ? __fmul_rn(.5f, __fadd_rn(zk, zc[i])) : zk;                 // :295-297
            const float hi = k < s - 1 ? __fmul_rn(.5f, __fadd_rn(zc[i + 2], zk)) : zk;
            out[i] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), rr[i]));                        // :301
        }
    }
    *reinterpret_cast<float4*>(z_out + (size_t)ray * s + k0) = make_float4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------------------------------------
// fine, generic shapes: one warp per ray.  smem per warp: cdf[nb] | bins[nb] | sort[npad]
// ------------------------------------------------------------------------------------------------
constexpr int kFineWarps = 4;

__global__ void __launch_bounds__(kFineWarps* kWarp)
    sample_fine_generic_kernel(const float* __restrict__ z_coarse, const float* __restrict__ w_coarse,
                               const float* __restrict__ u, int u_stride, const RngKey rng, const bool use_rng, float* __restrict__ z_fine,
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

    for (int i = lane; i < sc; i += kWarp) {
        const float zi = zc[i];
        sorted[i] = zi;
        if (i < nb) bins[i] = __fmul_rn(.5f, __fadd_rn(zc[i + 1], zi));                            // :310
    }
    for (int i = sc + n_new + lane; i < npad; i += kWarp) sorted[i] = __int_as_float(0x7f800000);

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
    double run = incl - lane_pdf;
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

    const float* urow = use_rng ? nullptr : u + (size_t)ray * u_stride;
    for (int s = lane; s < n_new; s += kWarp) {
        const unsigned long long el = (unsigned long long)ray * n_new + s;
        const float us = use_rng ? rng_pick(rng_uniform4(rng, el >> 2), el) : urow[s];
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
    for (int k = 2; k <= npad; k <<= 1) {                                                       // :314
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

// ------------------------------------------------------------------------------------------------
// fine, the model's shape (64 coarse depths, 32*NPL new samples): values in registers, a few small smem rows per ray.
// The kernel is NOT HBM-bound: 1.8 KB of traffic per ray against ~540 (ordered uniforms) / ~780 (random uniforms) warp
// instructions and ~120 shared-memory wavefronts; ncu shows 75-82 % of the issue slots and 76-80 % of the L1 data pipe
// in use.  Every stage is therefore written for instruction and wavefront count:
//   * the cdf stage (fp64 reduction and scan, IEEE divisions) runs for two rays at once, one per half warp;
//   * cdf and mid points interleaved as float2 -> the two gathers of a sample are two LDS.64;
//   * searchsorted(right=True) over the 63 cdf entries and the coarse depths' ranks among the sorted samples are
//     branch-free descents over breadth-first copies of the tables (LDS, FSETP, LEA, predicated add; no bank conflicts);
//   * the new samples are sorted (only when they are not already in order) by a bitonic network in its "flip"
//     form -- every compare-exchange is ascending, so the direction is a compile-time constant inside a lane and
//     one lane-bit predicate across lanes;
//   * the merge with the (sorted) coarse depths is by rank from the coarse side only: coarse depth k belongs in slot
//     k + #(samples < depth); the occupancy of the merged row is OR-reduced into TOT/32 words, from which every lane
//     derives which value belongs in the slots it stores (coalesced float2 rows): no scatter, no scan, no staging row;
//   * a zero numerator is kept out of the division (its slow path costs the whole warp ~110 instructions).
// ------------------------------------------------------------------------------------------------
constexpr int kFastWarps = 8;
constexpr int kSc = 64;

// Shared-memory accesses by 32-bit shared-space address.  The descents below keep absolute addresses in registers so
// that a step is LDS [reg+imm], FSETP, predicated add (written with C++ pointers, nvcc re-derives base + offset and
// spends a fourth instruction per step).  volatile + "memory": ordered against the surrounding stores / __syncwarp.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}

template <int NPL>
__device__ __forceinline__ void bitonic_sort_registers(float (&v)[NPL], int lane) {
    constexpr int N = NPL * kWarp;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
        // flip stage: element idx meets idx ^ (k-1); the lower index keeps the minimum
        if (k <= NPL) {
#pragma unroll
            for (int r = 0; r < NPL; ++r) {
                const int q = r ^ (k - 1);
                if (r < q) {
                    const float a = v[r], b = v[q];
                    v[r] = fminf(a, b);
                    v[q] = fmaxf(a, b);
                }
            }
        } else {
            const int m = k / NPL - 1;                        // partner lane = lane ^ m, partner register = NPL-1-r
            const bool lower = (lane & ((m + 1) >> 1)) == 0;  // top bit of the mask decides who is the lower index
            float o[NPL];
#pragma unroll
            for (int r = 0; r < NPL; ++r) o[r] = __shfl_xor_sync(kFull, v[NPL - 1 - r], m);
#pragma unroll
            for (int r = 0; r < NPL; ++r) v[r] = lower ? fminf(v[r], o[r]) : fmaxf(v[r], o[r]);
        }
        // half cleaners: idx meets idx ^ j, ascending
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j >= NPL) {
                const int d = j / NPL;
                const bool lower = (lane & d) == 0;
#pragma unroll
                for (int r = 0; r < NPL; ++r) {
                    const float o = __shfl_xor_sync(kFull, v[r], d);
                    v[r] = lower ? fminf(v[r], o) : fmaxf(v[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < NPL; ++r) {
                    if ((r & j) == 0) {
                        const float a = v[r], b = v[r + j];
                        v[r] = fminf(a, b);
                        v[r + j] = fmaxf(a, b);
                    }
                }
            }
        }
    }
}

template <int NPL, int WARPS>
__global__ void __launch_bounds__(WARPS* kWarp)
    sample_fine_fast_kernel(const float* __restrict__ z_coarse, const float* __restrict__ w_coarse,
                            const float* __restrict__ u, int u_stride, const RngKey rng, const bool use_rng, float* __restrict__ z_fine,
                            float* __restrict__ samples_dbg, float* __restrict__ cdf_dbg, int* __restrict__ below_dbg,
                            int* __restrict__ above_dbg, int n_rays) {
    constexpr int NNEW = NPL * kWarp, TOT = kSc + NNEW, NB = kSc - 1;
    static_assert(TOT % 64 == 0 && (NNEW & (NNEW - 1)) == 0, "row shapes");
    struct __align__(16) Row {
        float2 cb[kSc];      // (cdf[k], bins[k]), 63 used: the two gathers of a sample
        float ss[NNEW];      // sorted new samples
        float zc[kSc];       // coarse depths (must follow ss: the merge reads both through one index)
        float ecdf[kSc];     // cdf[0..62] as an implicit search tree in breadth-first order (root at [1])
        float ess[NNEW];     // ss[0..NNEW-2] likewise
    };
    // Breadth-first ("Eytzinger") copies for the two descents: the nodes a warp can touch at level k are 2^k
    // consecutive words, so 32 lanes never meet in a bank on different words (a descent over the sorted array reads
    // nodes 2^(6-k) words apart -- up to 8 lanes per bank; ncu: 117 conflict cycles per ray, L1 data pipe saturated).
    constexpr int LOGN = NPL == 2 ? 6 : NPL == 4 ? 7 : 8;
    static_assert((1 << LOGN) == NNEW, "LOGN");
    // Two rays per warp.  The cdf stage is per-warp overhead (fp64 reduction and scan, IEEE divisions): it runs once for
    // both rays, one per half warp with four coarse entries per lane (four shuffle steps instead of five, half the
    // instructions per ray); everything after it needs all 32 lanes per ray and runs for one ray after the other.
    __shared__ Row s_rows[WARPS][2];
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int ray0 = (blockIdx.x * WARPS + warp) * 2;
    if (ray0 >= n_rays) return;
    const int n_here = min(2, n_rays - ray0);
    const int tz1 = __ffs(lane + 1) - 1;                              // ctz(lane + 1)
    {
        const int half = lane >> 4, gl = lane & 15;
        const int fray = ray0 + min(half, n_here - 1);                // (an odd last ray is prepared twice)
        Row& frow = s_rows[warp][half];
        // lane owns coarse indices 4*gl .. 4*gl+3
        const float4 z4 = __ldg(reinterpret_cast<const float4*>(z_coarse + (size_t)fray * kSc) + gl);
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_coarse + (size_t)fray * kSc) + gl);
        const float znext = __shfl_down_sync(kFull, z4.x, 1);         // (gl == 15: another ray's value, bin 63 is unused)
        *reinterpret_cast<float4*>(frow.zc + 4 * gl) = z4;
        const float zs[5] = {z4.x, z4.y, z4.z, z4.w, znext};
        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
        float bin[4], wk[4], pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bin[i] = __fmul_rn(.5f, __fadd_rn(zs[i + 1], zs[i]));                               // :310
            const bool interior = !((gl == 0 && i == 0) || (gl == 15 && i == 3));               // weights k = 1 .. 62
            wk[i] = interior ? __fadd_rn(ws[i], 1e-5f) : 0.f;                                   // :331
        }
        double total = ((double)wk[0] + (double)wk[1]) + ((double)wk[2] + (double)wk[3]);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
        const float norm = (float)total;                                                        // :332
#pragma unroll
        for (int i = 0; i < 4; ++i) pk[i] = wk[i] != 0.f ? __fdiv_rn(wk[i], norm) : 0.f;
        // cdf[k] = sum_{k' <= k} pdf(k'): torch's CPU cumsum keeps an fp64 running sum, rounded per prefix  :333
        double run[4];
        run[0] = (double)pk[0];
#pragma unroll
        for (int i = 1; i < 4; ++i) run[i] = run[i - 1] + (double)pk[i];
        double incl = run[3];
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const double v = __shfl_up_sync(kFull, incl, o);
            if (gl >= o) incl += v;
        }
        const double before = incl - run[3];
        float cdf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) cdf[i] = (float)(before + run[i]);                          // cdf[0] = 0  (:334)
        if (gl == 15) cdf[3] = __int_as_float(0x7f800000);                                      // entry 63: never read
        *reinterpret_cast<float4*>(&frow.cb[4 * gl]) = make_float4(cdf[0], bin[0], cdf[1], bin[1]);
        *reinterpret_cast<float4*>(&frow.cb[4 * gl + 2]) = make_float4(cdf[2], bin[2], cdf[3], bin[3]);
        // sorted index j = node t = j + 1 of the in-order numbering: level = 5 - ctz(t), slot = 2^level + (t >> (ctz(t)+1))
        const int tzg = __ffs(gl + 1) - 1;                            // ctz(gl + 1)
        *reinterpret_cast<float2*>(frow.ecdf + 32 + 2 * gl) = make_float2(cdf[0], cdf[2]);      // t = 4gl+1, 4gl+3: bottom level
        frow.ecdf[16 + gl] = cdf[1];                                                            // t = 4gl+2
        if (gl < 15) frow.ecdf[(8 >> tzg) + ((gl + 1) >> (tzg + 1))] = cdf[3];                  // t = 4(gl+1)
        if (cdf_dbg != nullptr && half < n_here) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (4 * gl + i < NB) cdf_dbg[(size_t)fray * NB + 4 * gl + i] = cdf[i];
        }
    }
    __syncwarp();

  for (int which = 0; which < n_here; ++which) {
    const int ray = ray0 + which;
    Row& row = s_rows[warp][which];
    const float2 z2 = *reinterpret_cast<const float2*>(row.zc + 2 * lane);     // this lane's coarse depths 2*lane, 2*lane+1
    float us[NPL];
    if (use_rng) {     // in-kernel draw (:341): this lane's NPL consecutive elements of the ray's row
        const unsigned long long el = (unsigned long long)ray * NNEW + lane * NPL;
        if constexpr (NPL % 4 == 0) {
#pragma unroll
            for (int i = 0; i < NPL / 4; ++i) {
                const float4 t = rng_uniform4(rng, (el >> 2) + i);
                us[4 * i] = t.x; us[4 * i + 1] = t.y; us[4 * i + 2] = t.z; us[4 * i + 3] = t.w;
            }
        } else {
            const float4 t = rng_uniform4(rng, el >> 2);
            us[0] = (el & 2ull) ? t.z : t.x;
            us[1] = (el & 2ull) ? t.w : t.y;
        }
    } else {
        const float* urow = u + (size_t)ray * u_stride + lane * NPL;
        if constexpr (NPL % 4 == 0) {
#pragma unroll
            for (int i = 0; i < NPL / 4; ++i) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(urow) + i);
                us[4 * i] = t.x; us[4 * i + 1] = t.y; us[4 * i + 2] = t.z; us[4 * i + 3] = t.w;
            }
        } else {
            const float2 t = __ldg(reinterpret_cast<const float2*>(urow));
            us[0] = t.x; us[1] = t.y;
        }
    }
    // ---- invert the cdf for this lane's NPL uniforms (:345-359) ----
    // searchsorted(right=True) = #(cdf[j] <= u) over 63 = 2^6 - 1 sorted entries: 6-step descent on byte offsets
    // node i -> 2i + (cdf <= u); after 6 levels i - 64 = #(cdf <= u).  On byte addresses a = E + 4i: a' = 2a - E (+4)
    const uint32_t cbp = smem_addr(row.cb), ecp = smem_addr(row.ecdf);
    uint32_t off[NPL];   // address of entry #(cdf <= u) of cb
#pragma unroll
    for (int r = 0; r < NPL; ++r) off[r] = ecp + 4;
#pragma unroll
    for (int level = 0; level < 6; ++level) {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            const float c = lds_f32(off[r]);
            off[r] = 2 * off[r] - ecp;
            if (c <= us[r]) off[r] += 4;
        }
    }
#pragma unroll
    for (int r = 0; r < NPL; ++r) off[r] = 2 * (off[r] - ecp) - 64 * 8 + cbp;
    float smp[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        const float2 lo = lds_f32x2(max(off[r], cbp + 8) - 8);                                  // below  :346
        const float2 hi = lds_f32x2(min(off[r], cbp + (NB - 1) * 8));                           // above  :347
        float denom = __fsub_rn(hi.x, lo.x);                                                    // :356
        if (denom < 1e-5f) denom = 1.f;                                                         // :357
        // (u - cdf_below) == 0 -- every ray's first sample under the deterministic linspace -- would send the whole warp
        //  through the division's slow path (FCHK flags a zero numerator): ~110 instructions per ray.  0 / denom = +0.
        const float num = __fsub_rn(us[r], lo.x);
        const float q = __fdiv_rn(num == 0.f ? 1.f : num, denom);
        const float t = num == 0.f ? 0.f : q;                                                   // :358
        smp[r] = __fadd_rn(lo.y, __fmul_rn(t, __fsub_rn(hi.y, lo.y)));                          // :359
    }
    if (samples_dbg != nullptr) {     // parity-test outputs, in the caller's order of uniforms
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            samples_dbg[(size_t)ray * NNEW + lane * NPL + r] = smp[r];
            const int pos = (int)(off[r] - cbp) / 8;
            if (below_dbg != nullptr) below_dbg[(size_t)ray * NNEW + lane * NPL + r] = max(pos - 1, 0);
            if (above_dbg != nullptr) above_dbg[(size_t)ray * NNEW + lane * NPL + r] = min(pos, NB - 1);
        }
    }

    // ---- sort the new samples (values only) (:314) ----
    // Sorted uniforms (the deterministic linspace of eval / rendering) give samples that are already in order: the
    // inverse cdf is monotone.  That is checked on the samples themselves (rounding could break it by an ulp), and
    // the 128-element bitonic network is skipped when it holds.
    {
        bool in_order = true;
#pragma unroll
        for (int r = 0; r + 1 < NPL; ++r) in_order &= smp[r] <= smp[r + 1];
        const float nxt = __shfl_down_sync(kFull, smp[0], 1);
        if (lane < kWarp - 1) in_order &= smp[NPL - 1] <= nxt;
        if (!__all_sync(kFull, in_order)) bitonic_sort_registers<NPL>(smp, lane);
    }
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int i = 0; i < NPL / 4; ++i)
            *reinterpret_cast<float4*>(row.ss + lane * NPL + 4 * i) = make_float4(smp[4 * i], smp[4 * i + 1], smp[4 * i + 2], smp[4 * i + 3]);
    } else {
        *reinterpret_cast<float2*>(row.ss + lane * NPL) = make_float2(smp[0], smp[1]);
    }
    // breadth-first copy of ss[0 .. NNEW-2] (ss[NNEW-1] is compared separately)
    const float ss_last = __shfl_sync(kFull, smp[NPL - 1], kWarp - 1);
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        const int t = lane * NPL + r + 1;
        const int tz = r + 1 < NPL ? __ffs(r + 1) - 1 : __ffs(NPL) - 1 + tz1;      // ctz(t)
        if (r + 1 < NPL || lane < kWarp - 1) row.ess[((NNEW / 2) >> tz) + (t >> (tz + 1))] = smp[r];
    }
    __syncwarp();

    // ---- merge: coarse depth k belongs in slot k + #(samples < depth); lower_bound over NNEW = 2^m sorted samples ----
    {
        const uint32_t esp = smem_addr(row.ess);
        const float zv[2] = {z2.x, z2.y};
        uint32_t p[2] = {esp + 4, esp + 4};
#pragma unroll
        for (int level = 0; level < LOGN; ++level) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float c = lds_f32(p[q]);
                p[q] = 2 * p[q] - esp;
                if (c < zv[q]) p[q] += 4;
            }
        }
        // (p - esp) / 4 - NNEW = #(ss[0..NNEW-2] < z), + the last sample: slot of coarse depth k = k + #(samples < depth)
        uint32_t slot[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (ss_last < zv[q]) p[q] += 4;
            slot[q] = ((p[q] - esp) >> 2) + (uint32_t)(2 * lane + q - NNEW);
        }
        // Occupancy of the merged row as TOT/32 words (one warp OR-reduction each): bit s = slot s holds a coarse depth.
        // From it every lane derives, for the slots it will STORE (64 i + 2 lane, + 1: coalesced float2 rows), how many
        // coarse depths precede the slot -- which is both the coarse index (if the slot is a coarse one) and, subtracted
        // from the slot, the sample index.  No scatter, no scan, no trip of the row through shared memory.
        constexpr int W = TOT / 32;
        uint32_t occ[W];
        int before[W + 1];
        before[0] = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const uint32_t mine = ((slot[0] >> 5) == (uint32_t)j ? 1u << (slot[0] & 31) : 0u) |
                                  ((slot[1] >> 5) == (uint32_t)j ? 1u << (slot[1] & 31) : 0u);
            occ[j] = __reduce_or_sync(kFull, mine);
            before[j + 1] = before[j] + __popc(occ[j]);
        }
        // one table: sorted samples, then the coarse depths (zc follows ss inside the row).  Coarse depths that are not
        // sorted collide on a slot and leave fewer than 64 bits: the sample index then runs on into zc, still inside the row.
        const uint32_t tab = smem_addr(row.ss);
        const bool upper = lane >= 16;
        const int bit = (2 * lane) & 31;
        float* dst = z_fine + (size_t)ray * TOT;
#pragma unroll
        for (int i = 0; i < TOT / 64; ++i) {
            const uint32_t word = upper ? occ[2 * i + 1] : occ[2 * i];
            const int r0 = (upper ? before[2 * i + 1] : before[2 * i]) + __popc(word & ((1u << bit) - 1u));
            const int is0 = (word >> bit) & 1, is1 = (word >> (bit + 1)) & 1;
            const int r1 = r0 + is0;
            const int s0 = 64 * i + 2 * lane;
            const int i0 = is0 ? NNEW + r0 : s0 - r0;
            const int i1 = is1 ? NNEW + r1 : s0 + 1 - r1;
            const float v0 = lds_f32(tab + 4 * i0), v1 = lds_f32(tab + 4 * i1);
            *reinterpret_cast<float2*>(dst + s0) = make_float2(v0, v1);
        }
    }
  }   // rays of this warp
}

}  // namespace snerf

using namespace snerf;

static int sample_coarse_impl(const float* near, const float* far, const float* t_vals, const float* t_rand, const RngKey rng,
                              const bool use_rng, float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream, const char* who) {
    SNERF_REQUIRE(n_rays >= 0 && n_samples >= 1, "%s: bad sizes (%d rays, %d samples)", who, n_rays, n_samples);
    if (n_rays == 0) return SNERF_OK;   // empty batch: nothing to launch (pointers may be null)
    SNERF_REQUIRE(near && far && t_vals && z_out, "%s: null pointer", who);
    const bool lindisp = (flags & SNERF_FLAG_LINDISP) != 0;
    const long long total = (long long)n_rays * n_samples;
    const bool aligned = ((reinterpret_cast<uintptr_t>(t_vals) | reinterpret_cast<uintptr_t>(t_rand) |
                           reinterpret_cast<uintptr_t>(z_out)) & 15) == 0;
    if (n_samples % 4 == 0 && aligned && total / 4 < (1LL << 31)) {
        const int q4 = n_samples / 4;
        const long long threads = (long long)n_rays * q4;
        const unsigned blocks = (unsigned)((threads + 255) / 256);
        if (n_samples == 64 && !lindisp)
            sample_coarse_vec4_kernel<64, false><<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, t_vals, t_rand, z_out, n_rays,
                                                                                         n_samples, q4, lindisp, rng, use_rng);
        else
            sample_coarse_vec4_kernel<0, true><<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, t_vals, t_rand, z_out, n_rays,
                                                                                        n_samples, q4, lindisp, rng, use_rng);
        SNERF_LAUNCH_OK("sample_coarse_vec4_kernel");
        return SNERF_OK;
    }
    const int blocks = (int)((total + 255) / 256);
    sample_coarse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(near, far, t_vals, t_rand, z_out, n_rays, n_samples, lindisp, rng, use_rng);
    SNERF_LAUNCH_OK("sample_coarse_kernel");
    return SNERF_OK;
}

extern "C" int snerf_sample_coarse(const float* near, const float* far, const float* t_vals, const float* t_rand,
                                   float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream) {
    return sample_coarse_impl(near, far, t_vals, t_rand, RngKey{0, 0}, false, z_out, n_rays, n_samples, flags, stream, "snerf_sample_coarse");
}

extern "C" int snerf_sample_coarse_rng(const float* near, const float* far, const float* t_vals, uint64_t seed, uint64_t offset,
                                       float* z_out, int n_rays, int n_samples, uint32_t flags, void* stream) {
    return sample_coarse_impl(near, far, t_vals, nullptr, RngKey{seed, offset}, true, z_out, n_rays, n_samples, flags, stream,
                              "snerf_sample_coarse_rng");
}

static int sample_fine_impl(const float* z_coarse, const float* weights_coarse, const float* u, int u_stride, const RngKey rng,
                            const bool use_rng, float* z_fine, float* samples_dbg, float* cdf_dbg, int32_t* below_dbg,
                            int32_t* above_dbg, int n_rays, int s_coarse, int n_new, void* stream, const char* who) {
    SNERF_REQUIRE(n_rays >= 0 && s_coarse >= 3 && n_new >= 1, "%s: bad sizes (%d rays, %d coarse, %d new)", who, n_rays, s_coarse, n_new);
    if (n_rays == 0) return SNERF_OK;
    SNERF_REQUIRE(z_coarse && weights_coarse && (u || use_rng) && z_fine, "%s: null pointer", who);
    SNERF_REQUIRE(use_rng || u_stride == 0 || u_stride >= n_new, "%s: u_stride %d < n_new %d", who, u_stride, n_new);
    if (s_coarse + n_new > 1024) return fail(SNERF_ERR_UNSUPPORTED, "%s: %d + %d samples > 1024", who, s_coarse, n_new);
    const bool aligned = ((reinterpret_cast<uintptr_t>(z_coarse) | reinterpret_cast<uintptr_t>(weights_coarse) |
                           reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(z_fine)) & 15) == 0 && u_stride % 4 == 0;
    if (s_coarse == kSc && aligned && (n_new == 64 || n_new == 128 || n_new == 256)) {
        const cudaStream_t st = (cudaStream_t)stream;
        // two rays per warp; the 256-sample row needs 8.5 KB of shared memory per warp, so its blocks have four warps
        if (n_new == 64)
            sample_fine_fast_kernel<2, kFastWarps><<<ceil_div(n_rays, 2 * kFastWarps), kFastWarps * kWarp, 0, st>>>(
                z_coarse, weights_coarse, u, u_stride, rng, use_rng, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg, n_rays);
        else if (n_new == 128)
            sample_fine_fast_kernel<4, kFastWarps><<<ceil_div(n_rays, 2 * kFastWarps), kFastWarps * kWarp, 0, st>>>(
                z_coarse, weights_coarse, u, u_stride, rng, use_rng, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg, n_rays);
        else
            sample_fine_fast_kernel<8, 4><<<ceil_div(n_rays, 2 * 4), 4 * kWarp, 0, st>>>(
                z_coarse, weights_coarse, u, u_stride, rng, use_rng, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg, n_rays);
        SNERF_LAUNCH_OK("sample_fine_fast_kernel");
        return SNERF_OK;
    }
    int npad = 2;
    while (npad < s_coarse + n_new) npad <<= 1;
    const size_t smem = (size_t)kFineWarps * (2 * (s_coarse - 1) + npad) * sizeof(float);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        SNERF_CUDA_OK(cudaFuncSetAttribute(sample_fine_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    sample_fine_generic_kernel<<<ceil_div(n_rays, kFineWarps), kFineWarps * kWarp, smem, (cudaStream_t)stream>>>(
        z_coarse, weights_coarse, u, u_stride, rng, use_rng, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg, n_rays, s_coarse,
        n_new, npad);
    SNERF_LAUNCH_OK("sample_fine_generic_kernel");
    return SNERF_OK;
}

extern "C" int snerf_sample_fine(const float* z_coarse, const float* weights_coarse, const float* u, int u_stride,
                                 float* z_fine, float* samples_dbg, float* cdf_dbg, int32_t* below_dbg,
                                 int32_t* above_dbg, int n_rays, int s_coarse, int n_new, void* stream) {
    return sample_fine_impl(z_coarse, weights_coarse, u, u_stride, RngKey{0, 0}, false, z_fine, samples_dbg, cdf_dbg, below_dbg, above_dbg,
                            n_rays, s_coarse, n_new, stream, "snerf_sample_fine");
}

extern "C" int snerf_sample_fine_rng(const float* z_coarse, const float* weights_coarse, uint64_t seed, uint64_t offset,
                                     float* z_fine, int n_rays, int s_coarse, int n_new, void* stream) {
    return sample_fine_impl(z_coarse, weights_coarse, nullptr, 0, RngKey{seed, offset}, true, z_fine, nullptr, nullptr, nullptr, nullptr,
                            n_rays, s_coarse, n_new, stream, "snerf_sample_fine_rng");
}

// out[e] = scale * (uniform [0,1) or standard normal) of element e of draw (seed, offset): the numbers the *_rng entry points consume
__global__ void __launch_bounds__(256) fill_random_kernel(float* __restrict__ out, long long n, int normal, float scale, const RngKey rng) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * b >= n) return;
    const float4 v = normal ? rng_normal4(rng, (unsigned long long)b) : rng_uniform4(rng, (unsigned long long)b);
    const float r[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    for (int i = 0; i < 4 && 4 * b + i < n; ++i) out[4 * b + i] = r[i];
}

extern "C" int snerf_fill_random(float* out, long long n, int normal, float scale, uint64_t seed, uint64_t offset, void* stream) {
    SNERF_REQUIRE(n >= 0, "snerf_fill_random: bad size");
    if (n == 0) return SNERF_OK;
    SNERF_REQUIRE(out != nullptr, "snerf_fill_random: null pointer");
    const long long blocks = ((n + 3) / 4 + 255) / 256;
    fill_random_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, normal, scale, RngKey{seed, offset});
    SNERF_LAUNCH_OK("fill_random_kernel");
    return SNERF_OK;
}
