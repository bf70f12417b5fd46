// Row a14 / N4 on the tensor path: the secondary-view visibility head (predict_visibility=True).
//
// Reference: src/models/SimpleNeRF01.py  MLP ctor :596-608 (views_output_linear has a fourth row), MLP.forward :640-649 (the
// view branch once more per other view), get_view_dependent_outputs :687-715, compute_other_view_dirs :317-325.
//
// Split of the work.  The view layer's input is [feature | enc_hi | PE(direction)] (:695), and only the last 27 columns
// differ between the ray's own view and the other views.  The shared part -- a 128 x 256 (+42) product per point -- is
// the view step of tc_forward_kernel (tcgen05); with SNERF_FLAG_VIS_HEAD its epilogue also writes that accumulator,
// without the per-ray bias, as bf16 [P,128] (`pre`).  What is left per point and view is
//     hv_v = relu(pre + b + W_view[:, dir columns] PE(dir_v)),   visibility_v = sigmoid(w_vis . hv_v + b_vis)
// i.e. 27 x 128 + 128 multiply-adds in fp32: tc_vis_kernel, thread o = column o of the view layer, 32 points per block
// round, the per-view direction encodings of the round in shared memory (read as broadcasts).
// Backward (tc_vis_kernel<true>, before snerf_mlp_backward): with dv_v = d visibility_v * sigmoid' and dY_v2,v = dv_v w_vis
// * [hv_v > 0] the kernel
//   * adds the fourth row's gradients (w_vis, b_vis) to `grads`,
//   * writes  extra = [hv_own > 0] dv_own w_vis + sum_v dY_v2,v  (fp32 [P,128]); the dgrad prologue adds it to dY_v, so the
//     chain, the view layer's weight-gradient job (G = dY_v^T [h8 | enc_hi | PE(own dir)], column sums) and the unmerge
//     step see the SUM over all views -- exact for every column the views share;
//   * adds the correction for the direction columns, which that job forms with the own direction for all views:
//     dW_view[:, dir] += sum_v dY_v2,v^T (PE(dir_v) - PE(dir_own)).
#include "common.cuh"
#include "tc_plan.cuh"

namespace snerf {

constexpr int kVisPts = 32;         // points per block round
constexpr int kVisThreads = 128;    // = view_width
constexpr int kVisEnc = 28;         // 27 encoding columns + one zero (float4 reads)
constexpr int kVisMaxOther = 8;

struct VisParams {
    const uint16_t* pre;            // bf16 [P,128]
    const float* view_bias;         // [n_rays,128]: b + W_view[:, dir] PE(own dir)  (tc_view_bias_kernel)
    const float* view_enc;          // [n_rays,32]: PE(own dir)
    const float *w_view, *b_view, *b_feat, *w_vis, *b_vis;
    const float *rays_o, *rays_d, *z, *rays_o2;
    float *vis, *vis2;              // forward: outputs; backward: the forward's outputs
    const float *d_vis, *d_vis2;    // backward, nullable
    float* extra;                   // backward: [P,128]
    float *g_w_view, *g_w_vis, *g_b_vis;
    long long n_points;
    int n_samples, n_other, ndc, view_in, dir_col0, view_degree;
};

__device__ __forceinline__ float vis_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

template <bool kBwd>
__global__ void __launch_bounds__(kVisThreads) tc_vis_kernel(const __grid_constant__ VisParams p) {
    extern __shared__ float sm[];
    const int nv = p.n_other, nk = 1 + nv;
    float* s_pe = sm;                                            // [nv][kVisPts][kVisEnc]
    float* s_diff = s_pe + (kBwd ? nv * kVisPts * kVisEnc : 0);  // (backward) PE(dir_v) - PE(own dir)
    float* s_val = s_diff + nv * kVisPts * kVisEnc;              // forward: [4 warps][nk][kVisPts] partial logits; backward: [nk][kVisPts] dv
    const int o = threadIdx.x, warp = o >> 5, lane = o & 31;

    // this thread's row of the direction columns, the bias shared by all views, its element of the fourth row
    float w[kVisEnc];
#pragma unroll
    for (int c = 0; c < kVisEnc; ++c) w[c] = c < 27 ? p.w_view[(size_t)o * p.view_in + p.dir_col0 + c] : 0.f;
    float bc = p.b_view[o];
    for (int j = 0; j < 256; ++j) bc = fmaf(p.w_view[(size_t)o * p.view_in + j], __ldg(p.b_feat + j), bc);   // feature bias through the view layer
    const float wv = p.w_vis[o];
    const float b_vis = p.b_vis[0];
    float gw = 0.f, gb = 0.f, gdir[kVisEnc];
#pragma unroll
    for (int c = 0; c < kVisEnc; ++c) gdir[c] = 0.f;

    const long long n_chunks = (p.n_points + kVisPts - 1) / kVisPts;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long p0 = chunk * kVisPts;
        // ---- phase A: directions from the other views' camera centres to the points, encoded (:317-325, :646) ----
        for (int idx = o; idx < nv * kVisPts; idx += kVisThreads) {
            const int v = idx / kVisPts, pi = idx % kVisPts;
            const long long pt = p0 + pi;
            float* pe = s_pe + (size_t)idx * kVisEnc;
            if (pt >= p.n_points) {
#pragma unroll
                for (int c = 0; c < kVisEnc; ++c) pe[c] = 0.f;
                if (kBwd) {
#pragma unroll
                    for (int c = 0; c < kVisEnc; ++c) s_diff[(size_t)idx * kVisEnc + c] = 0.f;
                }
                continue;
            }
            const int ray = (int)(pt / p.n_samples);
            const float ox = p.rays_o[ray * 3], oy = p.rays_o[ray * 3 + 1], oz = p.rays_o[ray * 3 + 2];
            const float dx = p.rays_d[ray * 3], dy = p.rays_d[ray * 3 + 1], dz = p.rays_d[ray * 3 + 2];
            float zz = p.z[pt];
            if (p.ndc) {                                                                             // :319-321 (near = 1)
                const float tn = -(1.f + oz) / dz;
                zz = (((oz + tn * dz) / (1.f - zz + 1e-6f)) - oz) / dz;
            }
            const float* o2 = p.rays_o2 + ((size_t)ray * nv + v) * 3;
            float d[3] = {__fadd_rn(ox, __fmul_rn(zz, dx)) - o2[0], __fadd_rn(oy, __fmul_rn(zz, dy)) - o2[1],
                          __fadd_rn(oz, __fmul_rn(zz, dz)) - o2[2]};                               // :322-323
            const float norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                    // :324
            d[0] /= norm; d[1] /= norm; d[2] /= norm;
            float e[kVisEnc];
#pragma unroll
            for (int c = 0; c < kVisEnc; ++c) e[c] = 0.f;
            e[0] = d[0]; e[1] = d[1]; e[2] = d[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < p.view_degree) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float sn, cs;
                        sincosf(d[c] * (float)(1 << k), &sn, &cs);
                        e[3 + 6 * k + c] = sn;
                        e[6 + 6 * k + c] = cs;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kVisEnc; ++c) pe[c] = e[c];
            if (kBwd) {
#pragma unroll
                for (int c = 0; c < kVisEnc; ++c) s_diff[(size_t)idx * kVisEnc + c] = c < 27 ? e[c] - p.view_enc[(size_t)ray * 32 + c] : 0.f;
            }
        }
        if (kBwd) {
            // dv[k][pi] = d visibility * sigmoid'  (k = 0: own view)
            for (int idx = o; idx < nk * kVisPts; idx += kVisThreads) {
                const int k = idx / kVisPts, pi = idx % kVisPts;
                const long long pt = p0 + pi;
                float dv = 0.f;
                if (pt < p.n_points) {
                    if (k == 0) {
                        const float a = p.vis[pt];
                        dv = p.d_vis ? p.d_vis[pt] * a * (1.f - a) : 0.f;
                    } else {
                        const float a = p.vis2[pt * nv + (k - 1)];
                        dv = p.d_vis2 ? p.d_vis2[pt * nv + (k - 1)] * a * (1.f - a) : 0.f;
                    }
                }
                s_val[idx] = dv;
                gb += dv;
            }
        }
        __syncthreads();
        // ---- phase B: column o of the view layer for every point and view of the round ----
        const int n_here = (int)((p.n_points - p0) < kVisPts ? (p.n_points - p0) : kVisPts);
        for (int pi = 0; pi < n_here; ++pi) {
            const long long pt = p0 + pi;
            const int ray = (int)(pt / p.n_samples);
            const float pre = __uint_as_float((uint32_t)p.pre[pt * 128 + o] << 16);
            const float h = pre + p.view_bias[(size_t)ray * 128 + o];
            float ex = 0.f;
            if (kBwd) {
                const float dv = s_val[pi];
                ex = h > 0.f ? dv * wv : 0.f;
                gw = fmaf(dv, fmaxf(h, 0.f), gw);
            } else {
                float part = wv * fmaxf(h, 0.f);
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
                if (lane == 0) s_val[(warp * nk) * kVisPts + pi] = part;
            }
            for (int v = 0; v < nv; ++v) {
                const float4* pe = reinterpret_cast<const float4*>(s_pe + (size_t)(v * kVisPts + pi) * kVisEnc);
                float h2 = pre + bc;
#pragma unroll
                for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                    const float4 e = pe[c4];
                    h2 = fmaf(w[4 * c4], e.x, fmaf(w[4 * c4 + 1], e.y, fmaf(w[4 * c4 + 2], e.z, fmaf(w[4 * c4 + 3], e.w, h2))));
                }
                if (kBwd) {
                    const float dv = s_val[(1 + v) * kVisPts + pi];
                    const float dyv = h2 > 0.f ? dv * wv : 0.f;
                    ex += dyv;
                    gw = fmaf(dv, fmaxf(h2, 0.f), gw);
                    const float4* df = reinterpret_cast<const float4*>(s_diff + (size_t)(v * kVisPts + pi) * kVisEnc);
#pragma unroll
                    for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                        const float4 e = df[c4];
                        gdir[4 * c4] = fmaf(dyv, e.x, gdir[4 * c4]);
                        gdir[4 * c4 + 1] = fmaf(dyv, e.y, gdir[4 * c4 + 1]);
                        gdir[4 * c4 + 2] = fmaf(dyv, e.z, gdir[4 * c4 + 2]);
                        gdir[4 * c4 + 3] = fmaf(dyv, e.w, gdir[4 * c4 + 3]);
                    }
                } else {
                    float part = wv * fmaxf(h2, 0.f);
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
                    if (lane == 0) s_val[(warp * nk + 1 + v) * kVisPts + pi] = part;
                }
            }
            if (kBwd) p.extra[pt * 128 + o] = ex;
        }
        __syncthreads();
        if (!kBwd) {
            // ---- phase C: the four warps' partial sums in a fixed order, bias, sigmoid (:710-713) ----
            for (int idx = o; idx < nk * kVisPts; idx += kVisThreads) {
                const int k = idx / kVisPts, pi = idx % kVisPts;
                const long long pt = p0 + pi;
                if (pt >= p.n_points) continue;
                float a = b_vis;
#pragma unroll
                for (int wq = 0; wq < kVisThreads / 32; ++wq) a += s_val[(wq * nk + k) * kVisPts + pi];
                a = vis_sigmoid(a);
                if (k == 0) p.vis[pt] = a;
                else p.vis2[pt * nv + (k - 1)] = a;
            }
            __syncthreads();
        }
    }
    if (kBwd) {
        atomicAdd(p.g_w_vis + o, gw);
#pragma unroll
        for (int c = 0; c < 27; ++c) atomicAdd(p.g_w_view + (size_t)o * p.view_in + p.dir_col0 + c, gdir[c]);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, s);
        if (lane == 0) atomicAdd(p.g_b_vis, gb);
    }
}

static size_t vis_smem_bytes(bool bwd, int n_other) {
    const size_t pe = (size_t)n_other * kVisPts * kVisEnc;
    const size_t val = (size_t)(bwd ? 1 : kVisThreads / 32) * (1 + n_other) * kVisPts;
    return ((bwd ? 2 : 1) * pe + val) * sizeof(float) + 16;
}

static int vis_params(const snerf_mlp_desc& d, const float* const* prm, const void* mlp_ws, int n_rays, int n_samples, int n_other,
                      uint32_t flags, VisParams& p, TcWorkspace& w) {
    const MlpDims m(d);
    SNERF_REQUIRE(m.has_view && m.venc == 27 && m.view_width == kVisThreads, "visibility head: needs the view branch with a degree-4 direction encoding");
    SNERF_REQUIRE(flags & SNERF_FLAG_VIS_HEAD, "visibility head (tensor path): snerf_mlp_forward must have run with SNERF_FLAG_VIS_HEAD (and these calls get the same flags)");
    SNERF_REQUIRE(n_other <= kVisMaxOther, "visibility head (tensor path): at most %d other views, got %d", kVisMaxOther, n_other);
    const TcPlan pl = build_plan(d, nullptr);
    w = tc_ws_layout(m, pl, n_rays, n_samples, flags);
    const uint8_t* wsb = (const uint8_t*)(((uintptr_t)mlp_ws + 1023) & ~(uintptr_t)1023);
    p.pre = (const uint16_t*)(wsb + w.vis_pre);
    p.view_bias = (const float*)(wsb + w.view_bias);
    p.view_enc = (const float*)(wsb + w.view_enc);
    p.w_view = prm[SNERF_P_VIEW_W]; p.b_view = prm[SNERF_P_VIEW_B]; p.b_feat = prm[SNERF_P_FEAT_B];
    p.w_vis = prm[SNERF_P_RGB_W] + 3 * m.view_width;                 // fourth row of views_output_linear (:710-711)
    p.b_vis = prm[SNERF_P_RGB_B] + 3;
    p.n_points = (long long)n_rays * n_samples;
    p.n_samples = n_samples; p.n_other = n_other; p.ndc = (flags & SNERF_FLAG_NDC) ? 1 : 0;
    p.view_in = m.view_in; p.dir_col0 = m.width + m.enc_hi; p.view_degree = d.view_degree;
    return SNERF_OK;
}

static int vis_smem_attr() {
    static bool done = false;
    if (!done) {   // eight other views need 57 KB in the backward kernel
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_vis_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vis_smem_bytes(false, kVisMaxOther)));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_vis_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vis_smem_bytes(true, kVisMaxOther)));
        done = true;
    }
    return SNERF_OK;
}

static int vis_grid(long long n_points) {
    const long long chunks = (n_points + kVisPts - 1) / kVisPts;
    const long long cap = (long long)num_sms() * 8;
    return (int)(chunks < cap ? chunks : cap);
}

int tc_visibility_forward(const snerf_mlp_desc& d, const float* const* prm, const void* mlp_ws, const float* rays_o,
                          const float* rays_d, const float* z, const float* rays_o2, float* visibility, float* visibility2,
                          int n_rays, int n_samples, int n_other, uint32_t flags, cudaStream_t st) {
    VisParams p{};
    TcWorkspace w{};
    const int rc = vis_params(d, prm, mlp_ws, n_rays, n_samples, n_other, flags, p, w);
    if (rc != SNERF_OK) return rc;
    if (p.n_points == 0) return SNERF_OK;
    p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.rays_o2 = rays_o2;
    p.vis = visibility; p.vis2 = visibility2;
    if (vis_smem_attr() != SNERF_OK) return SNERF_ERR_CUDA;
    tc_vis_kernel<false><<<vis_grid(p.n_points), kVisThreads, vis_smem_bytes(false, n_other), st>>>(p);
    SNERF_LAUNCH_OK("tc_vis_kernel<forward>");
    return SNERF_OK;
}

int tc_visibility_backward(const snerf_mlp_desc& d, const float* const* prm, void* mlp_ws, const float* rays_o,
                           const float* rays_d, const float* z, const float* rays_o2, const float* visibility,
                           const float* visibility2, const float* d_visibility, const float* d_visibility2,
                           float* const* grads, int n_rays, int n_samples, int n_other, uint32_t flags, cudaStream_t st) {
    VisParams p{};
    TcWorkspace w{};
    const int rc = vis_params(d, prm, mlp_ws, n_rays, n_samples, n_other, flags, p, w);
    if (rc != SNERF_OK) return rc;
    if (p.n_points == 0) return SNERF_OK;
    SNERF_REQUIRE(flags & SNERF_FLAG_SAVE_FOR_BWD, "visibility_backward: the forward must have run with SNERF_FLAG_SAVE_FOR_BWD");
    uint8_t* wsb = (uint8_t*)(((uintptr_t)mlp_ws + 1023) & ~(uintptr_t)1023);
    p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.rays_o2 = rays_o2;
    p.vis = const_cast<float*>(visibility); p.vis2 = const_cast<float*>(visibility2);
    p.d_vis = d_visibility; p.d_vis2 = d_visibility2;
    p.extra = (float*)(wsb + w.vis_extra);
    p.g_w_view = grads[SNERF_P_VIEW_W];
    p.g_w_vis = grads[SNERF_P_RGB_W] + 3 * kVisThreads;
    p.g_b_vis = grads[SNERF_P_RGB_B] + 3;
    if (vis_smem_attr() != SNERF_OK) return SNERF_ERR_CUDA;
    tc_vis_kernel<true><<<vis_grid(p.n_points), kVisThreads, vis_smem_bytes(true, n_other), st>>>(p);
    SNERF_LAUNCH_OK("tc_vis_kernel<backward>");
    return SNERF_OK;
}

}  // namespace snerf
