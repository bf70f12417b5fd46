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
// i.e. 27 x 128 + 128 multiply-adds in fp32.
// Forward (tc_vis_fwd_kernel): points per thread -- the encodings of two points x two views in registers, the direction
// columns of W_view in shared memory (every lane reads the same row: LDS.128 broadcasts, one row serves four FMAs per lane),
// the dot product with the fourth row accumulated in the thread: no exchange between threads at all.
// Backward (tc_vis_bwd_kernel, before snerf_mlp_backward): a thread owns two columns of the view layer, 32 points per block
// round, the direction encodings of the round in shared memory -- the layout in which the weight gradients (sums over points
// per column) stay in registers.  With dv_v = d visibility_v * sigmoid' and dY_v2,v = dv_v w_vis * [hv_v > 0] the kernel
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

// ---- forward: one thread per point --------------------------------------------------------------------------------
// direction from another view's camera centre to the sample point, encoded (:317-325, :646); e[27] = 1 (bias column)
__device__ __forceinline__ void vis_encode_dir(const VisParams& p, long long pt, int ray, int v, float (&e)[kVisEnc]) {
    const float ox = p.rays_o[ray * 3], oy = p.rays_o[ray * 3 + 1], oz = p.rays_o[ray * 3 + 2];
    const float dx = p.rays_d[ray * 3], dy = p.rays_d[ray * 3 + 1], dz = p.rays_d[ray * 3 + 2];
    float zz = p.z[pt];
    if (p.ndc) {                                                                             // :319-321 (near = 1)
        const float tn = -(1.f + oz) / dz;
        zz = (((oz + tn * dz) / (1.f - zz + 1e-6f)) - oz) / dz;
    }
    const float* o2 = p.rays_o2 + ((size_t)ray * p.n_other + v) * 3;
    float d[3] = {__fadd_rn(ox, __fmul_rn(zz, dx)) - o2[0], __fadd_rn(oy, __fmul_rn(zz, dy)) - o2[1],
                  __fadd_rn(oz, __fmul_rn(zz, dz)) - o2[2]};                               // :322-323
    const float norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                    // :324
    d[0] /= norm; d[1] /= norm; d[2] /= norm;
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
    e[kVisEnc - 1] = 1.f;
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Shared memory is the pacing resource of both kernels: an LDS.128 whose lanes all read the same 16 bytes still costs four
// wavefronts, i.e. one shared-memory cycle per broadcast float, so every broadcast value has to feed several FMAs per lane.
// Forward: a thread works on TWO points and two views at a time (one weight -> four FMAs per lane).
template <int NP>   // points per thread
__device__ __forceinline__ void vis_other_views(const VisParams& p, const float (*s_w)[kVisEnc], const float* s_wv, const long long (&pt)[NP],
                                                const int (&ray)[NP], const bool (&live)[NP], int v0, bool two, float b_vis) {
    float e[NP][2][kVisEnc];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        vis_encode_dir(p, pt[i], ray[i], v0, e[i][0]);
        vis_encode_dir(p, pt[i], ray[i], two ? v0 + 1 : v0, e[i][1]);
    }
    float acc[NP][2];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i][0] = acc[i][1] = 0.f;
    uint4 nx[NP];                      // the next 16 bytes of each row, requested one iteration ahead
#pragma unroll
    for (int i = 0; i < NP; ++i) nx[i] = __ldg(reinterpret_cast<const uint4*>(p.pre + pt[i] * 128));
    for (int ch = 0; ch < 16; ++ch) {
        uint32_t qq[NP][4];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            qq[i][0] = nx[i].x; qq[i][1] = nx[i].y; qq[i][2] = nx[i].z; qq[i][3] = nx[i].w;
            if (ch + 1 < 16) nx[i] = __ldg(reinterpret_cast<const uint4*>(p.pre + pt[i] * 128) + ch + 1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int o = 8 * ch + k;
            float h[NP][2], g[NP][2];       // two short chains per point and view instead of one of 28 dependent FMAs
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const float pre = (k & 1) ? bf16_hi(qq[i][k >> 1]) : bf16_lo(qq[i][k >> 1]);
                h[i][0] = h[i][1] = pre;
                g[i][0] = g[i][1] = 0.f;
            }
            const float4* wr = reinterpret_cast<const float4*>(&s_w[o][0]);
#pragma unroll
            for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                const float4 ww = wr[c4];
#pragma unroll
                for (int i = 0; i < NP; ++i)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        h[i][v] = fmaf(ww.x, e[i][v][4 * c4], fmaf(ww.y, e[i][v][4 * c4 + 1], h[i][v]));
                        g[i][v] = fmaf(ww.z, e[i][v][4 * c4 + 2], fmaf(ww.w, e[i][v][4 * c4 + 3], g[i][v]));
                    }
            }
            const float wv = s_wv[o];
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int v = 0; v < 2; ++v) acc[i][v] = fmaf(wv, fmaxf(h[i][v] + g[i][v], 0.f), acc[i][v]);
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (!live[i]) continue;
        p.vis2[pt[i] * p.n_other + v0] = vis_sigmoid(acc[i][0] + b_vis);
        if (two) p.vis2[pt[i] * p.n_other + v0 + 1] = vis_sigmoid(acc[i][1] + b_vis);
    }
}

constexpr int kVfPts = 2;           // points per thread in the forward kernel
__global__ void __launch_bounds__(kVisThreads) tc_vis_fwd_kernel(const __grid_constant__ VisParams p) {
    __shared__ __align__(16) float s_w[kVisThreads][kVisEnc];     // row o: the 27 direction columns of W_view, then the bias all views share
    __shared__ __align__(16) float s_wv[kVisThreads];             // fourth row of the view head
    const int t = threadIdx.x;
    for (int i = t; i < kVisThreads * 27; i += kVisThreads) {
        const int o = i / 27, c = i - 27 * o;
        s_w[o][c] = p.w_view[(size_t)o * p.view_in + p.dir_col0 + c];
    }
    {
        float bc = p.b_view[t];
        for (int j = 0; j < 256; ++j) bc = fmaf(p.w_view[(size_t)t * p.view_in + j], __ldg(p.b_feat + j), bc);   // feature bias through the view layer
        s_w[t][27] = bc;
        s_wv[t] = p.w_vis[t];
    }
    __syncthreads();
    const float b_vis = p.b_vis[0];
    const int nv = p.n_other;
    constexpr int kBlockPts = kVfPts * kVisThreads;
    for (long long base = (long long)blockIdx.x * kBlockPts; base < p.n_points; base += (long long)gridDim.x * kBlockPts) {
        long long pt[kVfPts];
        int ray[kVfPts];
        bool live[kVfPts];
#pragma unroll
        for (int i = 0; i < kVfPts; ++i) {
            const long long q = base + i * kVisThreads + t;
            live[i] = q < p.n_points;
            pt[i] = live[i] ? q : p.n_points - 1;          // a thread past the end repeats the last point and stores nothing
            ray[i] = (int)((unsigned)pt[i] / (unsigned)p.n_samples);       // n_points < 2^31 (host check)
        }
        // own view: the per-ray bias already holds its direction part
#pragma unroll
        for (int i = 0; i < kVfPts; ++i) {
            const uint4* prow = reinterpret_cast<const uint4*>(p.pre + pt[i] * 128);
            const float4* vb = reinterpret_cast<const float4*>(p.view_bias + (size_t)ray[i] * 128);
            float a = 0.f;
#pragma unroll 2
            for (int ch = 0; ch < 16; ++ch) {
                const uint4 q = __ldg(prow + ch);
                const float4 b0 = __ldg(vb + 2 * ch), b1 = __ldg(vb + 2 * ch + 1);
                const float4 w0 = *reinterpret_cast<const float4*>(s_wv + 8 * ch), w1 = *reinterpret_cast<const float4*>(s_wv + 8 * ch + 4);
                a = fmaf(w0.x, fmaxf(bf16_lo(q.x) + b0.x, 0.f), a); a = fmaf(w0.y, fmaxf(bf16_hi(q.x) + b0.y, 0.f), a);
                a = fmaf(w0.z, fmaxf(bf16_lo(q.y) + b0.z, 0.f), a); a = fmaf(w0.w, fmaxf(bf16_hi(q.y) + b0.w, 0.f), a);
                a = fmaf(w1.x, fmaxf(bf16_lo(q.z) + b1.x, 0.f), a); a = fmaf(w1.y, fmaxf(bf16_hi(q.z) + b1.y, 0.f), a);
                a = fmaf(w1.z, fmaxf(bf16_lo(q.w) + b1.z, 0.f), a); a = fmaf(w1.w, fmaxf(bf16_hi(q.w) + b1.w, 0.f), a);
            }
            if (live[i]) p.vis[pt[i]] = vis_sigmoid(a + b_vis);                                  // :710-713
        }
        // other views, two per pass over the rows
        for (int v0 = 0; v0 < nv; v0 += 2) vis_other_views<kVfPts>(p, s_w, s_wv, pt, ray, live, v0, v0 + 1 < nv, b_vis);
    }
}

// ---- backward: a thread owns TWO columns of the view layer (one broadcast encoding value -> four FMAs per lane) ---------
constexpr int kVbThreads = kVisThreads / 2;
__global__ void __launch_bounds__(kVbThreads) tc_vis_bwd_kernel(const __grid_constant__ VisParams p) {
    extern __shared__ float sm[];
    const int nv = p.n_other, nk = 1 + nv;
    float* s_pe = sm;                                            // [nv][kVisPts][kVisEnc]  PE(dir_v)
    float* s_own = s_pe + nv * kVisPts * kVisEnc;                // [kVisPts][kVisEnc]      PE(own dir) of the point's ray
    float* s_dv = s_own + kVisPts * kVisEnc;                     // [nk][kVisPts]  d visibility * sigmoid'  (k = 0: own view)
    __shared__ int s_ray[kVisPts];
    const int t = threadIdx.x, lane = t & 31;

    // per column (t and t + 64): its row of the direction columns, the bias shared by all views, its element of the fourth row
    float w[2][kVisEnc], bc[2], wv[2], gw[2] = {0.f, 0.f}, gdir[2][kVisEnc];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int o = t + q * kVbThreads;
#pragma unroll
        for (int c = 0; c < kVisEnc; ++c) {
            w[q][c] = c < 27 ? p.w_view[(size_t)o * p.view_in + p.dir_col0 + c] : 0.f;
            gdir[q][c] = 0.f;
        }
        float b = p.b_view[o];
        for (int j = 0; j < 256; ++j) b = fmaf(p.w_view[(size_t)o * p.view_in + j], __ldg(p.b_feat + j), b);   // feature bias through the view layer
        bc[q] = b;
        wv[q] = p.w_vis[o];
    }
    float gb = 0.f;

    const long long n_chunks = (p.n_points + kVisPts - 1) / kVisPts;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long p0 = chunk * kVisPts;
        // ---- phase A: the direction encodings of the round (other views, own view); head gradients before the sigmoid ----
        for (int idx = t; idx < nv * kVisPts; idx += kVbThreads) {
            const int v = idx / kVisPts, pi = idx % kVisPts;
            const long long pt = p0 + pi;
            float e[kVisEnc];
#pragma unroll
            for (int c = 0; c < kVisEnc; ++c) e[c] = 0.f;
            if (pt < p.n_points) {
                vis_encode_dir(p, pt, (int)((unsigned)pt / (unsigned)p.n_samples), v, e);
                e[kVisEnc - 1] = 0.f;                          // (the bias is added separately here)
            }
#pragma unroll
            for (int c = 0; c < kVisEnc; ++c) s_pe[(size_t)idx * kVisEnc + c] = e[c];
        }
        if (t < kVisPts) {
            const long long pt = p0 + t;
            const int ray = pt < p.n_points ? (int)((unsigned)pt / (unsigned)p.n_samples) : 0;       // n_points < 2^31 (host check)
            s_ray[t] = ray;
            for (int c = 0; c < kVisEnc; ++c) s_own[t * kVisEnc + c] = (c < 27 && pt < p.n_points) ? p.view_enc[(size_t)ray * 32 + c] : 0.f;
        }
        for (int idx = t; idx < nk * kVisPts; idx += kVbThreads) {
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
            s_dv[idx] = dv;
            gb += dv;
        }
        __syncthreads();
        // ---- phase B: this thread's two columns of the view layer for every point and view of the round ----
        const int n_here = (int)((p.n_points - p0) < kVisPts ? (p.n_points - p0) : kVisPts);
        // the next point's two global operands are requested one iteration ahead (the stores of `extra` would otherwise pin
        // every load behind them: ~600 cycles of exposed latency per point at three warps per scheduler)
        uint16_t pre_nx[2];
        float vb_nx[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            pre_nx[q] = __ldg(p.pre + p0 * 128 + t + q * kVbThreads);
            vb_nx[q] = __ldg(p.view_bias + (size_t)s_ray[0] * 128 + t + q * kVbThreads);
        }
        for (int pi = 0; pi < n_here; ++pi) {
            const long long pt = p0 + pi;
            const float dv0 = s_dv[pi];
            float pre[2], vbias[2], ex[2], dysum[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) { pre[q] = __uint_as_float((uint32_t)pre_nx[q] << 16); vbias[q] = vb_nx[q]; }
            if (pi + 1 < n_here) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    pre_nx[q] = __ldg(p.pre + (pt + 1) * 128 + t + q * kVbThreads);
                    vb_nx[q] = __ldg(p.view_bias + (size_t)s_ray[pi + 1] * 128 + t + q * kVbThreads);
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float h = pre[q] + vbias[q];
                ex[q] = h > 0.f ? dv0 * wv[q] : 0.f;
                gw[q] = fmaf(dv0, fmaxf(h, 0.f), gw[q]);
                dysum[q] = 0.f;
            }
            for (int v = 0; v < nv; ++v) {
                float e[kVisEnc];
                const float4* pe = reinterpret_cast<const float4*>(s_pe + (size_t)(v * kVisPts + pi) * kVisEnc);
#pragma unroll
                for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                    const float4 x = pe[c4];
                    e[4 * c4] = x.x; e[4 * c4 + 1] = x.y; e[4 * c4 + 2] = x.z; e[4 * c4 + 3] = x.w;
                }
                const float dv = s_dv[(1 + v) * kVisPts + pi];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float s0 = pre[q] + bc[q], s1 = 0.f, s2 = 0.f, s3 = 0.f;      // four short chains instead of one of 28 dependent FMAs
#pragma unroll
                    for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                        s0 = fmaf(w[q][4 * c4], e[4 * c4], s0); s1 = fmaf(w[q][4 * c4 + 1], e[4 * c4 + 1], s1);
                        s2 = fmaf(w[q][4 * c4 + 2], e[4 * c4 + 2], s2); s3 = fmaf(w[q][4 * c4 + 3], e[4 * c4 + 3], s3);
                    }
                    const float h2 = (s0 + s1) + (s2 + s3);
                    const float dyv = h2 > 0.f ? dv * wv[q] : 0.f;
                    ex[q] += dyv;
                    dysum[q] += dyv;
                    gw[q] = fmaf(dv, fmaxf(h2, 0.f), gw[q]);
#pragma unroll
                    for (int c = 0; c < kVisEnc; ++c) gdir[q][c] = fmaf(dyv, e[c], gdir[q][c]);
                }
            }
            if (nv > 0) {      // the chain forms these columns with the own direction for all views: take that share back out
                const float4* po = reinterpret_cast<const float4*>(s_own + (size_t)pi * kVisEnc);
#pragma unroll
                for (int c4 = 0; c4 < kVisEnc / 4; ++c4) {
                    const float4 x = po[c4];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        gdir[q][4 * c4] = fmaf(-dysum[q], x.x, gdir[q][4 * c4]);
                        gdir[q][4 * c4 + 1] = fmaf(-dysum[q], x.y, gdir[q][4 * c4 + 1]);
                        gdir[q][4 * c4 + 2] = fmaf(-dysum[q], x.z, gdir[q][4 * c4 + 2]);
                        gdir[q][4 * c4 + 3] = fmaf(-dysum[q], x.w, gdir[q][4 * c4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) p.extra[pt * 128 + t + q * kVbThreads] = ex[q];
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int o = t + q * kVbThreads;
        atomicAdd(p.g_w_vis + o, gw[q]);
#pragma unroll
        for (int c = 0; c < 27; ++c) atomicAdd(p.g_w_view + (size_t)o * p.view_in + p.dir_col0 + c, gdir[q][c]);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, s);
    if (lane == 0) atomicAdd(p.g_b_vis, gb);
}

static size_t vis_bwd_smem_bytes(int n_other) {
    return ((size_t)(n_other + 1) * kVisPts * kVisEnc + (size_t)(1 + n_other) * kVisPts) * sizeof(float) + 16;
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
    if (!done) {   // eight other views: 33 KB of encodings per round in the backward kernel
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_vis_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vis_bwd_smem_bytes(kVisMaxOther)));
        done = true;
    }
    return SNERF_OK;
}

// persistent grids: per_sm blocks per SM (the backward kernel's registers allow four), fewer for small inputs
static int vis_grid(long long n_points, int pts_per_block, int per_sm) {
    const long long chunks = (n_points + pts_per_block - 1) / pts_per_block;
    const long long cap = (long long)num_sms() * per_sm;
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
    tc_vis_fwd_kernel<<<vis_grid(p.n_points, kVfPts * kVisThreads, 3), kVisThreads, 0, st>>>(p);
    SNERF_LAUNCH_OK("tc_vis_fwd_kernel");
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
    tc_vis_bwd_kernel<<<vis_grid(p.n_points, kVisPts, 6), kVbThreads, vis_bwd_smem_bytes(n_other), st>>>(p);
    SNERF_LAUNCH_OK("tc_vis_bwd_kernel");
    return SNERF_OK;
}

}  // namespace snerf
