// fp32 CUDA-core MLP ("precise" path, SNERF_FLAG_PRECISE): positional encoding, the 8x256 trunk
// with its skip concat, sigma / feature / view / rgb layers, forward and backward, as a chain of
// tiled fp32 GEMM launches.  It exists so that every non-GEMM piece of the hot path (encoding order,
// concat order, noise, heads, compositing gradients) can be checked against the fp32 reference to
// ~1e-5, and as the numerically strict mode of the drop-in; the throughput path is mlp_tc.cu.
//
// Reference behaviour: src/models/SimpleNeRF01.py  PositionalEncoder :525-557, MLP.forward :626-654,
// get_view_independent_outputs :656-685, get_view_dependent_outputs :687-715.
#include "common.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------------------
// workspace layout (floats)
// ------------------------------------------------------------------------------------------------
struct SimtLayout {
    size_t enc, venc, x5, h[8], head, xv, hv, rgbraw, dya, dyb, dhv, dhead, total;
    int x5_ld, xv_ld;
};

static SimtLayout simt_layout(const MlpDims& m, int n_rays, int n_samples) {
    SimtLayout L{};
    const size_t P = (size_t)n_rays * n_samples;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n, 64); return o; };
    L.x5_ld = m.trunk_in + m.width;
    L.xv_ld = m.view_in;
    L.enc = take(P * m.enc);
    L.venc = take((size_t)n_rays * (m.venc > 0 ? m.venc : 1));
    L.x5 = take(P * L.x5_ld);
    for (int l = 0; l < m.depth; ++l) L.h[l] = (l == m.skip_layer) ? 0 : take(P * m.width);
    L.head = take(P * m.head_out);
    L.xv = take(m.has_view ? P * L.xv_ld : 1);
    L.hv = take(m.has_view ? P * m.view_width : 1);
    L.rgbraw = take(P * 3);
    L.dya = take(P * m.width);
    L.dyb = take(P * m.width);
    L.dhv = take(m.has_view ? P * m.view_width : 1);
    L.dhead = take(P * 4);
    L.total = off;
    return L;
}

size_t simt_workspace_bytes(const MlpDims& m, const snerf_mlp_desc&, int n_rays, int n_samples, uint32_t) {
    return simt_layout(m, n_rays, n_samples).total * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// encoding
// ------------------------------------------------------------------------------------------------
// out[0..3*(1+2*deg)) = x, sin(x*2^0), cos(x*2^0), sin(x*2^1), ...   (:537-551, sin before cos)
__device__ __forceinline__ void encode3(const float x[3], int degree, float* out, int stride = 1) {
    out[0] = x[0]; out[stride] = x[1]; out[2 * stride] = x[2];
    float freq = 1.f;
    for (int k = 0; k < degree; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s, co;
            sincosf(__fmul_rn(x[c], freq), &s, &co);   // accurate path (range reduced), |arg| up to ~2^9 * |x|
            out[(3 + 6 * k + c) * stride] = s;
            out[(6 + 6 * k + c) * stride] = co;
        }
        freq *= 2.f;
    }
}

__global__ void __launch_bounds__(128) simt_encode_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                          const float* __restrict__ view_dirs, const float* __restrict__ z,
                                                          float* __restrict__ enc, float* __restrict__ venc,
                                                          float* __restrict__ x5, float* __restrict__ xv, int n_rays, int s,
                                                          int pts_degree, int view_degree, int enc_dim, int trunk_in,
                                                          int x5_ld, int xv_ld, int width, int venc_dim) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long P = (long long)n_rays * s;
    if (p >= P) return;
    const int ray = (int)(p / s);
    const float zz = z[p];
    float x[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], zz));   // :140/:142
    float* e = enc + p * enc_dim;
    encode3(x, pts_degree, e);
    for (int c = 0; c < trunk_in; ++c) x5[p * x5_ld + c] = e[c];                    // skip concat: encoding first (:663)
    if (xv != nullptr) {
        float* row = xv + p * xv_ld + width;
        for (int c = trunk_in; c < enc_dim; ++c) *row++ = e[c];                     // :633
        if (venc_dim > 0) {
            float v[3] = {view_dirs[ray * 3], view_dirs[ray * 3 + 1], view_dirs[ray * 3 + 2]};
            float ve[32];
            encode3(v, view_degree, ve);
            for (int c = 0; c < venc_dim; ++c) row[c] = ve[c];                      // :695
            if (p % s == 0)
                for (int c = 0; c < venc_dim; ++c) venc[(size_t)ray * venc_dim + c] = ve[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tiled fp32 GEMM  C[M,N] (+)= A(M,K) * B(K,N)
//   TA=false: A[m*lda+k]   TA=true: A[k*lda+m]       TB=false: B[k*ldb+n]   TB=true: B[n*ldb+k]
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

struct GemmEpilogue {
    const float* bias;   // per n, nullable
    const float* mask;   // multiply by (mask[m*mask_ld+n] > 0), nullable
    int mask_ld;
    bool relu;
    bool accumulate;     // C += result (non-atomic)
    bool atomic;         // split-K: atomicAdd into C
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                    GemmEpilogue ep, int k_per_split) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4] = {};
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        for (int i = tid; i < BM * BK; i += 256) {
            int m, k;
            if (TA) { k = i / BM; m = i % BM; } else { m = i / BK; k = i % BK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < k_end) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
        for (int i = tid; i < BN * BK; i += 256) {
            int n, k;
            if (TB) { n = i / BK; k = i % BK; } else { k = i / BN; n = i % BN; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < k_end) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            float* c = C + (size_t)m * ldc + n;
            if (ep.atomic) {
                atomicAdd(c, v);
                continue;
            }
            if (ep.bias) v += ep.bias[n];
            if (ep.accumulate) v += *c;
            if (ep.relu) v = fmaxf(v, 0.f);
            if (ep.mask) v = ep.mask[(size_t)m * ep.mask_ld + n] > 0.f ? v : 0.f;
            *c = v;
        }
    }
}

template <bool TA, bool TB>
static int gemm(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                GemmEpilogue ep, int splits = 1) {
    if (M == 0 || N == 0) return SNERF_OK;
    int k_per_split = K;
    if (splits > 1) {
        k_per_split = ceil_div(ceil_div(K, splits), BK) * BK;
        splits = ceil_div(K, k_per_split);
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, k_per_split);
    SNERF_LAUNCH_OK("sgemm_kernel");
    return SNERF_OK;
}

static GemmEpilogue ep_linear(const float* bias, bool relu) { return GemmEpilogue{bias, nullptr, 0, relu, false, false}; }
static GemmEpilogue ep_dgrad(const float* mask, int mask_ld, bool accumulate) {
    return GemmEpilogue{nullptr, mask, mask_ld, false, accumulate, false};
}
static GemmEpilogue ep_wgrad() { return GemmEpilogue{nullptr, nullptr, 0, false, false, true}; }

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// sigma = relu(head[:,0] + noise)  (:668-672);  rgb = sigmoid(raw)  (:678 / :706)
__global__ void simt_finalize_kernel(const float* __restrict__ head, int head_out, const float* __restrict__ rgbraw,
                                     const float* __restrict__ noise, float* __restrict__ sigma, float* __restrict__ rgb,
                                     long long P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float s = head[p * head_out];
    if (noise) s += noise[p];
    sigma[p] = fmaxf(s, 0.f);
    const float* raw = (head_out == 4) ? head + p * 4 + 1 : rgbraw + p * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[p * 3 + c] = sigmoidf(raw[c]);
}

// d_head[:,0] = d_sigma * (sigma > 0); d_rgbraw = d_rgb * rgb * (1 - rgb)
__global__ void simt_head_grad_kernel(const float* __restrict__ sigma, const float* __restrict__ rgb,
                                      const float* __restrict__ d_sigma, const float* __restrict__ d_rgb,
                                      float* __restrict__ d_head, int head_out, float* __restrict__ d_rgbraw, long long P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    d_head[p * head_out] = sigma[p] > 0.f ? d_sigma[p] : 0.f;
    float* out = (head_out == 4) ? d_head + p * 4 + 1 : d_rgbraw + p * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = rgb[p * 3 + c];
        out[c] = d_rgb[p * 3 + c] * v * (1.f - v);
    }
}

// bias gradient: column sums of dY [P, n] accumulated into g[n]
__global__ void __launch_bounds__(256) simt_colsum_kernel(const float* __restrict__ dy, int ld, int n, long long P,
                                                          float* __restrict__ g, int rows_per_block) {
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = min(P, r0 + rows_per_block);
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float acc = 0.f;
        for (long long r = r0; r < r1; ++r) acc += dy[r * ld + c];
        atomicAdd(g + c, acc);
    }
}

static int colsum(cudaStream_t st, const float* dy, int ld, int n, long long P, float* g) {
    if (g == nullptr || P == 0) return SNERF_OK;
    const int rows = 256;
    simt_colsum_kernel<<<(int)((P + rows - 1) / rows), n >= 256 ? 256 : (n >= 128 ? 128 : 32), 0, st>>>(dy, ld, n, P, g, rows);
    SNERF_LAUNCH_OK("simt_colsum_kernel");
    return SNERF_OK;
}

#define TRY(expr)                      \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != SNERF_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// forward / backward drivers
// ------------------------------------------------------------------------------------------------
int simt_forward(const snerf_mlp_desc& d, const float* const* prm, const float* rays_o, const float* rays_d,
                 const float* view_dirs, const float* z, const float* noise, float* sigma, float* rgb, void* ws,
                 size_t ws_bytes, int n_rays, int n_samples, uint32_t flags, cudaStream_t st) {
    const MlpDims m(d);
    const SimtLayout L = simt_layout(m, n_rays, n_samples);
    SNERF_REQUIRE(ws_bytes >= L.total * sizeof(float), "mlp_forward: workspace too small (%zu < %zu)", ws_bytes,
                  L.total * sizeof(float));
    const long long P = (long long)n_rays * n_samples;
    if (P == 0) return SNERF_OK;
    float* w = (float*)ws;
    float* xv = m.has_view ? w + L.xv : nullptr;
    simt_encode_kernel<<<(int)((P + 127) / 128), 128, 0, st>>>(rays_o, rays_d, view_dirs, z, w + L.enc, w + L.venc, w + L.x5,
                                                              xv, n_rays, n_samples, d.pts_degree, d.view_degree, m.enc,
                                                              m.trunk_in, L.x5_ld, L.xv_ld, m.width, m.venc);
    SNERF_LAUNCH_OK("simt_encode_kernel");
    // trunk (:659-663)
    const float* x = w + L.x5;   // layer 0 reads the first trunk_in columns of the skip buffer
    int ldx = L.x5_ld;
    for (int l = 0; l < m.depth; ++l) {
        const int K = m.trunk_fan_in(l);
        float* y;
        int ldy;
        if (l == m.skip_layer) { y = w + L.x5 + m.trunk_in; ldy = L.x5_ld; } else { y = w + L.h[l]; ldy = m.width; }
        TRY((gemm<false, true>(st, (int)P, m.width, K, x, ldx, prm[2 * l], K, y, ldy, ep_linear(prm[2 * l + 1], true))));
        if (l == m.skip_layer) { x = w + L.x5; ldx = L.x5_ld; } else { x = y; ldx = ldy; }
    }
    // heads
    TRY((gemm<false, true>(st, (int)P, m.head_out, m.width, x, ldx, prm[SNERF_P_HEAD_W], m.width, w + L.head, m.head_out,
                           ep_linear(prm[SNERF_P_HEAD_B], false))));
    if (m.has_view) {
        TRY((gemm<false, true>(st, (int)P, m.width, m.width, x, ldx, prm[SNERF_P_FEAT_W], m.width, w + L.xv, L.xv_ld,
                               ep_linear(prm[SNERF_P_FEAT_B], false))));
        TRY((gemm<false, true>(st, (int)P, m.view_width, m.view_in, w + L.xv, L.xv_ld, prm[SNERF_P_VIEW_W], m.view_in,
                               w + L.hv, m.view_width, ep_linear(prm[SNERF_P_VIEW_B], true))));
        TRY((gemm<false, true>(st, (int)P, 3, m.view_width, w + L.hv, m.view_width, prm[SNERF_P_RGB_W], m.view_width,
                               w + L.rgbraw, 3, ep_linear(prm[SNERF_P_RGB_B], false))));
    }
    simt_finalize_kernel<<<(int)((P + 255) / 256), 256, 0, st>>>(w + L.head, m.head_out, w + L.rgbraw, noise, sigma, rgb, P);
    SNERF_LAUNCH_OK("simt_finalize_kernel");
    (void)flags;
    return SNERF_OK;
}

int simt_backward(const snerf_mlp_desc& d, const float* const* prm, const float* sigma, const float* rgb,
                  const float* d_sigma, const float* d_rgb, float* const* grads, void* ws, size_t ws_bytes, int n_rays,
                  int n_samples, uint32_t flags, cudaStream_t st) {
    const MlpDims m(d);
    const SimtLayout L = simt_layout(m, n_rays, n_samples);
    SNERF_REQUIRE(ws_bytes >= L.total * sizeof(float), "mlp_backward: workspace too small");
    const long long P = (long long)n_rays * n_samples;
    if (P == 0) return SNERF_OK;
    float* w = (float*)ws;
    const int splits = (int)max(1LL, min(512LL, P / 2048));
    const int Pi = (int)P;

    simt_head_grad_kernel<<<(int)((P + 255) / 256), 256, 0, st>>>(sigma, rgb, d_sigma, d_rgb, w + L.dhead, m.head_out,
                                                                 w + L.rgbraw, P);
    SNERF_LAUNCH_OK("simt_head_grad_kernel");

    // activations feeding the heads: output of the last trunk layer
    const int last = m.depth - 1;
    const float* h_last = (last == m.skip_layer) ? w + L.x5 + m.trunk_in : w + L.h[last];
    const int ld_last = (last == m.skip_layer) ? L.x5_ld : m.width;
    float* dy = w + L.dya;    // gradient w.r.t. the pre-activation of the layer being processed
    float* dx = w + L.dyb;

    // sigma (/rgb) head: dW = d_head^T h, db, and d h_last (no mask yet)
    TRY((gemm<true, false>(st, m.head_out, m.width, Pi, w + L.dhead, m.head_out, h_last, ld_last, grads[SNERF_P_HEAD_W],
                           m.width, ep_wgrad(), splits)));
    TRY(colsum(st, w + L.dhead, m.head_out, m.head_out, P, grads[SNERF_P_HEAD_B]));
    if (m.has_view) {
        // rgb head
        TRY((gemm<true, false>(st, 3, m.view_width, Pi, w + L.rgbraw, 3, w + L.hv, m.view_width, grads[SNERF_P_RGB_W],
                               m.view_width, ep_wgrad(), splits)));
        TRY(colsum(st, w + L.rgbraw, 3, 3, P, grads[SNERF_P_RGB_B]));
        // d hv = d_rgbraw W_rgb, masked by relu
        const bool vis_grad = (flags & SNERF_FLAG_VIS_GRAD) != 0;   // simt_visibility_backward pre-filled dhv and dyb
        TRY((gemm<false, false>(st, Pi, m.view_width, 3, w + L.rgbraw, 3, prm[SNERF_P_RGB_W], m.view_width, w + L.dhv,
                                m.view_width, ep_dgrad(w + L.hv, m.view_width, vis_grad))));
        // view layer
        TRY((gemm<true, false>(st, m.view_width, m.view_in, Pi, w + L.dhv, m.view_width, w + L.xv, L.xv_ld,
                               grads[SNERF_P_VIEW_W], m.view_in, ep_wgrad(), splits)));
        TRY(colsum(st, w + L.dhv, m.view_width, m.view_width, P, grads[SNERF_P_VIEW_B]));
        // d feature = d_hv W_view[:, :width]  (feature has no activation)
        TRY((gemm<false, false>(st, Pi, m.width, m.view_width, w + L.dhv, m.view_width, prm[SNERF_P_VIEW_W], m.view_in, dx,
                                m.width, ep_dgrad(nullptr, 0, vis_grad))));
        // feature layer
        TRY((gemm<true, false>(st, m.width, m.width, Pi, dx, m.width, h_last, ld_last, grads[SNERF_P_FEAT_W], m.width,
                               ep_wgrad(), splits)));
        TRY(colsum(st, dx, m.width, m.width, P, grads[SNERF_P_FEAT_B]));
        // d h_last = d_feature W_feat + d_head W_head, then relu mask
        TRY((gemm<false, false>(st, Pi, m.width, m.width, dx, m.width, prm[SNERF_P_FEAT_W], m.width, dy, m.width,
                                ep_dgrad(nullptr, 0, false))));
        TRY((gemm<false, false>(st, Pi, m.width, m.head_out, w + L.dhead, m.head_out, prm[SNERF_P_HEAD_W], m.width, dy,
                                m.width, ep_dgrad(h_last, ld_last, true))));
    } else {
        TRY((gemm<false, false>(st, Pi, m.width, m.head_out, w + L.dhead, m.head_out, prm[SNERF_P_HEAD_W], m.width, dy,
                                m.width, ep_dgrad(h_last, ld_last, false))));
    }
    // trunk, last layer first.  dy = gradient w.r.t. layer l's pre-activation (already relu-masked)
    for (int l = last; l >= 0; --l) {
        const int K = m.trunk_fan_in(l);
        const float* x;
        int ldx;
        if (l == 0) { x = w + L.x5; ldx = L.x5_ld; }
        else if (l - 1 == m.skip_layer) { x = w + L.x5; ldx = L.x5_ld; }
        else { x = w + L.h[l - 1]; ldx = m.width; }
        TRY((gemm<true, false>(st, m.width, K, Pi, dy, m.width, x, ldx, grads[2 * l], K, ep_wgrad(), splits)));
        TRY(colsum(st, dy, m.width, m.width, P, grads[2 * l + 1]));
        if (l == 0) break;
        // dgrad into the previous layer's output (hidden part only: the encoding columns need no gradient)
        const int col0 = (l - 1 == m.skip_layer) ? m.trunk_in : 0;
        const float* hprev = x + col0;
        TRY((gemm<false, false>(st, Pi, m.width, m.width, dy, m.width, prm[2 * l] + col0, K, dx, m.width,
                                ep_dgrad(hprev, ldx, false))));
        float* t = dy; dy = dx; dx = t;
    }
    (void)flags;
    return SNERF_OK;
}


// ------------------------------------------------------------------------------------------------
// Row a14 / N4: the secondary-view visibility head (predict_visibility=True), precise path only.
// Reference: MLP ctor :596-608 (views_output_linear has a fourth row), MLP.forward :640-649 (the view branch once more per
// other view), get_view_dependent_outputs :687-715, compute_other_view_dirs :317-325.
// Runs on the workspace a snerf_mlp_forward(PRECISE | SAVE_FOR_BWD) of the same MLP left behind: xv = [feature | enc_hi |
// PE(view_dir)] and hv = relu(view layer) per point.  Own workspace: xv2 [P, xv_ld], hv2 [n_other][P, view_width],
// dv [P, 1 + n_other] (pre-sigmoid gradients), dhv2 [P, view_width].
// ------------------------------------------------------------------------------------------------
struct VisLayout {
    size_t xv2, hv2, dv, dhv2, total;
};

static VisLayout vis_layout(const MlpDims& m, int n_rays, int n_samples, int n_other) {
    VisLayout V{};
    const size_t P = (size_t)n_rays * n_samples;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align_up(n < 1 ? 1 : n, 64); return o; };
    V.xv2 = take(P * m.view_in);
    V.hv2 = take((size_t)n_other * P * m.view_width);
    V.dv = take(P * (1 + n_other));
    V.dhv2 = take(P * m.view_width);
    V.total = off;
    return V;
}

size_t simt_visibility_workspace_bytes(const MlpDims& m, int n_rays, int n_samples, int n_other) {
    return vis_layout(m, n_rays, n_samples, n_other).total * sizeof(float);
}

// xv2[p] = [xv[p, :view_in - venc] | PE(unit vector from the other view's camera centre to the point)]
__global__ void __launch_bounds__(128) vis_build_input_kernel(const float* __restrict__ xv, const float* __restrict__ rays_o,
                                                              const float* __restrict__ rays_d, const float* __restrict__ z,
                                                              const float* __restrict__ rays_o2, float* __restrict__ xv2,
                                                              int n_rays, int s, int n_other, int v, int xv_ld, int keep,
                                                              int view_degree, bool ndc) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)n_rays * s) return;
    const int ray = (int)(p / s);
    const float ox = rays_o[ray * 3], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
    const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
    float zz = z[p];
    if (ndc) {                                                                                   // :319-321 (near = 1)
        const float tn = -(1.f + oz) / dz;
        zz = (((oz + tn * dz) / (1.f - zz + 1e-6f)) - oz) / dz;
    }
    const float* o2 = rays_o2 + ((size_t)ray * n_other + v) * 3;
    float d[3] = {__fadd_rn(ox, __fmul_rn(zz, dx)) - o2[0], __fadd_rn(oy, __fmul_rn(zz, dy)) - o2[1],
                  __fadd_rn(oz, __fmul_rn(zz, dz)) - o2[2]};                                     // :322-323
    const float norm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                          // :324
    d[0] /= norm; d[1] /= norm; d[2] /= norm;
    const float* src = xv + p * xv_ld;
    float* dst = xv2 + p * xv_ld;
    for (int c = 0; c < keep; ++c) dst[c] = src[c];
    float ve[32];
    encode3(d, view_degree, ve);
    for (int c = 0; c < xv_ld - keep; ++c) dst[keep + c] = ve[c];
}

__global__ void vis_sigmoid_kernel(float* __restrict__ x, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = sigmoidf(x[i]);
}

// dv[p, 0] = d_vis * vis (1 - vis); dv[p, 1 + v] = d_vis2 * vis2 (1 - vis2)   (null gradients count as zero)
__global__ void vis_head_grad_kernel(const float* __restrict__ vis, const float* __restrict__ vis2,
                                     const float* __restrict__ d_vis, const float* __restrict__ d_vis2,
                                     float* __restrict__ dv, long long P, int n_other) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float a = vis[p];
    dv[p * (1 + n_other)] = d_vis ? d_vis[p] * a * (1.f - a) : 0.f;
    for (int v = 0; v < n_other; ++v) {
        const float b = vis2[p * n_other + v];
        dv[p * (1 + n_other) + 1 + v] = d_vis2 ? d_vis2[p * n_other + v] * b * (1.f - b) : 0.f;
    }
}

int simt_visibility_forward(const snerf_mlp_desc& d, const float* const* prm, const void* mlp_ws, const float* rays_o,
                            const float* rays_d, const float* z, const float* rays_o2, float* visibility, float* visibility2,
                            void* vis_ws, size_t vis_ws_bytes, int n_rays, int n_samples, int n_other, uint32_t flags,
                            cudaStream_t st) {
    const MlpDims m(d);
    SNERF_REQUIRE(m.has_view && m.venc > 0, "visibility head: the MLP needs a view branch with view directions");
    const SimtLayout L = simt_layout(m, n_rays, n_samples);
    const VisLayout V = vis_layout(m, n_rays, n_samples, n_other);
    SNERF_REQUIRE(vis_ws_bytes >= V.total * sizeof(float), "visibility_forward: workspace too small (%zu < %zu)", vis_ws_bytes,
                  V.total * sizeof(float));
    const long long P = (long long)n_rays * n_samples;
    if (P == 0) return SNERF_OK;
    const float* w = (const float*)mlp_ws;
    float* vw = (float*)vis_ws;
    const float* w_vis = prm[SNERF_P_RGB_W] + 3 * m.view_width;      // fourth row of views_output_linear (:710-711)
    const float* b_vis = prm[SNERF_P_RGB_B] + 3;
    // own view: visibility = sigmoid(hv . w_vis + b_vis)
    TRY((gemm<false, true>(st, (int)P, 1, m.view_width, w + L.hv, m.view_width, w_vis, m.view_width, visibility, 1,
                           ep_linear(b_vis, false))));
    vis_sigmoid_kernel<<<(int)((P + 255) / 256), 256, 0, st>>>(visibility, P);
    SNERF_LAUNCH_OK("vis_sigmoid_kernel");
    for (int v = 0; v < n_other; ++v) {                                                          // :646-649
        vis_build_input_kernel<<<(int)((P + 127) / 128), 128, 0, st>>>(w + L.xv, rays_o, rays_d, z, rays_o2, vw + V.xv2, n_rays,
                                                                      n_samples, n_other, v, L.xv_ld, m.view_in - m.venc,
                                                                      d.view_degree, (flags & SNERF_FLAG_NDC) != 0);
        SNERF_LAUNCH_OK("vis_build_input_kernel");
        float* hv2 = vw + V.hv2 + (size_t)v * P * m.view_width;
        TRY((gemm<false, true>(st, (int)P, m.view_width, m.view_in, vw + V.xv2, L.xv_ld, prm[SNERF_P_VIEW_W], m.view_in, hv2,
                               m.view_width, ep_linear(prm[SNERF_P_VIEW_B], true))));
        TRY((gemm<false, true>(st, (int)P, 1, m.view_width, hv2, m.view_width, w_vis, m.view_width, visibility2 + v, n_other,
                               ep_linear(b_vis, false))));
    }
    if (n_other > 0) {
        vis_sigmoid_kernel<<<(int)((P * n_other + 255) / 256), 256, 0, st>>>(visibility2, P * n_other);
        SNERF_LAUNCH_OK("vis_sigmoid_kernel");
    }
    return SNERF_OK;
}

// Must run BEFORE simt_backward(... | SNERF_FLAG_VIS_GRAD) of the same step: it leaves d hv (own view) in the MLP workspace's
// dhv region and the other views' d feature in its dyb region, which that call then accumulates onto.
int simt_visibility_backward(const snerf_mlp_desc& d, const float* const* prm, void* mlp_ws, const float* rays_o,
                             const float* rays_d, const float* z, const float* rays_o2, const float* visibility,
                             const float* visibility2, const float* d_visibility, const float* d_visibility2,
                             float* const* grads, void* vis_ws, size_t vis_ws_bytes, int n_rays, int n_samples, int n_other,
                             uint32_t flags, cudaStream_t st) {
    const MlpDims m(d);
    const SimtLayout L = simt_layout(m, n_rays, n_samples);
    const VisLayout V = vis_layout(m, n_rays, n_samples, n_other);
    SNERF_REQUIRE(vis_ws_bytes >= V.total * sizeof(float), "visibility_backward: workspace too small");
    const long long P = (long long)n_rays * n_samples;
    if (P == 0) return SNERF_OK;
    float* w = (float*)mlp_ws;
    float* vw = (float*)vis_ws;
    const int Pi = (int)P, ldv = 1 + n_other;
    const int splits = (int)max(1LL, min(512LL, P / 2048));
    const float* w_vis = prm[SNERF_P_RGB_W] + 3 * m.view_width;
    float* g_w_vis = grads[SNERF_P_RGB_W] + 3 * m.view_width;
    float* g_b_vis = grads[SNERF_P_RGB_B] + 3;
    vis_head_grad_kernel<<<(int)((P + 255) / 256), 256, 0, st>>>(visibility, visibility2, d_visibility, d_visibility2, vw + V.dv, P,
                                                                n_other);
    SNERF_LAUNCH_OK("vis_head_grad_kernel");
    // own view: row-3 weight / bias gradients, and d hv left (relu-masked) for the main backward to accumulate onto
    TRY((gemm<true, false>(st, 1, m.view_width, Pi, vw + V.dv, ldv, w + L.hv, m.view_width, g_w_vis, m.view_width, ep_wgrad(), splits)));
    TRY(colsum(st, vw + V.dv, ldv, 1, P, g_b_vis));
    TRY((gemm<false, false>(st, Pi, m.view_width, 1, vw + V.dv, ldv, w_vis, m.view_width, w + L.dhv, m.view_width,
                            ep_dgrad(w + L.hv, m.view_width, false))));
    for (int v = 0; v < n_other; ++v) {
        const float* hv2 = vw + V.hv2 + (size_t)v * P * m.view_width;
        const float* dv = vw + V.dv + 1 + v;
        TRY((gemm<true, false>(st, 1, m.view_width, Pi, dv, ldv, hv2, m.view_width, g_w_vis, m.view_width, ep_wgrad(), splits)));
        TRY(colsum(st, dv, ldv, 1, P, g_b_vis));
        TRY((gemm<false, false>(st, Pi, m.view_width, 1, dv, ldv, w_vis, m.view_width, vw + V.dhv2, m.view_width,
                                ep_dgrad(hv2, m.view_width, false))));
        // view layer of this pass: its input is rebuilt (feature and enc_hi copied, the other view's direction encoded)
        vis_build_input_kernel<<<(int)((P + 127) / 128), 128, 0, st>>>(w + L.xv, rays_o, rays_d, z, rays_o2, vw + V.xv2, n_rays,
                                                                      n_samples, n_other, v, L.xv_ld, m.view_in - m.venc,
                                                                      d.view_degree, (flags & SNERF_FLAG_NDC) != 0);
        SNERF_LAUNCH_OK("vis_build_input_kernel");
        TRY((gemm<true, false>(st, m.view_width, m.view_in, Pi, vw + V.dhv2, m.view_width, vw + V.xv2, L.xv_ld, grads[SNERF_P_VIEW_W],
                               m.view_in, ep_wgrad(), splits)));
        TRY(colsum(st, vw + V.dhv2, m.view_width, m.view_width, P, grads[SNERF_P_VIEW_B]));
        // d feature of the other views, summed in the main backward's d-feature buffer
        TRY((gemm<false, false>(st, Pi, m.width, m.view_width, vw + V.dhv2, m.view_width, prm[SNERF_P_VIEW_W], m.view_in, w + L.dyb,
                                m.width, ep_dgrad(nullptr, 0, v > 0))));
    }
    if (n_other == 0) SNERF_CUDA_OK(cudaMemsetAsync(w + L.dyb, 0, (size_t)P * m.width * sizeof(float), st));
    return SNERF_OK;
}

}  // namespace snerf
