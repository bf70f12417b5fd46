// bf16 tcgen05 MLP (the throughput path): positional encoding + 8x256 trunk + heads as ONE persistent,
// warp-specialised kernel per MLP evaluation.
//
// Reference behaviour: src/models/SimpleNeRF01.py  run_network :363-392, PositionalEncoder :525-557,
// MLP.forward :626-654, get_view_independent_outputs :656-685, get_view_dependent_outputs :687-715.
//
// Forward kernel: 74 clusters of two CTAs (one CTA per SM, 736 threads); a pair works on 256-point super tiles with
// tcgen05.mma.cta_group::2 and keeps two of them in flight (see the comment above the kernel):
//   warps 0-3   encoder        rays + depth -> point -> positional encoding (kept in registers, the single encoding
//                              panel is rewritten before each use)
//   warp 4      stash writer   (training) one 64 KB bulk copy of the slot's activation panels per job -> HBM
//   warp 5      weight loader  this CTA's half of every packed bf16 weight chunk, L2 -> smem ring by bulk async copy
//   warps 6-21  epilogue       TMEM -> regs -> bias/ReLU -> bf16 -> swizzled smem panel (the next layer's A operand);
//                              four warps per SM sub-partition, one 64-column panel each; hidden layers use one packed
//                              HFMA2.BF16.RELU per value pair; the rgb head (and the sigma + rgb head of an MLP without a
//                              view branch) are fp32 dot products on the un-rounded activations, the sigma head of an MLP
//                              with a view branch is accumulator column 128 of the view step (tc_plan.cuh); in training
//                              also the ReLU sign bits for the backward pass
//   warp 22     MMA issuer (leader CTA: converged warp, one elected lane) / weight relay (peer CTA)
// Activations never leave the SM in eval mode.
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"
#include "tc_plan.cuh"
#include "rng.cuh"

namespace snerf {
using namespace tc;

// packed image = bf16 weight chunks, then (MLPs with a view branch) the fp32 merged matrix W_vf the chunks were cut from
static size_t packed_chunk_bytes(const snerf_mlp_desc& d) { return align_up(build_plan(d, nullptr).packed_bytes, 1024); }
size_t tc_packed_bytes(const snerf_mlp_desc& d) { return packed_chunk_bytes(d) + (d.view_width > 0 ? kMergedWeightBytes : 0); }

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 torch.nn.Linear weights -> bf16 swizzled [N x 64] chunks
// ------------------------------------------------------------------------------------------------
struct PackParams {
    int n;
    PackChunk c[kMaxPack];
};

__global__ void __launch_bounds__(256) tc_pack_kernel(const __grid_constant__ PackParams pp, uint8_t* __restrict__ packed) {
    const PackChunk& c = pp.c[blockIdx.y];
    const int total = c.n_rows * 8;   // one thread per 16-byte chunk (8 bf16)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i >> 3, ch = i & 7;
        const bool has_src = n < c.src_rows;
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float f[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = ch * 8 + h * 2 + e;
                float v = 0.f;
                if (has_src && kk >= c.k_lo && kk < c.k_hi) {
                    v = c.transposed ? c.src[(size_t)(c.row0 + kk - c.k_lo) * c.ld + c.col0 + n]
                                     : c.src[(size_t)(c.row0 + n) * c.ld + c.col0 + (kk - c.k_lo)];
                }
                f[e] = v;
            }
            w[h] = pack_bf16(f[0], f[1]);
        }
        *reinterpret_cast<uint4*>(packed + c.dst_off + swz_offset(c.dst_row0 + n, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// W_vf[o][k] = sum_j W_view[o][j] W_feat[j][k]   (fp32; 128 x 256 x 256): blocks 0..63 form 16 x 32 output tiles,
// block 64 the feature bias seen through the view layer, bvf[o] = sum_j W_view[o][j] b_feat[j], stored behind W_vf
__global__ void __launch_bounds__(256) tc_merge_view_feature_kernel(const float* __restrict__ w_view, int view_in,
                                                                    const float* __restrict__ w_feat, const float* __restrict__ b_feat,
                                                                    float* __restrict__ wvf) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (blockIdx.x == 64) {                                       // warp ty takes rows ty, ty + 8, ...
        for (int o = ty; o < 128; o += 8) {
            float a = 0.f;
            for (int j = tx; j < 256; j += 32) a = fmaf(w_view[(size_t)o * view_in + j], b_feat[j], a);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
            if (tx == 0) wvf[128 * 256 + o] = a;
        }
        return;
    }
    __shared__ float sa[16][33], sb[32][33];
    const int o0 = (blockIdx.x >> 3) * 16, k0 = (blockIdx.x & 7) * 32;
    float acc[2] = {0.f, 0.f};
    for (int j0 = 0; j0 < 256; j0 += 32) {
#pragma unroll
        for (int i = 0; i < 2; ++i) sa[ty + 8 * i][tx] = w_view[(size_t)(o0 + ty + 8 * i) * view_in + j0 + tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) sb[ty + 8 * i][tx] = w_feat[(size_t)(j0 + ty + 8 * i) * 256 + k0 + tx];
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float w = sb[c][tx];
            acc[0] = fmaf(sa[ty][c], w, acc[0]);
            acc[1] = fmaf(sa[ty + 8][c], w, acc[1]);
        }
        __syncthreads();
    }
    wvf[(size_t)(o0 + ty) * 256 + k0 + tx] = acc[0];
    wvf[(size_t)(o0 + ty + 8) * 256 + k0 + tx] = acc[1];
}

int tc_pack(const snerf_mlp_desc& d, const float* const* prm, void* packed, cudaStream_t st) {
    float* wvf = nullptr;
    if (d.view_width > 0) {
        wvf = (float*)((uint8_t*)packed + packed_chunk_bytes(d));
        tc_merge_view_feature_kernel<<<65, 256, 0, st>>>(prm[SNERF_P_VIEW_W], MlpDims(d).view_in, prm[SNERF_P_FEAT_W], prm[SNERF_P_FEAT_B], wvf);
        SNERF_LAUNCH_OK("tc_merge_view_feature_kernel");
    }
    const TcPlan pl = build_plan(d, prm, wvf);
    PackParams pp;
    pp.n = pl.n_pack;
    for (int i = 0; i < pl.n_pack; ++i) pp.c[i] = pl.pack[i];
    tc_pack_kernel<<<dim3(8, pl.n_pack), 256, 0, st>>>(pp, (uint8_t*)packed);
    SNERF_LAUNCH_OK("tc_pack_kernel");
    return SNERF_OK;
}

// per-ray part of the view layer: vb[ray][o] = b_view[o] + sum_c W_view[o][col0 + c] * PE(view_dir)[c]   (:640, :695)
//                                              + sum_j W_view[o][j] b_feat[j]   (the feature bias seen through the merged matrix)
// One block per kRaysPerBlock rays: the encodings are computed element-parallel (accurate sin/cos: fp32 consumers), then
// thread o keeps its weight row in registers and forms the output of every ray of the block.
constexpr int kVbRays = 16;
__global__ void __launch_bounds__(128) tc_view_bias_kernel(const float* __restrict__ view_dirs, const float* __restrict__ w_view,
                                                           const float* __restrict__ b_view, const float* __restrict__ b_merged,
                                                           float* __restrict__ vb, int n_rays,
                                                           int view_degree, int view_in, int col0, int venc,
                                                           const float* __restrict__ pts_d, const float* __restrict__ cam_o,
                                                           const float* __restrict__ cam_d, float4* __restrict__ rayc,
                                                           float* __restrict__ view_enc) {
    const int ray0 = blockIdx.x * kVbRays;
    // fused evaluation: the per-ray constants of the compositing arithmetic -- |d| (:436 / :441) and, for NDC depths, the two
    // factors of convert_depth_from_ndc (:498, :501), with composite.cu's expressions
    if (rayc != nullptr && threadIdx.x < kVbRays && ray0 + threadIdx.x < n_rays) {
        const int ray = ray0 + threadIdx.x;
        const float d0 = pts_d[ray * 3], d1 = pts_d[ray * 3 + 1], d2 = pts_d[ray * 3 + 2];
        float tn = 0.f, k0 = 0.f;
        if (cam_o != nullptr) {
            const float oz = cam_o[ray * 3 + 2], dz = cam_d[ray * 3 + 2];
            tn = -(1.f + oz) / dz;
            k0 = (oz + tn * dz) / dz;
        }
        rayc[ray] = make_float4(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), tn, k0, 0.f);
    }
    __shared__ float ve[kVbRays][32];
    for (int e = threadIdx.x; e < kVbRays * 32; e += blockDim.x) {
        const int r = e >> 5, idx = e & 31, ray = ray0 + r;
        float val = 0.f;                                                   // columns >= venc and rays past the end stay zero
        if (ray < n_rays && idx < venc) {
            if (idx < 3) {
                val = view_dirs[ray * 3 + idx];
            } else {
                const int j = idx - 3, k = j / 6, rem = j - 6 * k, c = rem >= 3 ? rem - 3 : rem;
                const float x = view_dirs[ray * 3 + c] * (float)(1 << k);          // same arithmetic as encode_point_accurate
                val = rem >= 3 ? cosf(x) : sinf(x);
            }
        }
        ve[r][idx] = val;
        if (view_enc != nullptr && ray < n_rays) view_enc[(size_t)ray * 32 + idx] = val;     // visibility head: PE(view_dir) per ray (vis_tc.cu)
    }
    __syncthreads();
    const int o = threadIdx.x;
    float w[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) w[c] = c < venc ? w_view[(size_t)o * view_in + col0 + c] : 0.f;
    const float bias = b_view[o] + b_merged[o];
    for (int r = 0; r < kVbRays && ray0 + r < n_rays; ++r) {
        float acc = bias;
#pragma unroll
        for (int c = 0; c < 32; ++c) acc = fmaf(w[c], ve[r][c], acc);
        vb[(size_t)(ray0 + r) * 128 + o] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// forward kernel (CTA pairs, two tiles in flight)
// ------------------------------------------------------------------------------------------------
// A cluster of two CTAs works on "super tiles" of 256 points: CTA r owns rows [128 r, 128 r + 128) -- its activation
// panels, its half of every weight chunk (N/2 rows of B) and its 128 accumulator rows in its own TMEM -- and the
// leader's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256) for both.  Each pair keeps TWO super tiles in
// flight (slots 0/1, one 256-column accumulator each) and the issuer alternates between them step by step, so the
// epilogue of one tile (TMEM -> bias/ReLU -> bf16 -> smem) runs under the MMAs of the other.
// Measured (tools/pair_probe.py, tools/mma_rate.py): 129 cycles per M256 N256 K16 pair MMA against 165 cycles per
// M128 N256 K16 single-CTA MMA, i.e. 2.5x the issue rate per row, and half the weight bytes per SM.
constexpr int kFwdThreads = 736;                                       // 23 warps, see the role table
constexpr int kEpiWarps = 16;                                           // 4 per SM sub-partition, one 64-column panel each
constexpr int kPStages = 4;                                             // weight ring: this CTA's half chunks; 4 = one whole layer
constexpr uint32_t kPStageBytes = 16384;
constexpr uint32_t kOffH = 0;                                           // [2 slots][4 panels]
constexpr uint32_t kOffE = 2 * 65536;                                   // encoding panel (one, rewritten before every use)
constexpr uint32_t kOffRing = kOffE + kPanelBytes;                      // 147456
constexpr uint32_t kOffConst = kOffRing + kPStages * kPStageBytes;      // 212992
constexpr uint32_t kConstBias16 = 0;                                    // [9][256] bf16 (packed bias + ReLU path)
constexpr uint32_t kConstBias32 = 9 * 256 * 2;                          // [256] fp32: bias of the last trunk layer (fp32 head path)
constexpr uint32_t kConstHeadW = kConstBias32 + 256 * 4;                // [4][256] fp32
constexpr uint32_t kConstRgbW = kConstHeadW + 4 * 256 * 4;              // [3][128] fp32
constexpr uint32_t kConstMisc = kConstRgbW + 3 * 128 * 4;               // head bias[4], rgb bias[4]
constexpr uint32_t kConstPart = kConstMisc + 64;                        // [3][128][4] fp32 partial head sums of panels 1-3
constexpr uint32_t kConstBytes = 17920;
constexpr uint32_t kOffBars = kOffConst + kConstBytes;
constexpr uint32_t kFwdSmem = kOffBars + 512 + 1024;                    // + alignment slack
static_assert(kConstPart + 3 * 128 * 4 * 4 <= kConstBytes, "constant area overflow");
static_assert(kFwdSmem <= 232448, "shared memory budget");

struct FwdParams {
    const uint8_t* packed;
    const float* bias[9];          // trunk 0..7, feature
    const float *w_head, *b_head, *w_rgb, *b_rgb;
    const float* view_bias;        // [n_rays,128]
    const float *rays_o, *rays_d, *z, *noise;
    RngKey noise_rng;              // use_rng: sigma noise drawn in the head epilogue (element = point index) instead of read from `noise`
    float noise_std;
    int use_rng;
    float *sigma, *rgb;            // null in the fused evaluation (the values stay in registers)
    // fused evaluation (row X1): compositing arithmetic in the rgb-head epilogue, one record per 32 samples of a ray
    float* seg;                    // [n_points / 32][kSegFloats]; null = off
    float *alpha_out, *wloc;       // nullable [n_points]: alpha, and the weights relative to the segment's first sample
    const float4* rayc;            // per ray: |d|, and the two factors of the NDC depth conversion (tc_view_bias_kernel)
    int ndc;
    uint8_t* stash;                // null in eval
    uint8_t* bits;                 // (training) ReLU sign bits of the trunk activations, kBitsTileBytes per tile
    int stash_pieces;              // 4 (default): the stash leaves one panel per bulk-copy request, paced; 1: one request per job (SNERF_STASH_PIECES)
    uint8_t* vis_pre;              // (visibility head, SNERF_FLAG_VIS_HEAD) bf16 [n_points,128]: the view layer's accumulator without the per-ray bias; else null
    long long* trace;              // debug: clock64 timestamps of pair 0 (tools/trace_fwd.py), normally null
    int debug;                     // debug (timing experiments, results become garbage): bit0 no panel stores, bit1 no TMEM loads / epilogue math, bit2 no weight copies
    long long n_points;
    int n_samples, n_tiles, n_steps, pts_degree, head_out;
    uint32_t tile_stash_bytes;
    TcStep steps[kMaxSteps];
};

// One record per warp-aligned run of 32 samples (n_samples % 32 == 0, so a run never straddles two rays).  With
// w' = alpha * (transmittance counted from the run's first sample) and the run's first depths as reference points:
//   P = prod (1 - alpha + 1e-10), A = sum w', C = sum w' rgb, S1 = sum w' (z - z_ref), S2 = sum w' (z - z_ref)^2
// (metric and ndc depths).  composite_fold_kernel multiplies by the transmittance that reaches the run and adds the runs of a
// ray up; the second moment about the ray's depth follows from S1, S2 without cancellation (z - z_ref spans one run only).
constexpr int kSegFloats = 12;
enum { SEG_P = 0, SEG_A, SEG_CR, SEG_CG, SEG_CB, SEG_S1M, SEG_S1N, SEG_S2M, SEG_S2N, SEG_ZM, SEG_ZN };

struct FwdBars {
    uint64_t w_full[kPStages];     // leader: own bytes + the peer's relay (2 arrivals); peer: own bytes (1)
    uint64_t w_empty[kPStages];    // MMA commit, multicast to both CTAs
    uint64_t acc_full[2];          // per slot: MMA commit, multicast
    uint64_t tile_ready[2];        // leader only: 2 x 8 epilogue warps -- accumulator drained, next A operand in smem
    uint64_t stash_ready[2];       // local: 8 epilogue warps
    uint64_t stash_done[2];        // local: stash writer finished reading the slot's panels
    uint64_t enc_ready;            // leader only: 2 x 4 encoder warps; one phase per use of the encoding panel
    uint64_t enc_free;             // MMA commit, multicast: the MMAs reading the encoding panel have completed
    uint32_t tmem_base;
};

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

// bf16x2 (acc + bias), optionally ReLU, in one HFMA2: inputs are the packed accumulator pair and the packed bias pair
__device__ __forceinline__ uint32_t bias_act_bf16x2(float lo, float hi, uint32_t bias2, bool relu) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(lo, hi);
    const __nv_bfloat162 one = __floats2bfloat162_rn(1.f, 1.f);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&bias2);
    const __nv_bfloat162 r = relu ? __hfma2_relu(x, one, b) : __hfma2(x, one, b);
    return *reinterpret_cast<const uint32_t*>(&r);
}

// kSave: training (activation stash + ReLU sign bits); a template parameter so that the epilogue's inner loop carries no test of it
template <bool kTrace, bool kSave>
__global__ void __launch_bounds__(kFwdThreads, 1) tc_forward_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    FwdBars* bars = (FwdBars*)(smem + kOffBars);
    __nv_bfloat16* s_bias16 = (__nv_bfloat16*)(smem + kOffConst + kConstBias16);
    float* s_bias32 = (float*)(smem + kOffConst + kConstBias32);
    float* s_whead = (float*)(smem + kOffConst + kConstHeadW);
    float* s_wrgb = (float*)(smem + kOffConst + kConstRgbW);
    float* s_misc = (float*)(smem + kOffConst + kConstMisc);
    float* s_part = (float*)(smem + kOffConst + kConstPart);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    constexpr bool save = kSave;
    const int dbg = kTrace ? p.debug : 0;          // stage switches exist in the traced build only
    const uint32_t rank = cluster_rank();
    // roles by warp id -- the sub-partition arbiter issues the highest eligible warp id first, so the latency-critical
    // epilogue warps sit on top and the helpers below them:
    //   0-3 encoders | 4 stash writer | 5 weight loader | 6-21 epilogue | 22 MMA issuer (leader) / weight relay (peer)
    constexpr int kWarpStash = 4, kWarpLoader = 5, kWarpEpi0 = 6, kWarpMma = 22;

    // ---- setup ----
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPStages; ++i) { mbar_init(&bars->w_full[i], rank == 0 ? 2 : 1); mbar_init(&bars->w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->acc_full[i], 1);
            mbar_init(&bars->tile_ready[i], 2 * kEpiWarps);
            mbar_init(&bars->stash_ready[i], kEpiWarps);
            mbar_init(&bars->stash_done[i], 1);
        }
        mbar_init(&bars->enc_ready, 2 * 4);
        mbar_init(&bars->enc_free, 1);
        mbar_fence_init();
    }
    if (warp == kWarpMma) tmem_alloc2<512>(&bars->tmem_base);
    for (int i = threadIdx.x; i < 9 * 256; i += kFwdThreads) {
        const float* b = p.bias[i >> 8];
        s_bias16[i] = __float2bfloat16_rn(b ? b[i & 255] : 0.f);
    }
    for (int i = threadIdx.x; i < 256; i += kFwdThreads) s_bias32[i] = p.bias[7][i];
    for (int i = threadIdx.x; i < p.head_out * 256; i += kFwdThreads) s_whead[i] = p.w_head[i];
    if (p.w_rgb) for (int i = threadIdx.x; i < 3 * 128; i += kFwdThreads) s_wrgb[i] = p.w_rgb[i];
    if (threadIdx.x < 4) s_misc[threadIdx.x] = threadIdx.x < p.head_out ? p.b_head[threadIdx.x] : 0.f;
    if (threadIdx.x >= 4 && threadIdx.x < 8) s_misc[threadIdx.x] = (p.b_rgb && threadIdx.x < 7) ? p.b_rgb[threadIdx.x - 4] : 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers are initialised before either signals the other
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    // pair pi takes super tiles pi, pi + n_pairs, ...; every pair runs the same count (tiles >= n_tiles are dummies)
    const int n_pairs = (int)gridDim.x / 2, pi = (int)blockIdx.x / 2;
    const int n_super = (p.n_tiles + 1) / 2;
    const int my_super = (n_super + n_pairs - 1) / n_pairs;
    auto tile_of = [&](int i) { return 2 * (pi + i * n_pairs) + (int)rank; };

    if (warp == kWarpLoader) {
        // ======================= weight loader: this CTA's N/2 rows of every chunk =======================
        // When both slots are occupied and a step's chunks fit the ring, the chunks are loaded ONCE and used by both
        // tiles (the issuer releases a stage after the second tile's MMAs); otherwise once per tile.
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s) {
                    const TcStep& st = p.steps[s];
                    const uint32_t bytes = (uint32_t)st.n_rows * kRowBytes, half = bytes / 2;
                    const int reps = (two && st.n_chunks > kPStages) ? 2 : 1;
                    for (int r = 0; r < reps; ++r)
                        for (int c = 0; c < st.n_chunks; ++c, ++cnt) {
                            const uint32_t stage = cnt % kPStages, round = cnt / kPStages;
                            if (round > 0) mbar_wait(&bars->w_empty[stage], (round - 1) & 1);
                            if (dbg & 4) { mbar_arrive(&bars->w_full[stage]); continue; }
                            mbar_arrive_expect_tx(&bars->w_full[stage], half);
                            bulk_g2s(smem + kOffRing + stage * kPStageBytes, p.packed + st.w_off + (uint32_t)c * bytes + rank * half, half,
                                     &bars->w_full[stage]);
                        }
                }
            }
        }
    } else if (warp == kWarpMma && rank != 0) {
        // ======================= peer: relay "my half of the chunk has landed" to the leader =======================
        // (a bulk copy cannot complete on the other CTA's mbarrier: tools/pair_probe.py mode 1 never completes)
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s) {
                    const TcStep& st = p.steps[s];
                    const int n = ((two && st.n_chunks > kPStages) ? 2 : 1) * st.n_chunks;
                    for (int c = 0; c < n; ++c, ++cnt) {
                        const uint32_t stage = cnt % kPStages;
                        mbar_wait_spin(&bars->w_full[stage], (cnt / kPStages) & 1);
                        mbar_arrive_cluster(cluster_addr(&bars->w_full[stage], 0));
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ======================= MMA issuer (leader CTA, one elected lane) =======================
        // The pair MMA queue is deep, so barrier latencies hide behind queued work as long as this warp's own instruction
        // stream stays short: the whole warp walks the job list converged (uniform datapath, no per-instruction election
        // loops), descriptors are 32-bit low words advanced by adds, and the barrier of the NEXT chunk / job is probed
        // (non-blocking) right after the current MMAs are issued; the blocking wait is only the fallback.
        {
            uint32_t stage = 0, wpar = 0;                      // ring position and the parity of its current phase
            uint32_t eu = 0;                                   // uses of the encoding panel so far
            const uint32_t ring_lo = desc_lo_kmajor(smem_u32(smem + kOffRing));
            const uint32_t h_lo = desc_lo_kmajor(smem_u32(smem + kOffH)), e_lo = desc_lo_kmajor(smem_u32(smem + kOffE));
            bool w_ok = false, t_ok = false;
            // release: this is the last tile using the step's weight chunks; landed: the chunks were seen by the other tile
            auto issue_job = [&](const int x, const int g, const int s, const bool release, const bool landed) {
                const TcStep& st = p.steps[s];
                const uint32_t jx = (uint32_t)(g * p.n_steps + s);
                const uint32_t d_tmem = tmem + x * 256;
                const uint32_t idesc = umma_idesc(256, st.n_rows, false, false);
                const bool tr = kTrace && p.trace && blockIdx.x == 0 && g == 1 && lane == 0;
                // the epilogue of the slot's previous job has drained the accumulator and written this job's A panels
                if (jx > 0 && !t_ok) mbar_wait(&bars->tile_ready[x], (jx - 1) & 1);
                if (tr) p.trace[(x * 16 + s) * 16 + 0] = clock64();
                const int nc = st.n_chunks;
                for (int c = 0; c < nc; ++c) {
                    const int pn = st.panel[c], nk = st.ksteps[c];
                    if (!landed && !w_ok) mbar_wait(&bars->w_full[stage], wpar);
                    if (pn == kPanelE) mbar_wait(&bars->enc_ready, eu & 1);
                    tc_fence_after();
                    if (tr && c < 5) p.trace[(x * 16 + s) * 16 + 4 + c] = clock64();
                    const uint32_t a_lo = pn == kPanelE ? e_lo : h_lo + x * (65536 >> 4) + pn * (kPanelBytes >> 4);
                    const uint32_t b_lo = ring_lo + stage * (kPStageBytes >> 4);
                    uint64_t* done = &bars->w_empty[stage];
                    if (elect_one()) {
                        umma2_lo(d_tmem, a_lo, b_lo, idesc, c != 0);
                        if (nk == 4) {
                            umma2_lo(d_tmem, a_lo + 2, b_lo + 2, idesc, 1);
                            umma2_lo(d_tmem, a_lo + 4, b_lo + 4, idesc, 1);
                            umma2_lo(d_tmem, a_lo + 6, b_lo + 6, idesc, 1);
                        } else {
                            for (int k = 1; k < nk; ++k) umma2_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, idesc, 1);
                        }
                        if (release) umma_commit2(done, 3);                          // frees the slot in both CTAs' rings
                        if (pn == kPanelE) umma_commit2(&bars->enc_free, 3);         // the encoders may rewrite the panel
                    }
                    __syncwarp();
                    if (pn == kPanelE) ++eu;
                    if (++stage == kPStages) { stage = 0; wpar ^= 1; }
                    if (!landed) w_ok = mbar_test_wait(&bars->w_full[stage], wpar);  // next chunk: probe now, blocking wait as fallback
                    if (tr && c < 5) p.trace[(x * 16 + s) * 16 + 9 + c] = clock64();
                }
                if (elect_one()) umma_commit2(&bars->acc_full[x], 3);
                __syncwarp();
                if (tr) p.trace[(x * 16 + s) * 16 + 1] = clock64();
            };
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s) {
                    const uint32_t jx = (uint32_t)(g * p.n_steps + s);
                    const bool shared = two && p.steps[s].n_chunks <= kPStages;
                    const uint32_t stage0 = stage, wpar0 = wpar;
                    issue_job(0, g, s, !shared, false);
                    if (two) {
                        t_ok = jx > 0 && mbar_test_wait(&bars->tile_ready[1], (jx - 1) & 1);
                        if (shared) { stage = stage0; wpar = wpar0; }
                        issue_job(1, g, s, true, shared);
                        if (shared) w_ok = mbar_test_wait(&bars->w_full[stage], wpar);
                    }
                    // next job: slot 0, index jx + 1 within the slot (also across the tile boundary)
                    t_ok = mbar_test_wait(&bars->tile_ready[0], jx & 1);
                }
            }
        }
    } else if (warp >= kWarpEpi0 && warp < kWarpEpi0 + kEpiWarps) {
        // ======================= epilogue =======================
        // warp (q, j): TMEM lanes / tile rows 32q..32q+31, the 64-column panel j of every step.  Four warps per SM
        // sub-partition hide each other's TMEM / shared-memory latencies; each works in 16-column units with the next
        // unit's TMEM load in flight (small register footprint: 23 warps fit the register file).  The epilogue is the
        // pacing stage of the chain, so the hidden-layer path is kept to the bare instruction count: one F2FP + one
        // HFMA2(.RELU) per value pair, store offsets precomputed, the step kind dispatched once per job.
        const int q = warp & 3, j = (warp - kWarpEpi0) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + j * 64;
        const uint32_t ready0 = cluster_addr(&bars->tile_ready[0], 0), ready1 = cluster_addr(&bars->tile_ready[1], 0);
        // chunk c of this thread's panel row lives at (panel + row_base) ^ (c << 4)  (128-byte swizzle)
        const uint32_t row_base = smem_u32(smem + kOffH) + (uint32_t)row * kRowBytes + (((uint32_t)row & 7u) << 4);
        auto job = [&](const int x, const int g, const int s) {
            const TcStep& st = p.steps[s];
            const int kind = st.kind;
            const uint32_t jx = (uint32_t)(g * p.n_steps + s);
            // fused evaluation: this row's depths and its ray's constants are requested before the wait for the accumulator
            float pf_z = 0.f, pf_zn = 0.f;
            float4 pf_rc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.seg != nullptr && kind == EPI_VIEW && j == 0) {
                const long long pt0 = (long long)tile_of(2 * g + x) * kTileRows + row;
                if (pt0 < p.n_points) {
                    const int ray = (int)((unsigned)pt0 / (unsigned)p.n_samples), k = (int)pt0 - ray * p.n_samples;
                    pf_z = __ldg(p.z + pt0);
                    pf_zn = k == p.n_samples - 1 ? (p.ndc ? 1.f : 1e10f) : __ldg(p.z + pt0 + 1);        // :433 / :438
                    pf_rc = __ldg(p.rayc + ray);
                }
            }
            // training: the sigma noise of this row (Philox + Box-Muller, ~200 instructions) is drawn while the warp would otherwise
            // wait for the accumulator, not on the critical path behind it
            float pf_noise = 0.f;
            if (j == 0 && (kind == EPI_VIEW || kind == EPI_RELU_HEAD4) && (p.noise != nullptr || p.use_rng)) {
                const long long pt0 = (long long)tile_of(2 * g + x) * kTileRows + row;
                if (pt0 < p.n_points)
                    pf_noise = p.noise ? __ldg(p.noise + pt0)
                                       : p.noise_std * rng_pick(rng_normal4(p.noise_rng, (unsigned long long)pt0 >> 2), (unsigned long long)pt0);
            }
            mbar_wait(&bars->acc_full[x], jx & 1);
            tc_fence_after();
            const bool tr = kTrace && p.trace && blockIdx.x == 0 && g == 1 && warp == kWarpEpi0 && lane == 0;
            if (tr) p.trace[(x * 16 + s) * 16 + 2] = clock64();
            const bool own = j * 64 < st.n_rows && !(dbg & 2);        // the step has this panel
            const bool writes_h = ((kind != EPI_VIEW) || save) && !(dbg & 1);
            const uint32_t acc_addr = lane_addr + x * 256;
            uint32_t dst_row = row_base + x * 65536 + j * kPanelBytes;
            asm volatile("" : "+r"(dst_row));       // keep the address in its register: ptxas otherwise re-derives it from the row per store
            if (kind == EPI_RELU) {
                // ---- hidden layers and the feature layer: bias (+ReLU) in packed bf16 ----
                if (own) {
                    uint32_t rr[2][16];
                    uint32_t mw[2] = {0u, 0u};
                    tmem_ld16_issue(acc_addr, rr[0]);
                    const uint4* bb = reinterpret_cast<const uint4*>(s_bias16 + st.bias_row * 256 + j * 64);
                    constexpr bool relu = true;
                    if (save && jx > 0) mbar_wait(&bars->stash_done[x], (jx - 1) & 1);   // the slot's panels have been copied out
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        tmem_ld_wait16(rr[u & 1]);
                        if (u + 1 < 4) tmem_ld16_issue(acc_addr + (u + 1) * 16, rr[(u + 1) & 1]);
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint4 b = bb[2 * u + i];
                            pk[4 * i + 0] = bias_act_bf16x2(__uint_as_float(rr[u & 1][8 * i + 0]), __uint_as_float(rr[u & 1][8 * i + 1]), b.x, relu);
                            pk[4 * i + 1] = bias_act_bf16x2(__uint_as_float(rr[u & 1][8 * i + 2]), __uint_as_float(rr[u & 1][8 * i + 3]), b.y, relu);
                            pk[4 * i + 2] = bias_act_bf16x2(__uint_as_float(rr[u & 1][8 * i + 4]), __uint_as_float(rr[u & 1][8 * i + 5]), b.z, relu);
                            pk[4 * i + 3] = bias_act_bf16x2(__uint_as_float(rr[u & 1][8 * i + 6]), __uint_as_float(rr[u & 1][8 * i + 7]), b.w, relu);
                        }
                        if (writes_h) {
                            sts128(dst_row ^ ((2 * u) << 4), pk[0], pk[1], pk[2], pk[3]);
                            sts128(dst_row ^ ((2 * u + 1) << 4), pk[4], pk[5], pk[6], pk[7]);
                        }
                        if (save && relu) mw[u >> 1] |= relu_bits_unit(pk, u);
                        if (tr) p.trace[512 + (x * 16 + s) * 8 + 1 + u] = clock64();
                    }
                    if (writes_h) fence_async_smem();   // generic-proxy stores -> async proxy (MMA operand fetch, bulk store)
                    if (save && relu) {
                        const int tile = tile_of(2 * g + x);
                        if (tile < p.n_tiles)
                            *reinterpret_cast<uint2*>(p.bits + (size_t)tile * kBitsTileBytes + (size_t)st.slot * kBitsSlotBytes + row * 32 + j * 8) = make_uint2(mw[0], mw[1]);
                    }
                }
                tc_fence_before();     // TMEM reads ordered before the next MMA into this accumulator
                if (tr) p.trace[512 + (x * 16 + s) * 8 + 5] = clock64();
                __syncwarp();
                if (lane == 0) {
                    if (save) mbar_arrive(&bars->stash_ready[x]);
                    mbar_arrive_cluster(x ? ready1 : ready0);
                    if (tr) p.trace[(x * 16 + s) * 16 + 3] = clock64();
                }
                return;
            }
            if (kind == EPI_VIEW) {
                // ---- view layer (128 columns; column 128 = sigma_pre) + rgb head: fp32 on the un-rounded activations.  The 128 columns are split over all
                // 16 warps (32 each: half a panel), since only the rgb head and -- in training -- the stash consume them ----
                const int tile = tile_of(2 * g + x);
                const long long pt = (long long)tile * kTileRows + row;
                const bool valid = pt < p.n_points;
                const int pan = j >> 1, u0 = 2 * (j & 1);
                float head[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t sig_raw = 0u;         // sigma_pre of this thread's row: accumulator column kSigmaCol (the head row rides this step)
                if (!(dbg & 2)) {
                    const uint32_t vaddr = tmem + ((uint32_t)(q * 32) << 16) + x * 256 + pan * 64 + u0 * 16;
                    uint32_t rr[2][16];
                    tmem_ld16_issue(vaddr, rr[0]);
                    tmem_ld16_issue(vaddr + 16, rr[1]);
                    if (j == 0) tmem_ld1_issue(tmem + ((uint32_t)(q * 32) << 16) + x * 256 + kSigmaCol, sig_raw);
                    if (save && jx > 0) mbar_wait(&bars->stash_done[x], (jx - 1) & 1);
                    const int ray = valid ? (int)((unsigned)pt / (unsigned)p.n_samples) : 0;   // host guarantees n_points < 2^31
                    const float* vbias = p.view_bias + (size_t)ray * 128 + pan * 64 + u0 * 16;
                    uint8_t* vdst = smem + kOffH + x * 65536 + pan * kPanelBytes + (uint32_t)row * kRowBytes;
                    const uint32_t rx = (uint32_t)row & 7u;
                    tmem_ld_wait16(rr[0]);
                    tmem_ld_wait16(rr[1]);
                    if (j == 0) tmem_ld_wait1(sig_raw);
                    if (p.vis_pre != nullptr && valid) {
                        // visibility head: the point part of the view layer (shared by every view, :691-695) leaves in bf16;
                        // this thread holds columns 32 j .. 32 j + 31 of its row
                        uint4* vp = reinterpret_cast<uint4*>(p.vis_pre + (size_t)pt * 256 + j * 64);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            vp[2 * k] = make_uint4(pack_bf16(__uint_as_float(rr[k][0]), __uint_as_float(rr[k][1])), pack_bf16(__uint_as_float(rr[k][2]), __uint_as_float(rr[k][3])),
                                                   pack_bf16(__uint_as_float(rr[k][4]), __uint_as_float(rr[k][5])), pack_bf16(__uint_as_float(rr[k][6]), __uint_as_float(rr[k][7])));
                            vp[2 * k + 1] = make_uint4(pack_bf16(__uint_as_float(rr[k][8]), __uint_as_float(rr[k][9])), pack_bf16(__uint_as_float(rr[k][10]), __uint_as_float(rr[k][11])),
                                                       pack_bf16(__uint_as_float(rr[k][12]), __uint_as_float(rr[k][13])), pack_bf16(__uint_as_float(rr[k][14]), __uint_as_float(rr[k][15])));
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int col0 = pan * 64 + (u0 + k) * 16;
                        float v[16];
                        const float4* bb = reinterpret_cast<const float4*>(vbias + k * 16);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 b = __ldg(bb + i);
                            v[4 * i + 0] = fmaxf(__uint_as_float(rr[k][4 * i + 0]) + b.x, 0.f);
                            v[4 * i + 1] = fmaxf(__uint_as_float(rr[k][4 * i + 1]) + b.y, 0.f);
                            v[4 * i + 2] = fmaxf(__uint_as_float(rr[k][4 * i + 2]) + b.z, 0.f);
                            v[4 * i + 3] = fmaxf(__uint_as_float(rr[k][4 * i + 3]) + b.w, 0.f);
                        }
#pragma unroll
                        for (int hh = 0; hh < 3; ++hh) {
                            const float4* w = reinterpret_cast<const float4*>(s_wrgb + hh * 128 + col0);
                            float a = head[hh];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float4 ww = w[i];
                                a = fmaf(v[4 * i], ww.x, fmaf(v[4 * i + 1], ww.y, fmaf(v[4 * i + 2], ww.z, fmaf(v[4 * i + 3], ww.w, a))));
                            }
                            head[hh] = a;
                        }
                        if (save) {      // hv goes to the stash only (it feeds no later layer)
                            const uint32_t c0 = (uint32_t)(2 * (u0 + k));
                            *reinterpret_cast<uint4*>(vdst + ((c0 ^ rx) << 4)) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                            *reinterpret_cast<uint4*>(vdst + (((c0 + 1) ^ rx) << 4)) = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
                        }
                    }
                    if (save) fence_async_smem();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (save) mbar_arrive(&bars->stash_ready[x]);
                    mbar_arrive_cluster(x ? ready1 : ready0);
                    if (tr) p.trace[(x * 16 + s) * 16 + 3] = clock64();
                }
                if (j > 0) {
#pragma unroll
                    for (int h = 0; h < 4; ++h) s_part[((j - 1) * 128 + row) * 4 + h] = head[h];
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                float col[3] = {0.f, 0.f, 0.f}, sg = 0.f;
                if (j == 0 && valid) {
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) {
                        const float4 o = *reinterpret_cast<const float4*>(s_part + (jj * 128 + row) * 4);
                        head[0] += o.x; head[1] += o.y; head[2] += o.z;
                    }
#pragma unroll
                    for (int h = 0; h < 3; ++h) col[h] = sigmoid_acc(head[h] + s_misc[4 + h]);                    // :704-707
                    if (p.rgb) {
#pragma unroll
                        for (int h = 0; h < 3; ++h) p.rgb[pt * 3 + h] = col[h];
                    }
                    sg = fmaxf(__uint_as_float(sig_raw) + s_misc[0] + pf_noise, 0.f);                            // :665-672
                    if (p.sigma) p.sigma[pt] = sg;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // the partial sums may be overwritten by the next head step
                if (p.seg != nullptr && j == 0) {
                    // ---- fused evaluation: volume_rendering :430-483 on this warp's 32 consecutive samples of one ray ----
                    // (same arithmetic as composite.cu: delta :435-441, alpha :446, transmittance :447, NDC depth :495-501)
                    const float zz = pf_z;
                    const float delta = (pf_zn - zz) * pf_rc.x;
                    float zm = zz;
                    if (p.ndc) {
                        const float guard = (zz == 1.f) ? 1e-3f : 0.f;
                        zm = pf_rc.z * (__frcp_rn(1.f - zz + guard) - 1.f) + pf_rc.y;
                    }
                    const float alpha = valid ? 1.f - __expf(-sg * delta) : 0.f;
                    float incl = valid ? (1.f - alpha) + 1e-10f : 1.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl *= v;
                    }
                    float tr_in = __shfl_up_sync(0xffffffffu, incl, 1);
                    if (lane == 0) tr_in = 1.f;
                    const float wl = alpha * tr_in;
                    const float zm_ref = __shfl_sync(0xffffffffu, zm, 0), zn_ref = __shfl_sync(0xffffffffu, zz, 0);
                    const float em = zm - zm_ref, en = zz - zn_ref;
                    // eight sums over the warp in nine shuffles: every step halves the values a lane still carries
                    float v8[8] = {wl, wl * col[0], wl * col[1], wl * col[2], wl * em, wl * en, wl * em * em, wl * en * en};
                    float v4[4], v2[2];
                    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float got = __shfl_xor_sync(0xffffffffu, h16 ? v8[i] : v8[i + 4], 16);
                        v4[i] = (h16 ? v8[i + 4] : v8[i]) + got;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float got = __shfl_xor_sync(0xffffffffu, h8 ? v4[i] : v4[i + 2], 8);
                        v2[i] = (h8 ? v4[i + 2] : v4[i]) + got;
                    }
                    float v1 = (h4 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? v2[0] : v2[1], 4);
                    v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
                    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);                // lane l: sum number (l >> 2)
                    if (valid) {
                        float* rec = p.seg + (size_t)((unsigned long long)pt >> 5) * kSegFloats;
                        if ((lane & 3) == 0) rec[SEG_A + (lane >> 2)] = v1;
                        if (lane == 31) rec[SEG_P] = incl;
                        if (lane == 1) rec[SEG_ZM] = zm_ref;
                        if (lane == 2) rec[SEG_ZN] = zn_ref;
                        if (p.alpha_out) p.alpha_out[pt] = alpha;
                        if (p.wloc) p.wloc[pt] = wl;
                    }
                }
                return;
            }
            // ---- last trunk layer of an MLP without a view branch, with its sigma + rgb head (4 rows): fp32 on the un-rounded
            // activations.  (With a view branch the last trunk layer is a plain EPI_RELU step and sigma rides the view step.) ----
            const int tile = tile_of(2 * g + x);
            const long long pt = (long long)tile * kTileRows + row;
            const bool valid = pt < p.n_points;
            float head[4] = {0.f, 0.f, 0.f, 0.f};
            auto head_units = [&](auto nh_tag) {
                constexpr int NH = decltype(nh_tag)::value;
                uint32_t rr[2][16];
                tmem_ld16_issue(acc_addr, rr[0]);
                if (save && jx > 0) mbar_wait(&bars->stash_done[x], (jx - 1) & 1);
                uint32_t mw[2] = {0u, 0u};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col0 = j * 64 + u * 16;
                    tmem_ld_wait16(rr[u & 1]);
                    if (u + 1 < 4) tmem_ld16_issue(acc_addr + (u + 1) * 16, rr[(u + 1) & 1]);
                    float v[16];
                    const float4* bb = reinterpret_cast<const float4*>(s_bias32 + col0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b = bb[i];
                        v[4 * i + 0] = fmaxf(__uint_as_float(rr[u & 1][4 * i + 0]) + b.x, 0.f);
                        v[4 * i + 1] = fmaxf(__uint_as_float(rr[u & 1][4 * i + 1]) + b.y, 0.f);
                        v[4 * i + 2] = fmaxf(__uint_as_float(rr[u & 1][4 * i + 2]) + b.z, 0.f);
                        v[4 * i + 3] = fmaxf(__uint_as_float(rr[u & 1][4 * i + 3]) + b.w, 0.f);
                    }
#pragma unroll
                    for (int hh = 0; hh < NH; ++hh) {
                        const float4* w = reinterpret_cast<const float4*>(s_whead + hh * 256 + col0);
                        float a = head[hh];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 ww = w[i];
                            a = fmaf(v[4 * i], ww.x, fmaf(v[4 * i + 1], ww.y, fmaf(v[4 * i + 2], ww.z, fmaf(v[4 * i + 3], ww.w, a))));
                        }
                        head[hh] = a;
                    }
                    if (writes_h) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
                        sts128(dst_row ^ ((2 * u) << 4), pk[0], pk[1], pk[2], pk[3]);
                        sts128(dst_row ^ ((2 * u + 1) << 4), pk[4], pk[5], pk[6], pk[7]);
                        if (save) mw[u >> 1] |= relu_bits_unit(pk, u);
                    }
                }
                if (writes_h) fence_async_smem();
                if (save && tile < p.n_tiles)
                    *reinterpret_cast<uint2*>(p.bits + (size_t)tile * kBitsTileBytes + (size_t)st.slot * kBitsSlotBytes + row * 32 + j * 8) = make_uint2(mw[0], mw[1]);
            };
            if (own) head_units(std::integral_constant<int, 4>{});
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (save) mbar_arrive(&bars->stash_ready[x]);
                mbar_arrive_cluster(x ? ready1 : ready0);
                if (tr) p.trace[(x * 16 + s) * 16 + 3] = clock64();
            }
            // combine the panels: warps 1-3 of the row group hand their partial sums to warp 0
            if (j > 0) {
#pragma unroll
                for (int h = 0; h < 4; ++h) s_part[((j - 1) * 128 + row) * 4 + h] = head[h];
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
            if (j == 0 && valid) {
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const float4 o = *reinterpret_cast<const float4*>(s_part + (jj * 128 + row) * 4);
                    head[0] += o.x; head[1] += o.y; head[2] += o.z; head[3] += o.w;
                }
                head[0] += s_misc[0]; head[1] += s_misc[1]; head[2] += s_misc[2]; head[3] += s_misc[3];
                {
                    const float sg = fmaxf(head[0] + pf_noise, 0.f);                               // :668-672
                    if (p.sigma) p.sigma[pt] = sg;
#pragma unroll
                    for (int h = 0; h < 3; ++h) p.rgb[pt * 3 + h] = sigmoid_acc(head[1 + h]);     // :676-680
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // the partial sums may be overwritten by the next head step
        };
        for (int g = 0; 2 * g < my_super; ++g) {
            const bool two = 2 * g + 1 < my_super;
            for (int s = 0; s < p.n_steps; ++s) {
                job(0, g, s);
                if (two) job(1, g, s);
            }
        }
    } else if (warp < 4) {
        // ======================= encoder =======================
        // The encodings of the group's two tiles are computed once and kept in registers (packed bf16); the single
        // encoding panel is rewritten before each use (layer 0, skip layer, view layer of the points-augmented model) in
        // the issuer's order, as soon as the MMAs of the previous use have completed.
        const int row = warp * 32 + lane;
        const uint32_t enc_addr = cluster_addr(&bars->enc_ready, 0);
        uint32_t e_steps = 0;                       // bit s: step s reads the encoding panel
        for (int s = 0; s < p.n_steps; ++s)
            for (int c = 0; c < p.steps[s].n_chunks; ++c)
                if (p.steps[s].panel[c] == kPanelE) e_steps |= 1u << s;
        uint32_t eu = 0;
        for (int g = 0; 2 * g < my_super; ++g) {
            const bool two = 2 * g + 1 < my_super;
            uint32_t pk[2][32];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const long long pt = (long long)tile_of(2 * g + x) * kTileRows + row;
                float xx[3] = {0.f, 0.f, 0.f};
                if ((x == 0 || two) && pt < p.n_points) {
                    const int ray = (int)((unsigned)pt / (unsigned)p.n_samples);
                    const float zz = p.z[pt];
#pragma unroll
                    for (int c = 0; c < 3; ++c) xx[c] = fmaf(p.rays_d[ray * 3 + c], zz, p.rays_o[ray * 3 + c]);   // :140/:142
                }
                float enc[64];
#pragma unroll
                for (int k = 0; k < 64; ++k) enc[k] = 0.f;
                encode_point(xx, p.pts_degree, enc);
#pragma unroll
                for (int k = 0; k < 32; ++k) pk[x][k] = pack_bf16(enc[2 * k], enc[2 * k + 1]);
            }
            for (int s = 0; s < p.n_steps; ++s) {
                if (!((e_steps >> s) & 1u)) continue;
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    if (x == 1 && !two) continue;
                    if (eu > 0) mbar_wait(&bars->enc_free, (eu - 1) & 1);
                    uint8_t* dst = smem + kOffE;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(dst + swz_offset(row, c)) = make_uint4(pk[x][4 * c], pk[x][4 * c + 1], pk[x][4 * c + 2], pk[x][4 * c + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(enc_addr);
                    ++eu;
                }
            }
        }
    } else if (warp == kWarpStash) {
        // ======================= stash writer (training) =======================
        // One SM moves at most ~32 bytes per clock to global memory (tools/store_probe.py: 31.6 B/clk with two stores in
        // flight, 25.8 when every store is waited for before the next is issued), and a training forward needs 64 KB per job =
        // ~22 B/clk at the pace of the eval kernel: the store path is the resource to keep busy.  The writer therefore keeps
        // TWO bulk stores in flight: the slot's panels are released (stash_done) when the NEXT store has been issued and the
        // previous one has finished reading shared memory.  The two slots alternate, so the released slot is never the one
        // the next store needs; with a single active slot (last, odd group) every store is waited for at once.
        if (save && lane == 0) {
            int pending = -1;                                  // slot whose store may still be reading its panels
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s)
                    for (int x = 0; x < (two ? 2 : 1); ++x) {
                        const TcStep& st = p.steps[s];
                        const uint32_t jx = (uint32_t)(g * p.n_steps + s);
                        const int tile = tile_of(2 * g + x);
                        mbar_wait_sleep(&bars->stash_ready[x], jx & 1, 64);
                        if (tile < p.n_tiles && p.stash_pieces > 1) {
                            // Paced: one 16 KB panel per request and at most two requests queued.  The weight chunks come through the same
                            // bulk-copy unit of the SM: behind two 64 KB stores a chunk arrived up to ~2 000 cycles late (clock64 trace of the
                            // issuer's waits, profiles/r2_weight_loader.md); behind 32 KB it does not.  Forward -6.5 %, dgrad -6.6 %.
                            const uint32_t total = (uint32_t)(st.n_rows / 64) * kPanelBytes, piece = 65536u / (uint32_t)p.stash_pieces;
                            for (uint32_t off = 0; off < total; off += piece) {
                                bulk_s2g(p.stash + (size_t)tile * p.tile_stash_bytes + (size_t)st.slot * 65536 + off, smem + kOffH + x * 65536 + off, piece);
                                bulk_commit();
                                bulk_wait_read<1>();
                                if (off == 0 && pending >= 0) { mbar_arrive(&bars->stash_done[pending]); pending = -1; }
                            }
                            if (two) pending = x;
                            else { bulk_wait_read<0>(); mbar_arrive(&bars->stash_done[x]); }
                        } else if (tile < p.n_tiles) {
                            bulk_s2g(p.stash + (size_t)tile * p.tile_stash_bytes + (size_t)st.slot * 65536, smem + kOffH + x * 65536,
                                     (uint32_t)(st.n_rows / 64) * kPanelBytes);
                            bulk_commit();
                            if (pending >= 0) { bulk_wait_read<1>(); mbar_arrive(&bars->stash_done[pending]); pending = -1; }
                            if (two) pending = x;
                            else { bulk_wait_read<0>(); mbar_arrive(&bars->stash_done[x]); }
                        } else {
                            if (pending >= 0) { bulk_wait_read<0>(); mbar_arrive(&bars->stash_done[pending]); pending = -1; }
                            mbar_arrive(&bars->stash_done[x]);
                        }
                    }
            }
            if (pending >= 0) { bulk_wait_read<0>(); mbar_arrive(&bars->stash_done[pending]); }
            bulk_wait_all<0>();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while the pair may still read its smem, signal its barriers or use its TMEM
    if (warp == kWarpMma) tmem_dealloc2<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// workspace + drivers
// ------------------------------------------------------------------------------------------------
size_t tc_workspace_bytes(const MlpDims& m, const snerf_mlp_desc& d, int n_rays, int n_samples, uint32_t flags) {
    return tc_ws_layout(m, build_plan(d, nullptr), n_rays, n_samples, flags).total;
}

static long long* g_trace = nullptr;   // set by snerfdbg_set_trace (debug only)
static int g_fwd_debug = 0;

int tc_forward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* rays_o, const float* rays_d,
               const float* view_dirs, const float* z, const float* noise, float* sigma, float* rgb, void* ws, size_t ws_bytes,
               int n_rays, int n_samples, uint32_t flags, cudaStream_t st, const unsigned long long* rng_seed_offset, float noise_std,
               const FusedRun* fused) {
    const MlpDims m(d);
    const TcPlan pl = build_plan(d, prm);
    const TcWorkspace w = tc_ws_layout(m, pl, n_rays, n_samples, flags);
    SNERF_REQUIRE(ws_bytes >= w.total, "mlp_forward: workspace too small (%zu < %zu)", ws_bytes, w.total);
    SNERF_REQUIRE(((uintptr_t)packed & 15) == 0 && ((uintptr_t)ws & 15) == 0, "mlp_forward: packed/workspace must be 16-byte aligned");
    uint8_t* wsb = (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    if (m.has_view) {
        tc_view_bias_kernel<<<(n_rays + kVbRays - 1) / kVbRays, 128, 0, st>>>(view_dirs, prm[SNERF_P_VIEW_W], prm[SNERF_P_VIEW_B],
                                                    (const float*)((const uint8_t*)packed + align_up(pl.packed_bytes, 1024)) + 128 * 256,
                                                    (float*)(wsb + w.view_bias), n_rays, d.view_degree, m.view_in, m.width + m.enc_hi, m.venc,
                                                    rays_d, fused ? fused->cam_o : nullptr, fused ? fused->cam_d : nullptr,
                                                    fused ? (float4*)(wsb + w.view_bias + (size_t)n_rays * 128 * sizeof(float)) : nullptr,
                                                    (flags & SNERF_FLAG_VIS_HEAD) ? (float*)(wsb + w.view_enc) : nullptr);
        SNERF_LAUNCH_OK("tc_view_bias_kernel");
    }
    FwdParams p{};
    p.packed = (const uint8_t*)packed;
    for (int l = 0; l < 8; ++l) p.bias[l] = prm[2 * l + 1];
    p.bias[8] = nullptr;          // (the feature bias reaches the view layer through the per-ray bias table)
    p.w_head = prm[SNERF_P_HEAD_W]; p.b_head = prm[SNERF_P_HEAD_B];
    p.w_rgb = m.has_view ? prm[SNERF_P_RGB_W] : nullptr; p.b_rgb = m.has_view ? prm[SNERF_P_RGB_B] : nullptr;
    p.view_bias = (const float*)(wsb + w.view_bias);
    p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.noise = noise; p.sigma = sigma; p.rgb = rgb;
    if (rng_seed_offset != nullptr && noise == nullptr && noise_std != 0.f) {
        p.noise_rng = RngKey{rng_seed_offset[0], rng_seed_offset[1]};
        p.noise_std = noise_std;
        p.use_rng = 1;
    }
    p.stash = (flags & SNERF_FLAG_SAVE_FOR_BWD) ? wsb + w.act : nullptr;
    p.bits = (flags & SNERF_FLAG_SAVE_FOR_BWD) ? wsb + w.bits : nullptr;
    p.vis_pre = ((flags & SNERF_FLAG_VIS_HEAD) && m.has_view) ? wsb + w.vis_pre : nullptr;
    if (fused) {
        p.seg = fused->seg; p.alpha_out = fused->alpha; p.wloc = fused->wloc; p.ndc = fused->cam_o != nullptr ? 1 : 0;
        p.rayc = (const float4*)(wsb + w.view_bias + (size_t)n_rays * 128 * sizeof(float));
    }
    {
        static int pieces = -1;
        if (pieces < 0) { const char* e = getenv("SNERF_STASH_PIECES"); pieces = e ? atoi(e) : 4; }
        p.stash_pieces = (pieces == 2 || pieces == 4 || pieces == 8 || pieces == 16) ? pieces : 1;
    }
    p.trace = g_trace; p.debug = g_fwd_debug;
    p.n_points = (long long)n_rays * n_samples;
    p.n_samples = n_samples; p.n_tiles = w.n_tiles; p.n_steps = pl.n_fwd; p.pts_degree = d.pts_degree; p.head_out = m.head_out;
    p.tile_stash_bytes = pl.tile_stash_bytes;
    for (int s = 0; s < pl.n_fwd; ++s) p.steps[s] = pl.fwd[s];
    SNERF_REQUIRE(p.n_points < (1LL << 31), "mlp_forward: more than 2^31 points in one call");
    static bool attr = false;
    if (!attr) {
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_forward_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
        attr = true;
    }
    const int grid = pair_grid(w.n_tiles);
    if (p.trace) {
        if (p.stash) SNERF_CUDA_OK(launch_clustered(tc_forward_kernel<true, true>, grid, kFwdThreads, kFwdSmem, st, p));
        else SNERF_CUDA_OK(launch_clustered(tc_forward_kernel<true, false>, grid, kFwdThreads, kFwdSmem, st, p));
    } else {
        if (p.stash) SNERF_CUDA_OK(launch_clustered(tc_forward_kernel<false, true>, grid, kFwdThreads, kFwdSmem, st, p));
        else SNERF_CUDA_OK(launch_clustered(tc_forward_kernel<false, false>, grid, kFwdThreads, kFwdSmem, st, p));
    }
    return SNERF_OK;
}

// ------------------------------------------------------------------------------------------------
// fused evaluation (row X1): fold of the per-run records into the per-ray maps, one warp per ray
// ------------------------------------------------------------------------------------------------
// Run k of a ray starts with transmittance T_k = prod_{j<k} P_j (:447); with w = T_k w':
//   acc = sum T_k A_k, rgb = sum T_k C_k, depth = sum T_k D_k / (acc + 1e-6)   (:449-459)
//   depth_var = sum_k T_k (S2_k - 2 S1_k (depth - z_ref,k) + A_k (depth - z_ref,k)^2)   (:454, :460; D_k = S1_k + A_k z_ref,k)
// and the weights of the run's samples are T_k times the run-relative weights the MLP kernel left in `wloc`.
constexpr int kFoldWarps = 8;
__global__ void __launch_bounds__(kFoldWarps * 32) composite_fold_kernel(const float* __restrict__ seg, const float* __restrict__ wloc,
                                                                         float* __restrict__ rgb_map, float* __restrict__ acc_out,
                                                                         float* __restrict__ depth, float* __restrict__ depth_var,
                                                                         float* __restrict__ depth_ndc, float* __restrict__ depth_var_ndc,
                                                                         float* __restrict__ weights, int n_rays, int n_samples, int white) {
    const int ray = blockIdx.x * kFoldWarps + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (ray >= n_rays) return;
    const int n_seg = n_samples / 32;                      // <= 32 (host check)
    float r[kSegFloats];
#pragma unroll
    for (int i = 0; i < kSegFloats; ++i) r[i] = 0.f;
    r[SEG_P] = 1.f;
    if (lane < n_seg) {
        const float4* src = reinterpret_cast<const float4*>(seg + ((size_t)ray * n_seg + lane) * kSegFloats);
        const float4 a = src[0], b = src[1], c = src[2];
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w; r[8] = c.x; r[9] = c.y; r[10] = c.z;
    }
    float incl = r[SEG_P];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl *= v;
    }
    float t_in = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) t_in = 1.f;
    auto wsum = [](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    const float A = r[SEG_A];
    const float acc = wsum(t_in * A);
    const float cr = wsum(t_in * r[SEG_CR]), cg = wsum(t_in * r[SEG_CG]), cb = wsum(t_in * r[SEG_CB]);
    const float inv = 1.f / (acc + 1e-6f);
    const float d_m = wsum(t_in * (r[SEG_S1M] + A * r[SEG_ZM])) * inv, d_n = wsum(t_in * (r[SEG_S1N] + A * r[SEG_ZN])) * inv;
    // run k about the ray's depth: sum w' (z - depth)^2 = S2 - 2 S1 (depth - z_ref) + A (depth - z_ref)^2
    const float gm = d_m - r[SEG_ZM], gn = d_n - r[SEG_ZN];
    const float v_m = wsum(t_in * (r[SEG_S2M] - 2.f * r[SEG_S1M] * gm + A * gm * gm));
    const float v_n = wsum(t_in * (r[SEG_S2N] - 2.f * r[SEG_S1N] * gn + A * gn * gn));
    if (lane == 0) {
        const float bg = white ? 1.f - acc : 0.f;                                                   // :463
        rgb_map[(size_t)ray * 3 + 0] = cr + bg;
        rgb_map[(size_t)ray * 3 + 1] = cg + bg;
        rgb_map[(size_t)ray * 3 + 2] = cb + bg;
        acc_out[ray] = acc;
        depth[ray] = d_m;
        depth_var[ray] = v_m;
        if (depth_ndc) { depth_ndc[ray] = d_n; depth_var_ndc[ray] = v_n; }
    }
    if (weights) {
        for (int k = 0; k < n_seg; ++k) {
            const float t = __shfl_sync(0xffffffffu, t_in, k);
            const size_t o = (size_t)ray * n_samples + k * 32 + lane;
            weights[o] = t * wloc[o];                                                               // :448
        }
    }
}

size_t tc_render_workspace_bytes(const MlpDims& m, const snerf_mlp_desc& d, int n_rays, int n_samples) {
    const size_t P = (size_t)n_rays * n_samples;
    return tc_workspace_bytes(m, d, n_rays, n_samples, 0) + align_up(P / 32 * kSegFloats * sizeof(float), 1024) +
           align_up(P * sizeof(float), 1024) + 1024;
}

int tc_render_forward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* pts_o,
                      const float* pts_d, const float* view_dirs, const float* z, const FusedComposite& fc, void* ws,
                      size_t ws_bytes, int n_rays, int n_samples, cudaStream_t st) {
    const MlpDims m(d);
    if (!m.has_view || n_samples % 32 != 0 || n_samples > 1024)
        return fail(SNERF_ERR_UNSUPPORTED, "snerf_render_forward: needs an MLP with a view branch and n_samples %% 32 == 0 (<= 1024), got %d", n_samples);
    SNERF_REQUIRE(ws_bytes >= tc_render_workspace_bytes(m, d, n_rays, n_samples), "snerf_render_forward: workspace too small");
    const size_t P = (size_t)n_rays * n_samples;
    const size_t base = tc_workspace_bytes(m, d, n_rays, n_samples, 0);
    uint8_t* wsb = (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    FusedRun fr{};
    fr.seg = (float*)(wsb + align_up(base, 1024));
    fr.wloc = fc.weights ? (float*)((uint8_t*)fr.seg + align_up(P / 32 * kSegFloats * sizeof(float), 1024)) : nullptr;
    fr.alpha = fc.alpha;
    fr.cam_o = fc.ndc ? fc.rays_o : nullptr;
    fr.cam_d = fc.ndc ? fc.rays_d : nullptr;
    const int rc = tc_forward(d, prm, packed, pts_o, pts_d, view_dirs, z, nullptr, nullptr, nullptr, ws, base, n_rays, n_samples, 0, st,
                              nullptr, 0.f, &fr);
    if (rc != SNERF_OK) return rc;
    composite_fold_kernel<<<(n_rays + kFoldWarps - 1) / kFoldWarps, kFoldWarps * 32, 0, st>>>(
        fr.seg, fr.wloc, fc.rgb_map, fc.acc, fc.depth, fc.depth_var, fc.ndc ? fc.depth_ndc : nullptr, fc.ndc ? fc.depth_var_ndc : nullptr,
        fc.weights, n_rays, n_samples, fc.white ? 1 : 0);
    SNERF_LAUNCH_OK("composite_fold_kernel");
    return SNERF_OK;
}

#ifdef SNERF_DEBUG   // probe kernels and debug entry points live in libsimplenerf_b200_dbg.so only (build.py --debug)
// ------------------------------------------------------------------------------------------------
// Descriptor probe: runs a host-specified list of tcgen05.mma instructions on host-provided operand
// images and returns the accumulator (tools/tc_probe.py).
// ------------------------------------------------------------------------------------------------
struct ProbeOp { uint32_t a_off, b_off, d_col, accumulate; };
__device__ int g_probe_chunk = 0, g_probe_waits = 0;

__global__ void __launch_bounds__(128) tc_probe_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes,
                                                       const uint8_t* __restrict__ b_img, uint32_t b_bytes,
                                                       float* __restrict__ d_out, const ProbeOp* __restrict__ ops, int n_ops,
                                                       uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                                                       uint32_t idesc, uint64_t desc_bits, int n_cols, long long* timing) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_load, bar_mma, bar_aux;
    __shared__ uint32_t tmem_base_s;
    __shared__ ProbeOp s_ops[256];
    for (int i = threadIdx.x; i < n_ops && i < 256; i += blockDim.x) s_ops[i] = ops[i];
    uint8_t* sa = smem;
    uint8_t* sb = smem + 65536;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_aux, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_load, a_bytes + b_bytes);
        bulk_g2s(sa, a_img, a_bytes, &bar_load);
        bulk_g2s(sb, b_img, b_bytes, &bar_load);
        mbar_wait(&bar_load, 0);
        tc_fence_after();
        const int chunk = *(volatile int*)&g_probe_chunk, nwaits = *(volatile int*)&g_probe_waits;
        int next_commit = chunk > 0 ? chunk : 1 << 30;
        const long long t_start = clock64();
        for (int i = 0; i < n_ops; ++i) {
            const ProbeOp op = s_ops[i];
            const uint32_t aa = smem_u32(sa) + op.a_off, ba = smem_u32(sb) + op.b_off;
            const uint64_t ad = (uint64_t)((aa >> 4) & 0x3FFFu) | ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) |
                                ((uint64_t)((a_sbo >> 4) & 0x3FFFu) << 32) | desc_bits;
            const uint64_t bd = (uint64_t)((ba >> 4) & 0x3FFFu) | ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) |
                                ((uint64_t)((b_sbo >> 4) & 0x3FFFu) << 32) | desc_bits;
            umma(tmem_base + op.d_col, ad, bd, idesc, op.accumulate != 0);
            if (i + 1 == next_commit) {   // issue pattern of the chain kernels: commit + barrier probes per chunk
                next_commit += chunk;
                umma_commit(&bar_aux);
                for (int w = 0; w < nwaits; ++w) { mbar_wait(&bar_load, 0); tc_fence_after(); }
            }
        }
        const long long t_issued = clock64();
        umma_commit(&bar_mma);
        mbar_wait(&bar_mma, 0);
        const long long t_done = clock64();
        if (timing) { timing[0] = t_issued - t_start; timing[1] = t_done - t_start; }
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < n_cols; c += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
        float* row = d_out + (size_t)(warp * 32 + lane) * n_cols + c;
#pragma unroll
        for (int i = 0; i < 32; ++i) row[i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

#endif  // SNERF_DEBUG

int tc_selftest(float* host_max_err, cudaStream_t) {
    for (int i = 0; i < 4; ++i) host_max_err[i] = -1.f;
    return fail(SNERF_ERR_UNSUPPORTED, "use tools/tc_probe.py");
}

}  // namespace snerf

using namespace snerf;
extern "C" int snerf_has_tensor_path(void) { return 1; }

#ifdef SNERF_DEBUG
extern "C" void snerfdbg_set_trace(long long* device_buffer_512) { snerf::g_trace = device_buffer_512; }
extern "C" void snerfdbg_set_fwd_debug(int bits) { snerf::g_fwd_debug = bits; }
extern "C" void snerfdbg_set_probe_pattern(int chunk, int waits) {
    cudaMemcpyToSymbol(snerf::g_probe_chunk, &chunk, sizeof(int));
    cudaMemcpyToSymbol(snerf::g_probe_waits, &waits, sizeof(int));
}

// debug entry (not part of the public ABI): all pointers are device pointers
extern "C" int snerfdbg_probe(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, float* d_out,
                              const void* ops, int n_ops, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                              uint32_t idesc, uint64_t desc_bits, int n_cols, void* stream, long long* timing) {
    SNERF_REQUIRE(a_bytes <= 65536 && b_bytes <= 131072 && n_cols % 32 == 0 && n_cols <= 512 && n_ops <= 256, "probe: bad sizes");
    SNERF_CUDA_OK(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 131072));
    tc_probe_kernel<<<1, 128, 65536 + 131072, (cudaStream_t)stream>>>((const uint8_t*)a_img, a_bytes, (const uint8_t*)b_img,
                                                                    b_bytes, d_out, (const ProbeOp*)ops, n_ops, a_lbo, a_sbo,
                                                                    b_lbo, b_sbo, idesc, desc_bits, n_cols, timing);
    SNERF_LAUNCH_OK("tc_probe_kernel");
    return SNERF_OK;
}
#endif  // SNERF_DEBUG
