// bf16 tcgen05 MLP backward: (1) a fused dgrad chain kernel, (2) a persistent wgrad kernel.
//
// The forward kernel (mlp_tc.cu) left the bf16 activation panels of every layer in the HBM stash as raw
// swizzled smem images.  Backward of the reference MLP (src/models/SimpleNeRF01.py :626-715 under autograd):
//
//  dgrad kernel  -- same warp-specialised chain as the forward: per 128-point tile
//      prologue warps : d rgb_pre = d rgb * rgb (1-rgb), d sigma_pre = d sigma * [sigma>0],
//                       dY_v = (d rgb_pre W_rgb) * [hv>0]            (CUDA cores, fp32)
//      MMA chain      : d feature = dY_v W_view[:, :256];  d h8 = d feature W_feat + d head_pre W_head;
//                       d h_l = dY_l W_l[:, hidden]  for l = 7..1    (tcgen05, B = packed W^T chunks)
//      epilogue warps : ReLU mask from the stashed activation, bf16, back to smem as the next A operand
//      stash writer   : every dY panel -> HBM (raw panel image) for the wgrad kernel
//  wgrad kernel  -- per parameter matrix dW = dY^T X over all points: both operands are read MN-major from the
//      stashed panels (no transposes), accumulated in TMEM across all tiles of the CTA, flushed once with fp32
//      atomics.  Idle warps reduce the bias gradients (column sums of dY) and the tiny head matrices from the
//      same smem stages; the positional-encoding operands (layer 0, skip layer, view layer) are recomputed.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_plan.cuh"

namespace snerf {
using namespace tc;

// =================================================================================================
// dgrad chain kernel
// =================================================================================================
constexpr int kBwdThreads = 480;   // loader, MMA, 8 epilogue, 4 prologue, stash writer
constexpr int kBwdEpiWarps = 8;
constexpr uint32_t kBOffH = 0;
constexpr uint32_t kBOffP = 65536;                                  // 3 panels: dY_v (2) + head-pre (1)
constexpr uint32_t kBOffRing = kBOffP + 3 * kPanelBytes;             // 114688
constexpr uint32_t kBOffConst = kBOffRing + kStages * kStageBytes;   // 212992
constexpr uint32_t kBOffBars = kBOffConst + 2048;
constexpr uint32_t kBwdSmem = kBOffBars + 512 + 1024;

struct BwdParams {
    const uint8_t* packed;
    const uint8_t* act;
    uint8_t* dy;
    const float *sigma, *rgb, *d_sigma, *d_rgb, *w_rgb;
    long long n_points;
    int n_tiles, n_steps, has_view;
    uint32_t tile_stash_bytes;
    TcStep steps[kMaxSteps];
};

struct BwdBars {
    uint64_t w_full[kStages], w_empty[kStages], acc_full[2], panel_ready[4], panel_stored[4], pro_ready, pro_free, pro_stored;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kBwdThreads, 1) tc_dgrad_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    BwdBars* bars = (BwdBars*)(smem + kBOffBars);
    float* s_wrgb = (float*)(smem + kBOffConst);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], kCluster); }
        for (int i = 0; i < 2; ++i) mbar_init(&bars->acc_full[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&bars->panel_ready[i], kBwdEpiWarps * 32); mbar_init(&bars->panel_stored[i], 1); }
        mbar_init(&bars->pro_ready, 128);
        mbar_init(&bars->pro_free, 1);
        mbar_init(&bars->pro_stored, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
    if (p.w_rgb) for (int i = threadIdx.x; i < 3 * 128; i += kBwdThreads) s_wrgb[i] = p.w_rgb[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int my_tiles = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;   // uniform per cluster; tiles >= n_tiles are dummies
    constexpr uint16_t kClusterMask = (uint16_t)((1u << kCluster) - 1);

    if (warp == 0) {
        // ======================= weight loader (1/kCluster of every chunk, multicast to the cluster) =======================
        if (lane == 0) {
            const uint32_t rank = cluster_rank();
            uint32_t cnt = 0;
            for (int ti = 0; ti < my_tiles; ++ti)
                for (int s = 0; s < p.n_steps; ++s) {
                    const TcStep& st = p.steps[s];
                    const uint32_t bytes = (uint32_t)st.n_rows * kRowBytes;
                    const uint32_t slice = bytes / kCluster;
                    for (int c = 0; c < st.n_chunks; ++c, ++cnt) {
                        const uint32_t stage = cnt % kStages, round = cnt / kStages;
                        if (round > 0) mbar_wait(&bars->w_empty[stage], (round - 1) & 1);
                        mbar_arrive_expect_tx(&bars->w_full[stage], bytes);
                        bulk_g2s_multicast(smem + kBOffRing + stage * kStageBytes + rank * slice,
                                           p.packed + st.w_off + (uint32_t)c * bytes + rank * slice, slice,
                                           &bars->w_full[stage], kClusterMask);
                    }
                }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            uint32_t cnt = 0, it = 0;
            const uint32_t idesc = umma_idesc(128, 256, false, false);
            for (int ti = 0; ti < my_tiles; ++ti) {
                bool pro_waited = false;
                for (int s = 0; s < p.n_steps; ++s, ++it) {
                    const TcStep& st = p.steps[s];
                    const uint32_t d_tmem = tmem + (it & 1) * 256;
                    uint32_t waited = 0;
                    for (int c = 0; c < st.n_chunks; ++c, ++cnt) {
                        const int pn = st.panel[c];
                        uint32_t a_addr;
                        if (pn >= kPanelP) {
                            if (!pro_waited) { mbar_wait(&bars->pro_ready, ti & 1); pro_waited = true; }
                            a_addr = smem_u32(smem + kBOffP + (pn - kPanelP) * kPanelBytes);
                        } else {
                            if (it > 0 && !(waited & (1u << pn))) { mbar_wait(&bars->panel_ready[pn], (it - 1) & 1); waited |= 1u << pn; }
                            a_addr = smem_u32(smem + kBOffH + pn * kPanelBytes);
                        }
                        const uint32_t stage = cnt % kStages;
                        mbar_wait(&bars->w_full[stage], (cnt / kStages) & 1);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(smem + kBOffRing + stage * kStageBytes);
                        for (int k = 0; k < st.ksteps[c]; ++k)
                            umma(d_tmem, umma_desc_kmajor(a_addr, k), umma_desc_kmajor(b_addr, k), idesc, (c | k) != 0);
                        umma_commit_multicast(&bars->w_empty[stage], kClusterMask);
                    }
                    umma_commit(&bars->acc_full[it & 1]);
                    if (st.last_e_use) umma_commit(&bars->pro_free);
                    if (it > 0 && !(waited & 8u)) mbar_wait(&bars->panel_ready[3], (it - 1) & 1);
                }
            }
        }
    } else if (warp < 2 + kBwdEpiWarps) {
        // ======================= epilogue: ReLU mask, bf16, next A operand =======================
        // warp (q, hf): rows 32q..32q+31, columns [32 hf, 32 hf + 32) of every panel
        const int q = warp & 3, hf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = blockIdx.x + ti * gridDim.x;
            const uint8_t* act_tile = p.act + (size_t)tile * p.tile_stash_bytes;
            for (int s = 0; s < p.n_steps; ++s, ++it) {
                const TcStep& st = p.steps[s];
                const bool masked = st.kind == BWD_MASK && tile < p.n_tiles;   // dummy tiles carry zero gradients
                const uint8_t* mrow = act_tile + (size_t)st.mask_slot * 65536;
                bool acc_ready = false;
                for (int j = 0; j < 4; ++j) {
                    uint4 mk[4];
                    if (masked) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            mk[c] = __ldg(reinterpret_cast<const uint4*>(mrow + j * kPanelBytes + swz_offset(row, hf * 4 + c)));
                    }
                    if (!acc_ready) {
                        mbar_wait(&bars->acc_full[it & 1], (it >> 1) & 1);
                        tc_fence_after();
                        acc_ready = true;
                    }
                    float v[32];
                    tmem_ld32(lane_addr + (it & 1) * 256 + j * 64 + hf * 32, v);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
                    if (masked) {
                        const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t w[4] = {mk[c].x, mk[c].y, mk[c].z, mk[c].w};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                // dY * [h > 0] on a bf16 pair: __hgt2 yields 1.0 / 0.0 per half
                                const __nv_bfloat162 ind = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&w[h]), zero);
                                const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&pk[4 * c + h]), ind);
                                pk[4 * c + h] = *reinterpret_cast<const uint32_t*>(&r);
                            }
                        }
                    }
                    if (it > 0) mbar_wait(&bars->panel_stored[j], (it - 1) & 1);
                    uint8_t* dst = smem + kBOffH + j * kPanelBytes;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + swz_offset(row, hf * 4 + c)) =
                            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    fence_async_smem();
                    tc_fence_before();
                    mbar_arrive(&bars->panel_ready[j]);
                }
            }
        }
    } else if (warp < 2 + kBwdEpiWarps + 4) {
        // ======================= prologue: head gradients of the next tile =======================
        const int row = (warp - 2 - kBwdEpiWarps) * 32 + lane;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = blockIdx.x + ti * gridDim.x;
            const long long pt = (long long)tile * kTileRows + row;
            float ds = 0.f, g[3] = {0.f, 0.f, 0.f};
            if (pt < p.n_points) {
                ds = p.sigma[pt] > 0.f ? p.d_sigma[pt] : 0.f;                          // relu'(sigma_pre + noise)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float r = p.rgb[pt * 3 + c];
                    g[c] = p.d_rgb[pt * 3 + c] * r * (1.f - r);                        // sigmoid'
                }
            }
            uint4 mk[16];
            if (p.has_view) {
                const uint8_t* hv = p.act + (size_t)tile * p.tile_stash_bytes + (size_t)9 * 65536;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    mk[i] = tile < p.n_tiles ? __ldg(reinterpret_cast<const uint4*>(hv + (i >> 3) * kPanelBytes + swz_offset(row, i & 7)))
                                             : make_uint4(0u, 0u, 0u, 0u);
            }
            if (ti > 0) {
                mbar_wait(&bars->pro_free, (ti - 1) & 1);
                mbar_wait(&bars->pro_stored, (ti - 1) & 1);
            }
            uint8_t* pb = smem + kBOffP;
            if (p.has_view) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {   // i = panel * 8 + chunk; columns 8i .. 8i+7 of hv
                    const uint32_t w[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
                    float dv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = 8 * i + e;
                        const float d = g[0] * s_wrgb[col] + g[1] * s_wrgb[128 + col] + g[2] * s_wrgb[256 + col];
                        const uint32_t bits = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFFu);
                        dv[e] = bits != 0u ? d : 0.f;
                    }
                    const uint4 u = make_uint4(pack_bf16(dv[0], dv[1]), pack_bf16(dv[2], dv[3]), pack_bf16(dv[4], dv[5]),
                                               pack_bf16(dv[6], dv[7]));
                    *reinterpret_cast<uint4*>(pb + (i >> 3) * kPanelBytes + swz_offset(row, i & 7)) = u;
                }
            }
            // head-pre panel: column 0 = d sigma_pre, columns 1..3 = d rgb_pre when the rgb comes from the same head
            const uint4 h0 = p.has_view ? make_uint4(pack_bf16(ds, 0.f), 0u, 0u, 0u)
                                        : make_uint4(pack_bf16(ds, g[0]), pack_bf16(g[1], g[2]), 0u, 0u);
            *reinterpret_cast<uint4*>(pb + 2 * kPanelBytes + swz_offset(row, 0)) = h0;
            *reinterpret_cast<uint4*>(pb + 2 * kPanelBytes + swz_offset(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
            fence_async_smem();
            mbar_arrive(&bars->pro_ready);
        }
    } else {
        // ======================= stash writer =======================
        if (lane == 0) {
            uint32_t it = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int tile = blockIdx.x + ti * gridDim.x;
                uint8_t* base = p.dy + (size_t)tile * p.tile_stash_bytes;
                mbar_wait(&bars->pro_ready, ti & 1);
                if (p.has_view && tile < p.n_tiles) {
                    bulk_s2g(base + (size_t)9 * 65536, smem + kBOffP, 2 * kPanelBytes);
                    bulk_commit();
                    bulk_wait_read<0>();
                }
                mbar_arrive(&bars->pro_stored);
                for (int s = 0; s < p.n_steps; ++s, ++it) {
                    const TcStep& st = p.steps[s];
                    for (int j = 0; j < 4; ++j) {
                        mbar_wait(&bars->panel_ready[j], it & 1);
                        if (tile < p.n_tiles) {
                            bulk_s2g(base + (size_t)st.slot * 65536 + j * kPanelBytes, smem + kBOffH + j * kPanelBytes, kPanelBytes);
                            bulk_commit();
                        }
                    }
                    bulk_wait_read<0>();
                    for (int j = 0; j < 4; ++j) mbar_arrive(&bars->panel_stored[j]);
                }
            }
            bulk_wait_all<0>();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// =================================================================================================
// wgrad kernel
// =================================================================================================
constexpr int kWgThreads = 512;
constexpr int kWgStages = 3;
constexpr uint32_t kWgStageBytes = 65536;
constexpr uint32_t kHalfPanel = 8192;                               // 64 points x 128 bytes
constexpr uint32_t kWgOffDh = kWgStages * kWgStageBytes;             // [64][4] fp32 head-pre gradients of the current stage
constexpr uint32_t kWgOffBars = kWgOffDh + 1024;
constexpr uint32_t kWgSmem = kWgOffBars + 512 + 1024;
constexpr int kMaxJobs = 16;

struct WgSeg { uint16_t d_col, count, k_lo, k_hi, dst_col; };

struct WgJob {
    uint8_t kind;                 // 0 = MMA job, 1 = head matrix (streams h8), 2 = rgb matrix (streams hv)
    uint8_t a_slot, a_panels;     // dY stash slot, 4 (M=256) or 2 (M=128)
    uint8_t b_slot, b_panels;     // activation stash slot, panels streamed
    uint8_t b_enc, b_venc;        // computed operand panels appended after the streamed ones
    uint8_t n_segs;
    uint16_t aux_off;             // byte offset / 1024 of the computed panels inside a stage
    uint16_t pad;
    float* dw;
    float* db;
    int32_t ld;
    WgSeg seg[3];
};

struct WgParams {
    const uint8_t *act, *dy;
    const float *rays_o, *rays_d, *view_dirs, *z, *sigma, *rgb, *d_sigma, *d_rgb;
    long long n_points;
    int n_samples, n_tiles, n_jobs, pts_degree, view_degree, head_out;
    uint32_t tile_stash_bytes;
    WgJob jobs[kMaxJobs];
};

struct WgBars {
    uint64_t full[kWgStages], empty[kWgStages], aux_ready[kWgStages], acc_done, acc_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ float bf16_at(const uint8_t* panel_base, int r, int col) {
    const uint16_t bits = *reinterpret_cast<const uint16_t*>(panel_base + (col >> 6) * kHalfPanel + r * kRowBytes +
                                                             ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
    return __uint_as_float((uint32_t)bits << 16);
}

__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    WgBars* bars = (WgBars*)(smem + kWgOffBars);
    float* s_dh = (float*)(smem + kWgOffDh);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], 9);        // MMA commit + 8 reducer warps
            mbar_init(&bars->aux_ready[i], 128);
        }
        mbar_init(&bars->acc_done, 1);
        mbar_init(&bars->acc_free, 256);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_stages_per_job = my_tiles * 2;

    if (warp == 0) {
        // ======================= loader: stashed panels -> smem stages =======================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int jb = 0; jb < p.n_jobs; ++jb) {
                const WgJob& job = p.jobs[jb];
                const uint32_t bytes = (uint32_t)(job.a_panels + job.b_panels) * kHalfPanel;
                for (int sg = 0; sg < n_stages_per_job; ++sg, ++cnt) {
                    const int tile = blockIdx.x + (sg >> 1) * gridDim.x, half = sg & 1;
                    const uint32_t stage = cnt % kWgStages, round = cnt / kWgStages;
                    if (round > 0) mbar_wait(&bars->empty[stage], (round - 1) & 1);
                    uint8_t* dst = smem + stage * kWgStageBytes;
                    mbar_arrive_expect_tx(&bars->full[stage], bytes);
                    const size_t toff = (size_t)tile * p.tile_stash_bytes + (size_t)half * kHalfPanel;
                    for (int j = 0; j < job.a_panels; ++j)
                        bulk_g2s(dst + j * kHalfPanel, p.dy + toff + (size_t)job.a_slot * 65536 + j * kPanelBytes, kHalfPanel,
                                 &bars->full[stage]);
                    for (int j = 0; j < job.b_panels; ++j)
                        bulk_g2s(dst + 32768 + j * kHalfPanel, p.act + toff + (size_t)job.b_slot * 65536 + j * kPanelBytes,
                                 kHalfPanel, &bars->full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            uint32_t cnt = 0, flushes = 0;
            for (int jb = 0; jb < p.n_jobs; ++jb) {
                const WgJob& job = p.jobs[jb];
                const bool aux = job.b_enc || job.b_venc;
                const int n1 = job.b_panels * 64;
                const int n2 = ((int)job.b_enc + (int)job.b_venc) * 64;
                const int n_mb = job.a_panels / 2;
                if (job.kind == 0 && flushes > 0) { mbar_wait(&bars->acc_free, (flushes - 1) & 1); tc_fence_after(); }
                for (int sg = 0; sg < n_stages_per_job; ++sg, ++cnt) {
                    const uint32_t stage = cnt % kWgStages;
                    mbar_wait(&bars->full[stage], (cnt / kWgStages) & 1);
                    if (job.kind != 0) { mbar_arrive(&bars->empty[stage]); continue; }
                    if (aux) mbar_wait(&bars->aux_ready[stage], (cnt / kWgStages) & 1);
                    tc_fence_after();
                    const uint32_t sbase = smem_u32(smem + stage * kWgStageBytes);
                    for (int k = 0; k < 4; ++k) {
                        const bool accumulate = (sg | k) != 0;
                        for (int mb = 0; mb < n_mb; ++mb) {
                            const uint64_t ad = umma_desc_mnmajor(sbase + mb * 2 * kHalfPanel, k, kHalfPanel);
                            if (n1 > 0)
                                umma(tmem + mb * 256, ad, umma_desc_mnmajor(sbase + 32768, k, kHalfPanel),
                                     umma_idesc(128, n1, true, true), accumulate);
                            if (n2 > 0)
                                umma(tmem + mb * 256 + n1, ad, umma_desc_mnmajor(sbase + (uint32_t)job.aux_off * 1024u, k, kHalfPanel),
                                     umma_idesc(128, n2, true, true), accumulate);
                        }
                    }
                    umma_commit(&bars->empty[stage]);
                }
                if (job.kind == 0) { umma_commit(&bars->acc_done); ++flushes; }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ======================= reducers: bias sums, head matrices, TMEM flush =======================
        const int t = (warp - 4) * 32 + lane;            // column owned by this thread
        const int q = warp & 3, colhalf = (warp - 4) >> 2;
        uint32_t cnt = 0, flushes = 0;
        for (int jb = 0; jb < p.n_jobs; ++jb) {
            const WgJob& job = p.jobs[jb];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            float acc_b = 0.f;
            const int nh = job.kind == 1 ? p.head_out : 3;
            for (int sg = 0; sg < n_stages_per_job; ++sg, ++cnt) {
                const int tile = blockIdx.x + (sg >> 1) * gridDim.x, half = sg & 1;
                const uint32_t stage = cnt % kWgStages;
                const uint8_t* sb = smem + stage * kWgStageBytes;
                if (job.kind != 0) {
                    // head-pre gradients of the 64 points of this stage
                    if (t < 64) {
                        const long long pt = (long long)tile * kTileRows + half * 64 + t;
                        float d[4] = {0.f, 0.f, 0.f, 0.f};
                        if (pt < p.n_points) {
                            const float ds = p.sigma[pt] > 0.f ? p.d_sigma[pt] : 0.f;
                            float g[3];
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float r = p.rgb[pt * 3 + c];
                                g[c] = p.d_rgb[pt * 3 + c] * r * (1.f - r);
                            }
                            if (job.kind == 1) { d[0] = ds; if (p.head_out == 4) { d[1] = g[0]; d[2] = g[1]; d[3] = g[2]; } }
                            else { d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) s_dh[t * 4 + c] = d[c];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                mbar_wait(&bars->full[stage], (cnt / kWgStages) & 1);
                if (job.kind == 0) {
                    if (job.db != nullptr && t < job.a_panels * 64) {
                        for (int r = 0; r < 64; ++r) acc_b += bf16_at(sb, r, t);
                    }
                } else {
                    const int ncol = job.b_panels * 64;
                    if (t < ncol) {
                        for (int r = 0; r < 64; ++r) {
                            const float x = bf16_at(sb + 32768, r, t);
                            const float4 d = *reinterpret_cast<const float4*>(s_dh + r * 4);
                            acc[0] = fmaf(d.x, x, acc[0]); acc[1] = fmaf(d.y, x, acc[1]);
                            acc[2] = fmaf(d.z, x, acc[2]); acc[3] = fmaf(d.w, x, acc[3]);
                        }
                    }
                    if (t < nh) {
                        for (int r = 0; r < 64; ++r) acc_b += s_dh[r * 4 + t];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->empty[stage]);
            }
            // ---- job results ----
            if (job.kind == 0) {
                if (job.db != nullptr && t < job.a_panels * 64) atomicAdd(job.db + t, acc_b);
                mbar_wait(&bars->acc_done, flushes & 1);
                tc_fence_after();
                const int n_mb = job.a_panels / 2;
                for (int mb = 0; mb < n_mb; ++mb) {
                    const int out_row = mb * 128 + q * 32 + lane;
                    for (int sgi = 0; sgi < job.n_segs; ++sgi) {
                        const WgSeg sg = job.seg[sgi];
                        for (int c0 = colhalf * 32; c0 < sg.count; c0 += 64) {
                            float v[32];
                            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mb * 256 + sg.d_col + c0, v);
                            float* dst = job.dw + (size_t)out_row * job.ld + sg.dst_col;
                            const bool vec = ((job.ld | sg.dst_col | sg.k_lo) & 3) == 0 && c0 >= sg.k_lo && c0 + 32 <= sg.k_hi &&
                                             (reinterpret_cast<uintptr_t>(job.dw) & 15) == 0;
                            if (vec) {   // whole 32-column chunk valid and 16-byte aligned: 8 vector reductions
#pragma unroll
                                for (int i = 0; i < 32; i += 4) red_add_v4(dst + (c0 - sg.k_lo) + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const int n = c0 + i;
                                    if (n >= sg.k_lo && n < sg.k_hi) atomicAdd(dst + (n - sg.k_lo), v[i]);
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars->acc_free);
                ++flushes;
            } else {
                const int ncol = job.b_panels * 64;
                if (t < ncol)
                    for (int h = 0; h < nh; ++h) atomicAdd(job.dw + (size_t)h * job.ld + t, acc[h]);
                if (t < nh && job.db != nullptr) atomicAdd(job.db + t, acc_b);
            }
        }
    } else if (warp >= 12) {
        // ======================= encoders: recomputed encoding operands =======================
        const int e = (warp - 12) * 32 + lane;
        const int r = e >> 1, hf = e & 1;
        uint32_t cnt = 0;
        for (int jb = 0; jb < p.n_jobs; ++jb) {
            const WgJob& job = p.jobs[jb];
            const bool aux = job.kind == 0 && (job.b_enc || job.b_venc);
            for (int sg = 0; sg < n_stages_per_job; ++sg, ++cnt) {
                const int tile = blockIdx.x + (sg >> 1) * gridDim.x, half = sg & 1;
                const uint32_t stage = cnt % kWgStages, round = cnt / kWgStages;
                if (!aux) {   // keep one aux_ready phase per stage use so that parities stay in step with `full`
                    if (round > 0) mbar_wait(&bars->empty[stage], (round - 1) & 1);
                    mbar_arrive(&bars->aux_ready[stage]);
                    continue;
                }
                const long long pt = (long long)tile * kTileRows + half * 64 + r;
                const bool valid = pt < p.n_points;
                const int ray = valid ? (int)(pt / p.n_samples) : 0;
                float enc[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) enc[i] = 0.f;
                if (job.b_enc) {
                    float x[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        const float zz = p.z[pt];
#pragma unroll
                        for (int c = 0; c < 3; ++c) x[c] = fmaf(p.rays_d[ray * 3 + c], zz, p.rays_o[ray * 3 + c]);
                    }
                    encode_point(x, p.pts_degree, enc);
                }
                if (round > 0) mbar_wait(&bars->empty[stage], (round - 1) & 1);
                uint8_t* dst = smem + stage * kWgStageBytes + (uint32_t)job.aux_off * 1024u;
                if (job.b_enc) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ch = hf * 4 + c;
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = hf ? enc[32 + 8 * c + i] : enc[8 * c + i];
                        *reinterpret_cast<uint4*>(dst + swz_offset(r, ch)) =
                            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                    dst += kHalfPanel;
                }
                if (job.b_venc) {
                    float ve[64];
#pragma unroll
                    for (int i = 0; i < 64; ++i) ve[i] = 0.f;
                    if (hf == 0) {
                        float vd[3] = {p.view_dirs[ray * 3], p.view_dirs[ray * 3 + 1], p.view_dirs[ray * 3 + 2]};
                        encode_point(vd, p.view_degree, ve);
                        const int venc = 3 * (1 + 2 * p.view_degree);
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (i >= venc) ve[i] = 0.f;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ch = hf * 4 + c;
                        *reinterpret_cast<uint4*>(dst + swz_offset(r, ch)) =
                            make_uint4(pack_bf16(ve[8 * c], ve[8 * c + 1]), pack_bf16(ve[8 * c + 2], ve[8 * c + 3]),
                                       pack_bf16(ve[8 * c + 4], ve[8 * c + 5]), pack_bf16(ve[8 * c + 6], ve[8 * c + 7]));
                    }
                }
                fence_async_smem();
                mbar_arrive(&bars->aux_ready[stage]);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// =================================================================================================
// driver
// =================================================================================================
static WgJob mma_job(int a_slot, int a_panels, int b_slot, int b_panels, bool enc, bool venc, float* dw, int ld, float* db) {
    WgJob j{};
    j.kind = 0; j.a_slot = (uint8_t)a_slot; j.a_panels = (uint8_t)a_panels; j.b_slot = (uint8_t)b_slot;
    j.b_panels = (uint8_t)b_panels; j.b_enc = enc; j.b_venc = venc; j.dw = dw; j.ld = ld; j.db = db;
    j.aux_off = (uint16_t)((a_panels == 2 ? 16384 : 32768 + b_panels * 8192) / 1024);
    return j;
}

int tc_backward(const snerf_mlp_desc& d, const float* const* prm, const void* packed, const float* rays_o, const float* rays_d,
                const float* view_dirs, const float* z, const float* sigma, const float* rgb, const float* d_sigma,
                const float* d_rgb, float* const* grads, void* ws, size_t ws_bytes, int n_rays, int n_samples, uint32_t flags,
                cudaStream_t st) {
    const MlpDims m(d);
    const TcPlan pl = build_plan(d, prm);
    const TcWorkspace w = tc_ws_layout(m, pl, n_rays, n_samples, flags);
    SNERF_REQUIRE(ws_bytes >= w.total, "mlp_backward: workspace too small (%zu < %zu)", ws_bytes, w.total);
    uint8_t* wsb = (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    const long long P = (long long)n_rays * n_samples;
    const int grid = w.n_tiles < num_sms() ? w.n_tiles : num_sms();

    // ---- (1) dgrad chain ----
    BwdParams bp{};
    bp.packed = (const uint8_t*)packed; bp.act = wsb + w.act; bp.dy = wsb + w.dy;
    bp.sigma = sigma; bp.rgb = rgb; bp.d_sigma = d_sigma; bp.d_rgb = d_rgb;
    bp.w_rgb = m.has_view ? prm[SNERF_P_RGB_W] : nullptr;
    bp.n_points = P; bp.n_tiles = w.n_tiles; bp.n_steps = pl.n_bwd; bp.has_view = m.has_view ? 1 : 0;
    bp.tile_stash_bytes = pl.tile_stash_bytes;
    for (int s = 0; s < pl.n_bwd; ++s) bp.steps[s] = pl.bwd[s];
    static bool attr = false;
    if (!attr) {
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
        attr = true;
    }
    SNERF_CUDA_OK(launch_clustered(tc_dgrad_kernel, chain_grid(w.n_tiles), kBwdThreads, kBwdSmem, st, bp));

    // ---- (2) wgrad ----
    WgParams wp{};
    wp.act = wsb + w.act; wp.dy = wsb + w.dy;
    wp.rays_o = rays_o; wp.rays_d = rays_d; wp.view_dirs = view_dirs; wp.z = z;
    wp.sigma = sigma; wp.rgb = rgb; wp.d_sigma = d_sigma; wp.d_rgb = d_rgb;
    wp.n_points = P; wp.n_samples = n_samples; wp.n_tiles = w.n_tiles; wp.pts_degree = d.pts_degree;
    wp.view_degree = d.view_degree; wp.head_out = m.head_out; wp.tile_stash_bytes = pl.tile_stash_bytes;
    int nj = 0;
    auto seg = [](int d_col, int count, int k_lo, int k_hi, int dst_col) {
        WgSeg s; s.d_col = (uint16_t)d_col; s.count = (uint16_t)count; s.k_lo = (uint16_t)k_lo; s.k_hi = (uint16_t)k_hi;
        s.dst_col = (uint16_t)dst_col; return s;
    };
    for (int l = 0; l < m.depth; ++l) {
        const int fan_in = m.trunk_fan_in(l);
        if (l == 0) {                                   // dW0 = dY0^T E
            WgJob j = mma_job(0, 4, 0, 0, true, false, grads[0], fan_in, grads[1]);
            j.n_segs = 1; j.seg[0] = seg(0, 64, 0, m.trunk_in, 0);
            wp.jobs[nj++] = j;
            continue;
        }
        const int hcol0 = (l - 1 == m.skip_layer) ? m.trunk_in : 0;
        WgJob j = mma_job(l, 4, l - 1, 4, false, false, grads[2 * l], fan_in, grads[2 * l + 1]);   // dW_l[:, hidden] = dY_l^T h_l
        j.n_segs = 1; j.seg[0] = seg(0, 256, 0, 256, hcol0);
        wp.jobs[nj++] = j;
        if (hcol0 > 0) {                                // skip layer: dW_l[:, :trunk_in] = dY_l^T E
            WgJob e = mma_job(l, 4, 0, 0, true, false, grads[2 * l], fan_in, nullptr);
            e.n_segs = 1; e.seg[0] = seg(0, 64, 0, m.trunk_in, 0);
            wp.jobs[nj++] = e;
        }
    }
    if (m.has_view) {
        WgJob f = mma_job(8, 4, 7, 4, false, false, grads[SNERF_P_FEAT_W], m.width, grads[SNERF_P_FEAT_B]);   // dW_feat = dY_f^T h8
        f.n_segs = 1; f.seg[0] = seg(0, 256, 0, 256, 0);
        wp.jobs[nj++] = f;
        // dW_view = dY_v^T [feature | E(bands >= trunk_degree) | PE(view dir)]
        const bool hi = m.enc_hi > 0;
        WgJob v = mma_job(9, 2, 8, 4, hi, true, grads[SNERF_P_VIEW_W], m.view_in, grads[SNERF_P_VIEW_B]);
        v.seg[0] = seg(0, 256, 0, 256, 0);
        if (hi) {
            v.n_segs = 3;
            v.seg[1] = seg(256, 64, m.trunk_in, m.enc, m.width);
            v.seg[2] = seg(320, 64, 0, m.venc, m.width + m.enc_hi);
        } else {
            v.n_segs = 2;
            v.seg[1] = seg(256, 64, 0, m.venc, m.width);
        }
        wp.jobs[nj++] = v;
        WgJob r{};                                      // dW_rgb = d rgb_pre^T hv
        r.kind = 2; r.b_slot = 9; r.b_panels = 2; r.dw = grads[SNERF_P_RGB_W]; r.ld = m.view_width; r.db = grads[SNERF_P_RGB_B];
        wp.jobs[nj++] = r;
    }
    WgJob h{};                                          // dW_head = d head_pre^T h8
    h.kind = 1; h.b_slot = 7; h.b_panels = 4; h.dw = grads[SNERF_P_HEAD_W]; h.ld = m.width; h.db = grads[SNERF_P_HEAD_B];
    wp.jobs[nj++] = h;
    wp.n_jobs = nj;
    tc_wgrad_kernel<<<grid, kWgThreads, kWgSmem, st>>>(wp);
    SNERF_LAUNCH_OK("tc_wgrad_kernel");
    return SNERF_OK;
}

}  // namespace snerf
