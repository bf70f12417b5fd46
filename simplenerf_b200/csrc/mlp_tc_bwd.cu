// bf16 tcgen05 MLP backward: (1) a fused dgrad chain kernel, (2) a persistent wgrad kernel.
//
// The forward kernel (mlp_tc.cu) left the bf16 activation panels of every layer in the HBM stash as raw
// swizzled smem images.  Backward of the reference MLP (src/models/SimpleNeRF01.py :626-715 under autograd):
//
//  dgrad kernel  -- the CTA-pair machine of the forward kernel run backwards (two 256-point super tiles in flight):
//      prologue warps : d rgb_pre = d rgb * rgb (1-rgb), d sigma_pre = d sigma * [sigma>0],
//                       dY_v = (d rgb_pre W_rgb) * [hv>0]            (CUDA cores, fp32)
//      MMA chain      : d h8 = dY_v W_vf + d head_pre W_head   (W_vf = W_view[:, :256] W_feat, the merged view branch of
//                       tc_plan.cuh);  d h_l = dY_l W_l[:, hidden]  for l = 7..1    (tcgen05 cta_group::2, B = packed W^T chunks)
//      epilogue warps : ReLU mask from the sign bits the forward epilogue filed, bf16, back to smem as the next A operand
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
// dgrad chain kernel (CTA pairs, two tiles in flight -- the machine of the forward kernel run backwards)
// =================================================================================================
// roles by warp id: 0-3 prologue | 4 stash writer | 5 weight loader | 6-21 epilogue | 22 MMA issuer (leader) / relay (peer)
// Per pair and group: two 256-point super tiles occupy slots 0/1; the issuer alternates between them step by step and
// both tiles share the weight chunks of a step (mlp_tc.cu describes the protocol).  The prologue writes dY_v straight
// into panels 0-1 of the slot (they are free until the first epilogue of the tile) and the 16-column head-pre panel
// into one small buffer that is rewritten before each use.
constexpr int kBwdThreads = 736;          // 23 warps
constexpr int kBwdEpiWarps = 16;          // 4 per SM sub-partition, one 64-column panel each
constexpr int kBStages = 4;
constexpr uint32_t kBStageBytes = 16384;
constexpr uint32_t kBOffH = 0;                                          // [2 slots][4 panels]
constexpr uint32_t kBOffHP = 2 * 65536;                                 // head-pre panel
constexpr uint32_t kBOffRing = kBOffHP + kPanelBytes;                   // 147456
constexpr uint32_t kBOffConst = kBOffRing + kBStages * kBStageBytes;    // 212992: w_rgb [3][128] fp32
constexpr uint32_t kBOffBars = kBOffConst + 2048;
constexpr uint32_t kBwdSmem = kBOffBars + 512 + 1024;
static_assert(kBwdSmem <= 232448, "shared memory budget");

struct BwdParams {
    const uint8_t* packed;
    const uint8_t* act;
    const uint8_t* bits;          // ReLU sign bits written by the forward kernel (kBitsTileBytes per tile)
    RingCtl ring;                 // destination of the gradient panels
    const float *sigma, *rgb, *d_sigma, *d_rgb, *w_rgb;
    int stash_pieces;              // 4: the gradient panels leave one per bulk-copy request (paced); 1: one request per job
    const float* vis_extra;        // (SNERF_FLAG_VIS_GRAD) fp32 [n_points,128]: the visibility head's share of dY_v, masks applied (vis_tc.cu); else null
    long long n_points;
    int n_tiles, n_steps, has_view;
    int n_pairs;                  // CTA pairs running the chain
    uint32_t tile_stash_bytes;
    TcStep steps[kMaxSteps];
};

struct BwdBars {
    uint64_t w_full[kBStages];     // leader: own bytes + the peer's relay (2 arrivals); peer: own bytes (1)
    uint64_t w_empty[kBStages];    // MMA commit, multicast
    uint64_t acc_full[2];          // per slot: MMA commit, multicast
    uint64_t tile_ready[2];        // leader only: 2 x 8 epilogue warps
    uint64_t stash_ready[2];       // local: 8 epilogue warps (one phase per job)
    uint64_t stash_done[2];        // local: stash writer (one phase per stash event: dY_v of a tile, then every job)
    uint64_t pro_ready[2];         // leader only: 2 x 4 prologue warps, dY_v of the slot's tile is in panels 0-1
    uint64_t pro_local[2];         // local: 4 prologue warps -> stash writer
    uint64_t slot_free[2];         // local: stash writer -> prologue, one phase per tile: the tile's last gradient panel has been copied out
    uint64_t hp_ready;             // leader only: 2 x 4 prologue warps, one phase per use of the head-pre panel
    uint64_t hp_free;              // MMA commit, multicast
    uint32_t tmem_base;
};

// kVis: the prologue adds the visibility head's share of dY_v (a template parameter: the plain kernel keeps its registers)
template <bool kVis>
__device__ __forceinline__ void dgrad_role(const BwdParams& p, uint8_t* smem, const int pair_index) {
    BwdBars* bars = (BwdBars*)(smem + kBOffBars);
    float* s_wrgb = (float*)(smem + kBOffConst);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_rank();
    constexpr int kWarpStash = 4, kWarpLoader = 5, kWarpEpi0 = 6, kWarpMma = 22;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kBStages; ++i) { mbar_init(&bars->w_full[i], rank == 0 ? 2 : 1); mbar_init(&bars->w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->acc_full[i], 1);
            mbar_init(&bars->tile_ready[i], 2 * kBwdEpiWarps);
            mbar_init(&bars->stash_ready[i], kBwdEpiWarps);
            mbar_init(&bars->stash_done[i], 1);
            mbar_init(&bars->pro_ready[i], 2 * 4);
            mbar_init(&bars->pro_local[i], 4);
            mbar_init(&bars->slot_free[i], 1);
        }
        mbar_init(&bars->hp_ready, 2 * 4);
        mbar_init(&bars->hp_free, 1);
        mbar_fence_init();
    }
    if (warp == kWarpMma) tmem_alloc2<512>(&bars->tmem_base);
    if (p.w_rgb) for (int i = threadIdx.x; i < 3 * 128; i += kBwdThreads) s_wrgb[i] = p.w_rgb[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    // Group g of pair pi = super tiles 2 (g n_pairs + pi) and + 1 in slots 0 / 1: the four tiles a pair has in flight are
    // consecutive, and tiles are produced in (roughly) increasing order over the whole launch -- the order in which the
    // wgrad CTAs consume them, which is what lets a small ring make progress.  Every pair runs the
    // same number of whole groups; tiles >= n_tiles are dummies.
    const int n_pairs = p.n_pairs, pi = pair_index;
    const int n_super = (p.n_tiles + 1) / 2;
    const int my_super = 2 * ((n_super + 2 * n_pairs - 1) / (2 * n_pairs));
    auto tile_of = [&](int i) { return 2 * (2 * ((i >> 1) * n_pairs + pi) + (i & 1)) + (int)rank; };
    const int pv = p.has_view ? 1 : 0;
    const int n_events = p.n_steps + pv;                        // stash events per tile and slot

    if (warp == kWarpLoader) {
        // ======================= weight loader: this CTA's N/2 rows of every chunk, once per step when both tiles can share =======================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s) {
                    const TcStep& st = p.steps[s];
                    const uint32_t bytes = (uint32_t)st.n_rows * kRowBytes, half = bytes / 2;
                    const int reps = (two && st.n_chunks > kBStages) ? 2 : 1;
                    for (int r = 0; r < reps; ++r)
                        for (int c = 0; c < st.n_chunks; ++c, ++cnt) {
                            const uint32_t stage = cnt % kBStages, round = cnt / kBStages;
                            if (round > 0) mbar_wait(&bars->w_empty[stage], (round - 1) & 1);
                            mbar_arrive_expect_tx(&bars->w_full[stage], half);
                            bulk_g2s(smem + kBOffRing + stage * kBStageBytes, p.packed + st.w_off + (uint32_t)c * bytes + rank * half, half,
                                     &bars->w_full[stage]);
                        }
                }
            }
        }
    } else if (warp == kWarpMma && rank != 0) {
        // ======================= peer: relay "my half of the chunk has landed" to the leader =======================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int g = 0; 2 * g < my_super; ++g) {
                const bool two = 2 * g + 1 < my_super;
                for (int s = 0; s < p.n_steps; ++s) {
                    const TcStep& st = p.steps[s];
                    const int n = ((two && st.n_chunks > kBStages) ? 2 : 1) * st.n_chunks;
                    for (int c = 0; c < n; ++c, ++cnt) {
                        const uint32_t stage = cnt % kBStages;
                        mbar_wait_spin(&bars->w_full[stage], (cnt / kBStages) & 1);
                        mbar_arrive_cluster(cluster_addr(&bars->w_full[stage], 0));
                    }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ======================= MMA issuer (leader CTA; converged warp, one elected lane) =======================
        uint32_t stage = 0, wpar = 0, hu = 0;
        const uint32_t ring_lo = desc_lo_kmajor(smem_u32(smem + kBOffRing));
        const uint32_t h_lo = desc_lo_kmajor(smem_u32(smem + kBOffH)), hp_lo = desc_lo_kmajor(smem_u32(smem + kBOffHP));
        const uint32_t idesc = umma_idesc(256, 256, false, false);
        bool w_ok = false, t_ok = false;
        auto issue_job = [&](const int x, const int g, const int s, const bool release, const bool landed) {
            const TcStep& st = p.steps[s];
            const uint32_t jx = (uint32_t)(g * p.n_steps + s);
            const uint32_t d_tmem = tmem + x * 256;
            if (jx > 0 && !t_ok) mbar_wait(&bars->tile_ready[x], (jx - 1) & 1);
            if (s == 0 && pv) mbar_wait(&bars->pro_ready[x], g & 1);
            const int nc = st.n_chunks;
            for (int c = 0; c < nc; ++c) {
                const int pn = st.panel[c], nk = st.ksteps[c];
                const bool is_hp = pn == kPanelP + 2;
                if (!landed && !w_ok) mbar_wait(&bars->w_full[stage], wpar);
                if (is_hp) mbar_wait(&bars->hp_ready, hu & 1);
                tc_fence_after();
                const uint32_t a_lo = is_hp ? hp_lo : h_lo + x * (65536 >> 4) + (pn >= kPanelP ? pn - kPanelP : pn) * (kPanelBytes >> 4);
                const uint32_t b_lo = ring_lo + stage * (kBStageBytes >> 4);
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
                    if (release) umma_commit2(done, 3);
                    if (is_hp) umma_commit2(&bars->hp_free, 3);
                }
                __syncwarp();
                if (is_hp) ++hu;
                if (++stage == kBStages) { stage = 0; wpar ^= 1; }
                if (!landed) w_ok = mbar_test_wait(&bars->w_full[stage], wpar);
            }
            if (elect_one()) umma_commit2(&bars->acc_full[x], 3);
            __syncwarp();
        };
        for (int g = 0; 2 * g < my_super; ++g) {
            const bool two = 2 * g + 1 < my_super;
            for (int s = 0; s < p.n_steps; ++s) {
                const uint32_t jx = (uint32_t)(g * p.n_steps + s);
                const bool shared = two && p.steps[s].n_chunks <= kBStages;
                const uint32_t stage0 = stage, wpar0 = wpar;
                issue_job(0, g, s, !shared, false);
                if (two) {
                    t_ok = jx > 0 && mbar_test_wait(&bars->tile_ready[1], (jx - 1) & 1);
                    if (shared) { stage = stage0; wpar = wpar0; }
                    issue_job(1, g, s, true, shared);
                    if (shared) w_ok = mbar_test_wait(&bars->w_full[stage], wpar);
                }
                t_ok = mbar_test_wait(&bars->tile_ready[0], jx & 1);
            }
        }
    } else if (warp >= kWarpEpi0 && warp < kWarpEpi0 + kBwdEpiWarps) {
        // ======================= epilogue: ReLU mask, bf16, next A operand =======================
        // warp (q, j): rows 32q..32q+31, the 64-column panel j; 16-column TMEM units, the next one in flight.  The
        // ReLU mask of the whole panel row is two words of sign bits written by the forward epilogue (tc_plan.cuh).
        const int q = warp & 3, j = (warp - kWarpEpi0) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + j * 64;
        const uint32_t ready0 = cluster_addr(&bars->tile_ready[0], 0), ready1 = cluster_addr(&bars->tile_ready[1], 0);
        // chunk c of this thread's panel row lives at (panel + row_base) ^ (c << 4)  (128-byte swizzle)
        const uint32_t row_base = smem_u32(smem + kBOffH) + (uint32_t)row * kRowBytes + (((uint32_t)row & 7u) << 4);
        auto job = [&](const int x, const int g, const int s) {
            const TcStep& st = p.steps[s];
            const uint32_t jx = (uint32_t)(g * p.n_steps + s);
            const uint32_t ev = (uint32_t)(g * n_events + pv + s);          // this job's stash event within the slot
            constexpr bool mask = true;      // every step of the chain ends in a ReLU mask (the linear feature step is merged away)
            uint2 mw = make_uint2(0u, 0u);                                  // dummy tiles carry zero gradients
            if (mask) {
                const int tile = tile_of(2 * g + x);
                if (tile < p.n_tiles)
                    mw = __ldg(reinterpret_cast<const uint2*>(p.bits + (size_t)tile * kBitsTileBytes + (size_t)st.mask_slot * kBitsSlotBytes + row * 32 + j * 8));
            }
            mbar_wait(&bars->acc_full[x], jx & 1);
            tc_fence_after();
            const uint32_t acc_addr = lane_addr + x * 256;
            uint32_t dst_row = row_base + x * 65536 + j * kPanelBytes;
            asm volatile("" : "+r"(dst_row));       // keep the address in its register: ptxas otherwise re-derives it from the row per store
            uint32_t rr[2][16];
            tmem_ld16_issue(acc_addr, rr[0]);
            if (ev > 0) mbar_wait(&bars->stash_done[x], (ev - 1) & 1);     // the slot's panels have been copied out
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                tmem_ld_wait16(rr[u & 1]);
                if (u + 1 < 4) tmem_ld16_issue(acc_addr + (u + 1) * 16, rr[(u + 1) & 1]);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(rr[u & 1][2 * i]), __uint_as_float(rr[u & 1][2 * i + 1]));
                if (mask) relu_mask_unit((u >> 1) ? mw.y : mw.x, u, pk);
                sts128(dst_row ^ ((2 * u) << 4), pk[0], pk[1], pk[2], pk[3]);
                sts128(dst_row ^ ((2 * u + 1) << 4), pk[4], pk[5], pk[6], pk[7]);
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bars->stash_ready[x]);
                mbar_arrive_cluster(x ? ready1 : ready0);
            }
        };
        for (int g = 0; 2 * g < my_super; ++g) {
            const bool two = 2 * g + 1 < my_super;
            for (int s = 0; s < p.n_steps; ++s) {
                job(0, g, s);
                if (two) job(1, g, s);
            }
        }
    } else if (warp < 4) {
        // ======================= prologue: head gradients =======================
        // d rgb_pre = d rgb * rgb (1-rgb), d sigma_pre = d sigma * [sigma>0], dY_v = (d rgb_pre W_rgb) * [hv>0]  (CUDA cores, fp32)
        const int row = warp * 32 + lane;
        const uint32_t pro0 = cluster_addr(&bars->pro_ready[0], 0), pro1 = cluster_addr(&bars->pro_ready[1], 0);
        const uint32_t hp_addr = cluster_addr(&bars->hp_ready, 0);
        uint32_t hu = 0;
        for (int g = 0; 2 * g < my_super; ++g) {
            const bool two = 2 * g + 1 < my_super;
            uint4 hp[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                hp[x] = make_uint4(0u, 0u, 0u, 0u);
                if (x == 1 && !two) continue;
                const int tile = tile_of(2 * g + x);
                const long long pt = (long long)tile * kTileRows + row;
                float ds = 0.f, gr[3] = {0.f, 0.f, 0.f};
                if (pt < p.n_points) {
                    ds = p.sigma[pt] > 0.f ? p.d_sigma[pt] : 0.f;                          // relu'(sigma_pre + noise)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float r = p.rgb[pt * 3 + c];
                        gr[c] = p.d_rgb[pt * 3 + c] * r * (1.f - r);                       // sigmoid'
                    }
                }
                // head-pre row: column 0 = d sigma_pre, columns 1..3 = d rgb_pre when the rgb comes from the same head
                hp[x] = p.has_view ? make_uint4(pack_bf16(ds, 0.f), 0u, 0u, 0u) : make_uint4(pack_bf16(ds, gr[0]), pack_bf16(gr[1], gr[2]), 0u, 0u);
                if (p.has_view) {
                    uint4 mk[16];
                    const uint8_t* hv = p.act + (size_t)tile * p.tile_stash_bytes + (size_t)kSlotHv * 65536;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        mk[i] = tile < p.n_tiles ? __ldg(reinterpret_cast<const uint4*>(hv + (i >> 3) * kPanelBytes + swz_offset(row, i & 7)))
                                                 : make_uint4(0u, 0u, 0u, 0u);
                    // panels 0-1 of the slot are free once the previous tile's last gradient panel has been copied out.  (A wait on
                    // stash_done would alias: this warp runs many phases ahead of that barrier; slot_free has one phase per tile.)
                    if (g > 0) mbar_wait(&bars->slot_free[x], (g - 1) & 1);
                    uint8_t* pb = smem + kBOffH + x * 65536;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {   // i = panel * 8 + chunk; columns 8i .. 8i+7 of hv
                        const uint32_t w[4] = {mk[i].x, mk[i].y, mk[i].z, mk[i].w};
                        float dv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int col = 8 * i + e;
                            const float d = gr[0] * s_wrgb[col] + gr[1] * s_wrgb[128 + col] + gr[2] * s_wrgb[256 + col];
                            const uint32_t bits = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFFu);
                            dv[e] = bits != 0u ? d : 0.f;
                        }
                        if (kVis && pt < p.n_points) {      // fourth row of the view head, own and other views (:646-649, :710-713)
                            const float4* ex = reinterpret_cast<const float4*>(p.vis_extra + (size_t)pt * 128 + 8 * i);
                            const float4 a = __ldg(ex), b = __ldg(ex + 1);
                            dv[0] += a.x; dv[1] += a.y; dv[2] += a.z; dv[3] += a.w; dv[4] += b.x; dv[5] += b.y; dv[6] += b.z; dv[7] += b.w;
                        }
                        *reinterpret_cast<uint4*>(pb + (i >> 3) * kPanelBytes + swz_offset(row, i & 7)) =
                            make_uint4(pack_bf16(dv[0], dv[1]), pack_bf16(dv[2], dv[3]), pack_bf16(dv[4], dv[5]), pack_bf16(dv[6], dv[7]));
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&bars->pro_local[x]);
                        mbar_arrive_cluster(x ? pro1 : pro0);
                    }
                }
            }
            // the head-pre panel, rewritten before each use in the issuer's order (tile of slot 0, then slot 1)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                if (x == 1 && !two) continue;
                if (hu > 0) mbar_wait(&bars->hp_free, (hu - 1) & 1);
                uint8_t* pb = smem + kBOffHP;
                *reinterpret_cast<uint4*>(pb + swz_offset(row, 0)) = hp[x];
                *reinterpret_cast<uint4*>(pb + swz_offset(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(hp_addr);
                ++hu;
            }
        }
    } else if (warp == kWarpStash) {
        // ======================= stash writer: every gradient panel -> the ring read by the wgrad CTAs =======================
        // Entry tile % cap of the slot's ring; before overwriting an entry its previous tile must have been copied out by
        // every consumer, and a tile is announced (ready flag) once its bulk store has COMPLETED -- checked one event late
        // (wait_group 1), so that the store's latency hides under the next job.
        if (lane == 0) {
            const RingCtl& ring = p.ring;
            uint32_t* pend_flag = nullptr;
            uint32_t pend_val = 0;
            auto flush = [&]() {
                if (pend_flag) { fence_proxy_async_all(); st_release_gpu(pend_flag, pend_val); pend_flag = nullptr; }
            };
            auto emit = [&](const int tile, const int slot, const uint8_t* src, const uint32_t bytes) {
                const uint32_t e = (uint32_t)tile % ring.cap;
                if (ring.use_flags && (uint32_t)tile >= ring.cap) {
                    const uint32_t want = (uint32_t)tile - ring.cap + 1u;
                    for (int k = 0; k < ring.n_consumers[slot]; ++k) {
                        const uint32_t* f = ring.consumed + ((size_t)k * kDySlots + slot) * ring.cap + e;
                        if (!flag_reached(f, want)) {
                            if (pend_flag) { bulk_wait_all<0>(); flush(); }     // never sit on an unannounced tile while waiting
                            flag_wait_ge(f, want);
                        }
                    }
                    fence_proxy_async_all();
                }
                bulk_s2g(ring.base + ring_slot_off(slot, ring.cap) + (size_t)e * ring_entry_bytes(slot), src, bytes);
                bulk_commit();
                if (ring.use_flags) {
                    if (pend_flag) { bulk_wait_all<1>(); flush(); }
                    pend_flag = ring.ready + (size_t)slot * ring.cap + e;
                    pend_val = (uint32_t)tile + 1u;
                }
            };
            // Two stores in flight (see the forward kernel's stash writer: one SM stores ~32 B/clk at best, 26 when every
            // store is waited for): a slot's panels are released once the NEXT store -- always the other slot's -- has been
            // issued and this one has finished reading shared memory.
            int pending = -1;
            bool pending_free = false;
            auto release = [&](const bool keep_newest) {
                if (pending < 0) return;
                if (keep_newest) bulk_wait_read<1>(); else bulk_wait_read<0>();
                mbar_arrive(&bars->stash_done[pending]);
                if (pending_free) mbar_arrive(&bars->slot_free[pending]);
                pending = -1;
                pending_free = false;
            };
            auto event = [&](const int x, const int tile, const int slot, const uint32_t bytes, const bool last_of_tile) {
                if (tile < p.n_tiles && !ring.use_flags && p.stash_pieces > 1) {
                    // paced (see the forward kernel's stash writer): one panel per request, at most two requests queued in the SM's
                    // bulk-copy unit, so that a weight chunk requested meanwhile is not held up behind 128 KB of stores
                    const uint8_t* src = smem + kBOffH + x * 65536;
                    uint8_t* dst = ring.base + ring_slot_off(slot, ring.cap) + (size_t)((uint32_t)tile % ring.cap) * ring_entry_bytes(slot);
                    const uint32_t piece = 65536u / (uint32_t)p.stash_pieces;
                    for (uint32_t off = 0; off < bytes; off += piece) {
                        bulk_s2g(dst + off, src + off, piece);
                        bulk_commit();
                        bulk_wait_read<1>();
                        if (off == 0 && pending >= 0) {        // every earlier request has been read: the previous slot's panels are free
                            mbar_arrive(&bars->stash_done[pending]);
                            if (pending_free) mbar_arrive(&bars->slot_free[pending]);
                            pending = -1;
                        }
                    }
                    pending = x;
                    pending_free = last_of_tile;
                } else if (tile < p.n_tiles) {
                    emit(tile, slot, smem + kBOffH + x * 65536, bytes);
                    release(true);
                    pending = x;
                    pending_free = last_of_tile;
                } else {
                    release(false);
                    mbar_arrive(&bars->stash_done[x]);
                    if (last_of_tile) mbar_arrive(&bars->slot_free[x]);
                }
            };
            for (int g = 0; 2 * g < my_super; ++g) {
                if (pv) {
                    for (int x = 0; x < 2; ++x) {
                        mbar_wait(&bars->pro_local[x], g & 1);
                        event(x, tile_of(2 * g + x), kDySlotView, 2 * kPanelBytes, false);
                    }
                }
                for (int s = 0; s < p.n_steps; ++s)
                    for (int x = 0; x < 2; ++x) {
                        const uint32_t jx = (uint32_t)(g * p.n_steps + s);
                        mbar_wait(&bars->stash_ready[x], jx & 1);
                        event(x, tile_of(2 * g + x), p.steps[s].slot, 4 * kPanelBytes, s == p.n_steps - 1);
                    }
            }
            release(false);
            bulk_wait_all<0>();
            flush();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kWarpMma) tmem_dealloc2<512>(tmem);
}

// the chain alone (two-kernel form)
template <bool kVis>
__global__ void __launch_bounds__(kBwdThreads, 1) tc_dgrad_kernel(const __grid_constant__ BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    dgrad_role<kVis>(p, smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u), (int)blockIdx.x / 2);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
}

// =================================================================================================
// wgrad kernel
// =================================================================================================
constexpr int kWgThreads = 512;
constexpr int kWgStages = 3;
constexpr uint32_t kWgStageBytes = 65536;
constexpr uint32_t kHalfPanel = 8192;                               // 64 points x 128 bytes
constexpr uint32_t kWgOffDh = kWgStages * kWgStageBytes;             // [64][4] fp32 head-pre gradients of the current stage
constexpr int kRawSlots = 4;                                        // ring of per-stage point inputs for the encoder warps
constexpr uint32_t kWgOffRaw = kWgOffDh + 1024;                     // [kRawSlots][10][64] fp32: z, o(3), d(3), view dir(3)
constexpr uint32_t kWgOffBars = kWgOffRaw + kRawSlots * 10 * 64 * 4;
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
    uint8_t with_head;            // MMA job whose B operand is h8: the reducer warps also form dW_head = d head_pre^T h8
    uint8_t ring_k;               // which of the dY slot's consumers this job is (index into RingCtl::consumed)
    int16_t cta0, n_cta;          // CTAs [cta0, cta0 + n_cta) own this job; CTA cta0 + i takes tiles i, i + n_cta, ...
    float* dw;
    float* db;
    int32_t ld;
    WgSeg seg[3];
};

struct WgParams {
    const uint8_t* act;
    RingCtl ring;                 // source of the gradient panels
    int n_clusters, spread;       // fused launch: total clusters; spread != 0 interleaves chain pairs and weight-gradient clusters
    const float *rays_o, *rays_d, *view_dirs, *z, *sigma, *rgb, *d_sigma, *d_rgb;
    long long n_points;
    int n_samples, n_tiles, n_jobs, pts_degree, view_degree, head_out;
    uint32_t tile_stash_bytes;
    float *head_dw, *head_db;     // destination of the merged head-matrix gradient (with_head jobs)
    int head_ld;
    long long* trace;             // debug: per CTA {job, cycles} (tools/wgrad_balance.py), normally null
    int debug;                    // debug: bit0 encoders write zeros, bit1 no MMAs, bit2 no reducer work
    WgJob jobs[kMaxJobs];
};

struct WgBars {
    uint64_t full[kWgStages], empty[kWgStages], aux_ready[kWgStages], acc_done, acc_free, raw_ready[kRawSlots], raw_free[kRawSlots];
    uint32_t tmem_base;
};

__device__ __forceinline__ float bf16_at(const uint8_t* panel_base, int r, int col) {
    const uint16_t bits = *reinterpret_cast<const uint16_t*>(panel_base + (col >> 6) * kHalfPanel + r * kRowBytes +
                                                             ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
    return __uint_as_float((uint32_t)bits << 16);
}

__device__ __forceinline__ void wgrad_role(const WgParams& p, uint8_t* smem, const int cta) {
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
        for (int i = 0; i < kRawSlots; ++i) { mbar_init(&bars->raw_ready[i], 64); mbar_init(&bars->raw_free[i], 128); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    // one job per CTA for the whole launch: the accumulator of the job's weight matrix stays in TMEM over every tile
    // this CTA owns and is flushed once (fp32 reductions), so the flush traffic is one matrix per CTA
    int jb = -1;
    for (int j = 0; j < p.n_jobs; ++j)
        if (cta >= p.jobs[j].cta0 && cta < p.jobs[j].cta0 + p.jobs[j].n_cta) jb = j;
    const WgJob& job = p.jobs[jb < 0 ? 0 : jb];
    const int part = cta - job.cta0, nparts = job.n_cta;
    const RingCtl& ring = p.ring;
    const long long t_start = clock64();
    const int my_tiles = (jb < 0 || part >= p.n_tiles) ? 0 : (p.n_tiles - part + nparts - 1) / nparts;
    const int n_stages = my_tiles * 2;

    if (warp == 0) {
        // ======================= loader: stashed panels -> smem stages =======================
        if (lane == 0) {
            const int n_a = (p.debug & 64) ? 0 : job.a_panels, n_b = (p.debug & 32) ? 0 : job.b_panels;    // debug bits 5 / 6: skip the B / A copies
            const uint32_t bytes = (uint32_t)(n_a + n_b) * kHalfPanel;
            for (int sg = 0; sg < n_stages; ++sg) {
                const int tile = part + (sg >> 1) * nparts, half = sg & 1;
                const uint32_t stage = sg % kWgStages, round = sg / kWgStages;
                if (round > 0) mbar_wait_sleep(&bars->empty[stage], (round - 1) & 1, 32);
                uint8_t* dst = smem + stage * kWgStageBytes;
                const uint32_t e = (uint32_t)tile % ring.cap;
                if (job.a_panels && ring.use_flags && half == 0) {       // the dgrad pair has finished writing this tile's panels
                    flag_wait_ge(ring.ready + (size_t)job.a_slot * ring.cap + e, (uint32_t)tile + 1u);
                    fence_proxy_async_all();
                }
                if (bytes) mbar_arrive_expect_tx(&bars->full[stage], bytes); else mbar_arrive(&bars->full[stage]);
                const size_t toff = (size_t)tile * p.tile_stash_bytes + (size_t)half * kHalfPanel;
                const uint8_t* dy = ring.base + ring_slot_off(job.a_slot, ring.cap) + (size_t)e * ring_entry_bytes(job.a_slot) + (size_t)half * kHalfPanel;
                for (int j = 0; j < n_a; ++j)
                    bulk_g2s(dst + j * kHalfPanel, dy + j * kPanelBytes, kHalfPanel, &bars->full[stage]);
                for (int j = 0; j < n_b; ++j)
                    bulk_g2s(dst + 32768 + j * kHalfPanel, p.act + toff + (size_t)job.b_slot * 65536 + j * kPanelBytes,
                             kHalfPanel, &bars->full[stage]);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const bool aux = job.b_enc || job.b_venc;
            const int n1 = job.b_panels * 64;
            const int n2 = ((int)job.b_enc + (int)job.b_venc) * 64;
            const int n_mb = job.a_panels / 2;
            for (int sg = 0; sg < n_stages; ++sg) {
                const uint32_t stage = sg % kWgStages;
                mbar_wait(&bars->full[stage], (sg / kWgStages) & 1);
                if (job.a_panels && ring.use_flags && (sg & 1)) {        // both halves of the tile are in shared memory: the ring entry may be reused
                    const int tile = part + (sg >> 1) * nparts;
                    st_release_gpu(ring.consumed + ((size_t)job.ring_k * kDySlots + job.a_slot) * ring.cap + (uint32_t)tile % ring.cap,
                                   (uint32_t)tile + 1u);
                }
                if (job.kind != 0) { mbar_arrive(&bars->empty[stage]); continue; }
                if (aux) mbar_wait(&bars->aux_ready[stage], (sg / kWgStages) & 1);
                tc_fence_after();
                const uint32_t sbase = smem_u32(smem + stage * kWgStageBytes);
                for (int k = 0; k < ((p.debug & 2) ? 0 : 4); ++k) {
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
            if (job.kind == 0 && n_stages > 0) umma_commit(&bars->acc_done);
        }
    } else if (warp == 2 || warp == 3) {
        // ======================= fetchers: per-point inputs of the recomputed encodings -> smem ring =======================
        // (asynchronous 4-byte copies completing on an mbarrier: several stages stay in flight and the encoder warps,
        // whose proxy fence would otherwise wait for their own outstanding loads, never touch global memory)
        if (job.kind == 0 && (job.b_enc || job.b_venc)) {
            const int r = (warp - 2) * 32 + lane;
            float* s_raw = (float*)(smem + kWgOffRaw);
            for (int sg = 0; sg < n_stages; ++sg) {
                const int slot = sg % kRawSlots, round = sg / kRawSlots;
                const int tile = part + (sg >> 1) * nparts, half = sg & 1;
                long long pt = (long long)tile * kTileRows + half * 64 + r;
                if (pt >= p.n_points) pt = p.n_points - 1;     // rows past the end carry zero gradients; any finite input will do
                const int ray = p.n_points < (1LL << 31) ? (int)((unsigned)pt / (unsigned)p.n_samples) : (int)(pt / p.n_samples);
                if (round > 0) mbar_wait_sleep(&bars->raw_free[slot], (round - 1) & 1, 128);
                float* dst = s_raw + slot * 640 + r;
                if (job.b_enc) {
                    cp_async4(dst, p.z + pt);
#pragma unroll
                    for (int c = 0; c < 3; ++c) { cp_async4(dst + (1 + c) * 64, p.rays_o + ray * 3 + c); cp_async4(dst + (4 + c) * 64, p.rays_d + ray * 3 + c); }
                }
                if (job.b_venc) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) cp_async4(dst + (7 + c) * 64, p.view_dirs + ray * 3 + c);
                }
                cp_async_arrive_noinc(&bars->raw_ready[slot]);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ======================= reducers: bias sums, head matrices, TMEM flush =======================
        const int t = (warp - 4) * 32 + lane;            // column owned by this thread
        const int q = warp & 3, colhalf = (warp - 4) >> 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float acc_b = 0.f, acc_hb = 0.f;
        const bool head = job.kind != 0 || job.with_head;   // this job forms a head matrix from its streamed B operand
        const int nh = job.kind == 2 ? 3 : p.head_out;
        // smem byte offsets of column `col`, rows 8g + i (i = 0..7), inside a set of [64 x 64] half panels: + g * 1024
        auto column_offsets = [](int col, uint32_t (&off)[8]) {
            const uint32_t base = (uint32_t)(col >> 6) * kHalfPanel + (uint32_t)(col & 7) * 2u, chunk = (uint32_t)(col & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) off[i] = base + (uint32_t)i * kRowBytes + ((chunk ^ (uint32_t)i) << 4);
        };
        auto bf16_ld = [](const uint8_t* a) { return __uint_as_float((uint32_t)*reinterpret_cast<const uint16_t*>(a) << 16); };
        const bool do_bias = job.kind == 0 && job.db != nullptr && t < job.a_panels * 64 && !(p.debug & 4);
        uint32_t off_a[8], off_b[8];
        column_offsets(t, off_a);
        // head matrix: thread owns column hcol of the B operand and the row groups [g0, g0 + ng)
        const int ncol = job.b_panels * 64;                 // 256 (h8) or 128 (hv)
        const int hcol = ncol > 0 ? t % ncol : 0;
        const int ng = ncol == 128 ? 4 : 8, g0 = ncol == 128 ? (t >> 7) * 4 : 0;
        column_offsets(hcol, off_b);
        // head-pre gradients of one stage (64 points): threads 0..63 fetch the raw inputs one stage ahead (loads only, so
        // that nothing waits on them before the next iteration) and finish the arithmetic when the stage comes up
        struct RawHead { float sig, dsig, r[3], dr[3]; };
        auto fetch_dh = [&](int sg, RawHead& w) {
            w.sig = 0.f; w.dsig = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { w.r[c] = 0.f; w.dr[c] = 0.f; }
            if (!head || t >= 64 || sg >= n_stages) return;
            const int tile = part + (sg >> 1) * nparts, half = sg & 1;
            const long long pt = (long long)tile * kTileRows + half * 64 + t;
            if (pt >= p.n_points) return;
            if (job.kind == 2 || p.head_out == 4) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { w.r[c] = __ldg(p.rgb + pt * 3 + c); w.dr[c] = __ldg(p.d_rgb + pt * 3 + c); }
            }
            if (job.kind != 2) { w.sig = __ldg(p.sigma + pt); w.dsig = __ldg(p.d_sigma + pt); }
        };
        auto finish_dh = [&](const RawHead& w) {
            float g[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) g[c] = w.dr[c] * w.r[c] * (1.f - w.r[c]);             // sigmoid'
            if (job.kind == 2) return make_float4(g[0], g[1], g[2], 0.f);
            const float ds = w.sig > 0.f ? w.dsig : 0.f;                                       // relu'
            return p.head_out == 4 ? make_float4(ds, g[0], g[1], g[2]) : make_float4(ds, 0.f, 0.f, 0.f);
        };
        RawHead raw;
        fetch_dh(0, raw);
        for (int sg = 0; sg < n_stages; ++sg) {
            const uint32_t stage = sg % kWgStages;
            const uint8_t* sb = smem + stage * kWgStageBytes;
            if (head) {
                if (t < 64) *reinterpret_cast<float4*>(s_dh + t * 4) = finish_dh(raw);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                fetch_dh(sg + 1, raw);       // in flight while this stage is reduced
            }
            mbar_wait(&bars->full[stage], (sg / kWgStages) & 1);
            if (do_bias) {
#pragma unroll
                for (int g = 0; g < 8; ++g)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc_b += bf16_ld(sb + off_a[i] + g * 1024);
            }
            if (head) {
                const uint8_t* hb = sb + 32768 + g0 * 1024;
                const float* dh = s_dh + g0 * 32;
                if (nh == 1) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[0] = fmaf(dh[(g * 8 + i) * 4], bf16_ld(hb + off_b[i] + g * 1024), acc[0]);
                } else {
                    for (int g = 0; g < ng; ++g)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float x = bf16_ld(hb + off_b[i] + g * 1024);
                            const float4 d = *reinterpret_cast<const float4*>(dh + (g * 8 + i) * 4);
                            acc[0] = fmaf(d.x, x, acc[0]); acc[1] = fmaf(d.y, x, acc[1]);
                            acc[2] = fmaf(d.z, x, acc[2]); acc[3] = fmaf(d.w, x, acc[3]);
                        }
                }
                if (t < nh) {
                    for (int r = 0; r < 64; ++r) acc_hb += s_dh[r * 4 + t];
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->empty[stage]);
        }
        // ---- job results ----
        if (n_stages > 0) {
            if (head) {
                float* hdw = job.kind == 0 ? p.head_dw : job.dw;
                float* hdb = job.kind == 0 ? p.head_db : job.db;
                const int hld = job.kind == 0 ? p.head_ld : job.ld;
                if (ncol > 0)
                    for (int h = 0; h < nh; ++h) atomicAdd(hdw + (size_t)h * hld + hcol, acc[h]);
                if (t < nh && hdb != nullptr) atomicAdd(hdb + t, acc_hb);
            }
            if (job.kind == 0) {
                if (job.db != nullptr && t < job.a_panels * 64) atomicAdd(job.db + t, acc_b);
                mbar_wait(&bars->acc_done, 0);
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
            }
        }
    } else if (warp >= 12 && warp < 16) {     // (the fused launch has 23 warps: the rest idle through the weight-gradient role)
        // ======================= encoders: recomputed encoding operands =======================
        const int e = (warp - 12) * 32 + lane;
        const int r = e >> 1, hf = e & 1;
        const bool aux = job.kind == 0 && (job.b_enc || job.b_venc);
        if (aux) {
            const float* s_raw = (const float*)(smem + kWgOffRaw);
            const int venc = 3 * (1 + 2 * p.view_degree);
            for (int sg = 0; sg < n_stages; ++sg) {
                const uint32_t stage = sg % kWgStages, round = sg / kWgStages;
                const int slot = sg % kRawSlots;
                mbar_wait(&bars->raw_ready[slot], (sg / kRawSlots) & 1);
                const float* src = s_raw + slot * 640 + r;
                float x[3] = {0.f, 0.f, 0.f}, vdir[3] = {0.f, 0.f, 0.f};
                if (job.b_enc) {
                    const float zz = src[0];
#pragma unroll
                    for (int c = 0; c < 3; ++c) x[c] = fmaf(src[(4 + c) * 64], zz, src[(1 + c) * 64]);
                }
                if (job.b_venc) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) vdir[c] = src[(7 + c) * 64];
                }
                mbar_arrive(&bars->raw_free[slot]);
                uint32_t pe[16], pv[16];    // this thread's 32 columns, packed bf16 pairs
                if (p.debug & 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { pe[i] = 0u; pv[i] = 0u; }
                } else
                if (job.b_enc) {
                    if (hf == 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pe[i] = pack_bf16(encode_element(x, p.pts_degree, 2 * i), encode_element(x, p.pts_degree, 2 * i + 1));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pe[i] = pack_bf16(encode_element(x, p.pts_degree, 32 + 2 * i), encode_element(x, p.pts_degree, 33 + 2 * i));
                    }
                }
                if (job.b_venc && !(p.debug & 1)) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a = (hf == 0 && 2 * i < venc) ? encode_element(vdir, p.view_degree, 2 * i) : 0.f;
                        const float b = (hf == 0 && 2 * i + 1 < venc) ? encode_element(vdir, p.view_degree, 2 * i + 1) : 0.f;
                        pv[i] = pack_bf16(a, b);
                    }
                }
                if (round > 0) mbar_wait(&bars->empty[stage], (round - 1) & 1);
                uint8_t* dst = smem + stage * kWgStageBytes + (uint32_t)job.aux_off * 1024u;
                if (job.b_enc) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + swz_offset(r, hf * 4 + c)) = make_uint4(pe[4 * c], pe[4 * c + 1], pe[4 * c + 2], pe[4 * c + 3]);
                    dst += kHalfPanel;
                }
                if (job.b_venc) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + swz_offset(r, hf * 4 + c)) = make_uint4(pv[4 * c], pv[4 * c + 1], pv[4 * c + 2], pv[4 * c + 3]);
                }
                fence_async_smem();
                mbar_arrive(&bars->aux_ready[stage]);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (p.trace && threadIdx.x == 0) { p.trace[2 * cta] = jb; p.trace[2 * cta + 1] = clock64() - t_start; }
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// the weight-gradient jobs alone (two-kernel form)
__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    wgrad_role(p, smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u), (int)blockIdx.x);   // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
}

// Fused form: ONE launch in which CTA pairs [0, n_pairs) run the dgrad chain and the remaining CTAs run the weight-gradient
// jobs on the gradient panels the pairs publish through the ring -- producer and consumer are co-resident by construction
// (one CTA per SM, grid <= SM count), so the hand-off cannot starve, also when a profiler serialises kernels.
constexpr uint32_t kFusedSmem = kBwdSmem > kWgSmem ? kBwdSmem : kWgSmem;
__global__ void __launch_bounds__(kBwdThreads, 1) tc_backward_kernel(const __grid_constant__ BwdParams bp, const __grid_constant__ WgParams wp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // role of a cluster: chain pairs first, or spread evenly over the launch (the clusters of a launch are dealt round-robin
    // to the GPCs, so spreading gives every GPC both kinds of traffic)
    const int c = (int)blockIdx.x / 2, nc = wp.n_clusters;
    const int before = wp.spread ? (int)((long long)c * bp.n_pairs / nc) : (c < bp.n_pairs ? c : bp.n_pairs);
    const bool chain = wp.spread ? (int)((long long)(c + 1) * bp.n_pairs / nc) > before : c < bp.n_pairs;
    if (chain) {
        if (wp.debug & 16) return;                        // debug bit 4: weight-gradient jobs alone (timing experiments)
        dgrad_role<false>(bp, smem, before);
    } else {
        if (wp.debug & 8) return;                         // debug bit 3: chain alone
        wgrad_role(wp, smem, 2 * (c - before) + ((int)blockIdx.x & 1));
    }
}

// =================================================================================================
// gradients of the two matrices behind the merged view branch (tc_plan.cuh)
// =================================================================================================
// With G = dY_v^T h8 [128 x 256] and s = column sums of dY_v [128] (left in the workspace by the view layer's job):
//   dW_view[:, :256] += G W_feat^T + s b_feat^T      dW_view[:, 256:] += the job's encoding columns      db_view += s
//   dW_feat          += W_view[:, :256]^T G          db_feat += W_view[:, :256]^T s
// fp32 on the CUDA cores, 17 MFLOP per call: blocks 0..63 the first product (16 x 32 output tiles), 64..191 the second,
// 192.. the copies and the two vectors.
constexpr int kUnmergeBlocks = 64 + 128 + 5;
__global__ void __launch_bounds__(256) tc_unmerge_grads_kernel(const float* __restrict__ merged, const float* __restrict__ w_view,
                                                               const float* __restrict__ w_feat, const float* __restrict__ b_feat,
                                                               float* __restrict__ dw_view, float* __restrict__ db_view,
                                                               float* __restrict__ dw_feat, float* __restrict__ db_feat, int view_in) {
    __shared__ float sa[32][33], sb[32][33];
    const float* s_vec = merged + 128 * view_in;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // thread: column tx, rows ty and ty + 8
    int b = blockIdx.x;
    float acc[2] = {0.f, 0.f};
    if (b < 64) {
        const int o0 = (b >> 3) * 16, k0 = (b & 7) * 32;         // out[o][k] = sum_j G[o][j] W_feat[k][j]
        for (int j0 = 0; j0 < 256; j0 += 32) {
#pragma unroll
            for (int i = 0; i < 2; ++i) sa[ty + 8 * i][tx] = merged[(size_t)(o0 + ty + 8 * i) * view_in + j0 + tx];
#pragma unroll
            for (int i = 0; i < 4; ++i) sb[ty + 8 * i][tx] = w_feat[(size_t)(k0 + ty + 8 * i) * 256 + j0 + tx];
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float w = sb[tx][c];
                acc[0] = fmaf(sa[ty][c], w, acc[0]);
                acc[1] = fmaf(sa[ty + 8][c], w, acc[1]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int o = o0 + ty + 8 * i, k = k0 + tx;
            dw_view[(size_t)o * view_in + k] += acc[i] + s_vec[o] * b_feat[k];
        }
        return;
    }
    b -= 64;
    if (b < 128) {
        const int k0 = (b >> 3) * 16, j0 = (b & 7) * 32;         // out[k][j] = sum_o W_view[o][k] G[o][j]
        for (int o0 = 0; o0 < 128; o0 += 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (tx < 16) sa[ty + 8 * i][tx] = w_view[(size_t)(o0 + ty + 8 * i) * view_in + k0 + tx];      // [o][k]
                sb[ty + 8 * i][tx] = merged[(size_t)(o0 + ty + 8 * i) * view_in + j0 + tx];                   // [o][j]
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float g = sb[c][tx];
                acc[0] = fmaf(sa[c][ty], g, acc[0]);
                acc[1] = fmaf(sa[c][ty + 8], g, acc[1]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) dw_feat[(size_t)(k0 + ty + 8 * i) * 256 + j0 + tx] += acc[i];
        return;
    }
    b -= 128;
    if (b < 4) {                                                 // the encoding columns of dW_view, rows 32 b .. 32 b + 31
        const int extra = view_in - 256;
        for (int i = threadIdx.x; i < 32 * extra; i += 256) {
            const int o = 32 * b + i / extra, k = 256 + i % extra;
            dw_view[(size_t)o * view_in + k] += merged[(size_t)o * view_in + k];
        }
        return;
    }
    if (threadIdx.x < 128) db_view[threadIdx.x] += s_vec[threadIdx.x];
    float a = 0.f;                                               // db_feat[k] = sum_o W_view[o][k] s[o]
    for (int o = 0; o < 128; ++o) a = fmaf(w_view[(size_t)o * view_in + threadIdx.x], s_vec[o], a);
    db_feat[threadIdx.x] += a;
}

static long long* g_wg_trace = nullptr;
static cudaEvent_t g_split_event = nullptr;   // measurement: recorded between the dgrad and the wgrad launch (bench.py)
static int g_wg_debug = 0;   // set by snerfdbg_set_wgrad_trace (debug only)

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
    const bool fused = bwd_fused();
    // fused: the chain gets bwd_dgrad_pairs() CTA pairs (fewer for small inputs), the weight-gradient jobs the other SMs
    const int n_super = (w.n_tiles + 1) / 2;
    int n_pairs = fused ? bwd_dgrad_pairs() : num_sms() / 2;
    if (n_pairs > (n_super + 1) / 2) n_pairs = (n_super + 1) / 2;
    if (fused && 2 * n_pairs > num_sms() - 16) n_pairs = (num_sms() - 16) / 2;
    const int wg_ctas = fused ? (num_sms() - 2 * n_pairs) / 2 * 2 : num_sms();

    RingCtl ring{};
    ring.base = wsb + w.dy;
    ring.ready = (uint32_t*)(wsb + w.flags);
    ring.consumed = ring.ready + (size_t)kDySlots * w.ring_cap;
    ring.cap = w.ring_cap;
    ring.use_flags = fused ? 1 : 0;

    // ---- (1) dgrad chain ----
    BwdParams bp{};
    bp.packed = (const uint8_t*)packed; bp.act = wsb + w.act; bp.bits = wsb + w.bits;
    bp.sigma = sigma; bp.rgb = rgb; bp.d_sigma = d_sigma; bp.d_rgb = d_rgb;
    bp.w_rgb = m.has_view ? prm[SNERF_P_RGB_W] : nullptr;
    {
        static int pieces = -1;
        if (pieces < 0) { const char* e = getenv("SNERF_STASH_PIECES"); pieces = e ? atoi(e) : 4; }
        bp.stash_pieces = (pieces == 2 || pieces == 4 || pieces == 8 || pieces == 16) ? pieces : 1;
    }
    if (flags & SNERF_FLAG_VIS_GRAD) {
        SNERF_REQUIRE((flags & SNERF_FLAG_VIS_HEAD) && m.has_view, "mlp_backward: SNERF_FLAG_VIS_GRAD on the tensor path needs the forward's SNERF_FLAG_VIS_HEAD");
        SNERF_REQUIRE(!fused, "mlp_backward: SNERF_FLAG_VIS_GRAD is not built for the one-launch form (unset SNERF_BWD_RING)");
        bp.vis_extra = (const float*)(wsb + w.vis_extra);
    }
    bp.n_points = P; bp.n_tiles = w.n_tiles; bp.n_steps = pl.n_bwd; bp.has_view = m.has_view ? 1 : 0;
    bp.n_pairs = n_pairs;
    bp.tile_stash_bytes = pl.tile_stash_bytes;
    for (int s = 0; s < pl.n_bwd; ++s) bp.steps[s] = pl.bwd[s];
    static bool attr = false;
    if (!attr) {
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
        SNERF_CUDA_OK(cudaFuncSetAttribute(tc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
        attr = true;
    }

    // ---- (2) wgrad ----
    WgParams wp{};
    wp.act = wsb + w.act;
    wp.n_clusters = n_pairs + wg_ctas / 2;
    {
        static int spread = -1;
        if (spread < 0) { const char* e = getenv("SNERF_BWD_SPREAD"); spread = e ? atoi(e) : 1; }
        wp.spread = spread;
    }
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
    float* merged = (float*)(wsb + w.merged_grad);       // [128 x view_in] then [128]
    if (m.has_view) {
        // G = dY_v^T [h8 | E(bands >= trunk_degree) | PE(view dir)] and the column sums of dY_v go to the workspace; the
        // gradients of W_view, b_view, W_feat, b_feat follow from them (tc_unmerge_grads_kernel below)
        const bool hi = m.enc_hi > 0;
        WgJob v = mma_job(kDySlotView, 2, 7, 4, hi, true, merged, m.view_in, merged + 128 * m.view_in);
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
        r.kind = 2; r.b_slot = kSlotHv; r.b_panels = 2; r.dw = grads[SNERF_P_RGB_W]; r.ld = m.view_width; r.db = grads[SNERF_P_RGB_B];
        wp.jobs[nj++] = r;
    }
    wp.trace = g_wg_trace; wp.debug = g_wg_debug;
    wp.head_dw = grads[SNERF_P_HEAD_W]; wp.head_db = grads[SNERF_P_HEAD_B]; wp.head_ld = m.width;
    if (m.has_view) {
        // dW_head = d head_pre^T h8 rides on the view job, whose streamed B operand is the same h8
        for (int j = 0; j < nj; ++j)
            if (wp.jobs[j].kind == 0 && wp.jobs[j].a_slot == kDySlotView) wp.jobs[j].with_head = 1;
    } else {
        WgJob h{};
        h.kind = 1; h.b_slot = 7; h.b_panels = 4; h.dw = grads[SNERF_P_HEAD_W]; h.ld = m.width; h.db = grads[SNERF_P_HEAD_B];
        wp.jobs[nj++] = h;
    }
    wp.n_jobs = nj;
    // ring consumers: a dY slot is read by one job, the skip layer's by two (hidden part and encoding part)
    for (int j = 0; j < nj; ++j)
        if (wp.jobs[j].kind == 0 && wp.jobs[j].a_panels) wp.jobs[j].ring_k = ring.n_consumers[wp.jobs[j].a_slot]++;
    for (int sl = 0; sl < kDySlots; ++sl)
        SNERF_REQUIRE(ring.n_consumers[sl] <= kRingConsumers, "mlp_backward: too many consumers of gradient slot %d", sl);
    bp.ring = ring;
    wp.ring = ring;
    // CTAs per job in proportion to the measured cost of the job per tile
    {
        const int G = wg_ctas;
        double cost[kMaxJobs], total = 0.0;
        for (int j = 0; j < nj; ++j) {
            const WgJob& jb = wp.jobs[j];
            // measured with tools/wgrad_balance.py (cycles of work per tile, relative): streaming dominates, the
            // recomputed encodings and the CUDA-core head matrices add their own latency
            // (re-measured at the end of round 2, work per job = sum of its CTAs' cycles, streaming job = 100: layer 0 147-150, skip layer's
            // hidden part 103-109, its encoding part 135-140, view job + head 184 / 275 with the point encodings, head matrix alone 160,
            // rgb matrix 116-118)
            cost[j] = (jb.a_panels + jb.b_panels) * 12.5;                             // streaming job: 100 for 4 + 4 panels
            if (jb.kind == 0 && jb.b_panels == 4 && jb.seg[0].dst_col > 0) cost[j] = 105.0;   // skip layer, hidden part (offset columns)
            if (jb.kind == 0 && jb.b_panels == 0) cost[j] = jb.db != nullptr ? 147.0 : 137.0; // dY x encoding only (encoder-bound): layer 0 / skip layer
            if (jb.kind == 0 && jb.a_panels == 2) cost[j] = jb.b_enc ? 219.0 : 128.0; // view job: recomputed view-dir (+ point) encoding
            if (jb.with_head) cost[j] += 56.0;
            if (jb.kind == 1) cost[j] = 160.0;
            if (jb.kind == 2) cost[j] = 117.0;
            total += cost[j];
        }
        int n[kMaxJobs], used = 0;
        for (int j = 0; j < nj; ++j) {
            n[j] = (int)(G * cost[j] / total);
            if (n[j] < 1) n[j] = 1;
            used += n[j];
        }
        while (used < G) {   // hand the remaining CTAs to the most loaded jobs
            int best = 0;
            for (int j = 1; j < nj; ++j)
                if (cost[j] / n[j] > cost[best] / n[best]) best = j;
            ++n[best]; ++used;
        }
        while (used > G) {
            int best = -1;
            for (int j = 0; j < nj; ++j)
                if (n[j] > 1 && (best < 0 || cost[j] / n[j] < cost[best] / n[best])) best = j;
            --n[best]; --used;
        }
        int c0 = 0;
        for (int j = 0; j < nj; ++j) { wp.jobs[j].cta0 = (int16_t)c0; wp.jobs[j].n_cta = (int16_t)n[j]; c0 += n[j]; }
    }
    auto unmerge = [&]() -> int {
        if (!m.has_view) return SNERF_OK;
        tc_unmerge_grads_kernel<<<kUnmergeBlocks, 256, 0, st>>>(merged, prm[SNERF_P_VIEW_W], prm[SNERF_P_FEAT_W], prm[SNERF_P_FEAT_B],
                                                               grads[SNERF_P_VIEW_W], grads[SNERF_P_VIEW_B], grads[SNERF_P_FEAT_W],
                                                               grads[SNERF_P_FEAT_B], m.view_in);
        SNERF_LAUNCH_OK("tc_unmerge_grads_kernel");
        return SNERF_OK;
    };
    if (m.has_view) SNERF_CUDA_OK(cudaMemsetAsync(merged, 0, (size_t)(128 * m.view_in + 128) * sizeof(float), st));
    if (fused) {
        if (wp.debug & 24) { bp.ring.use_flags = 0; wp.ring.use_flags = 0; }     // one role alone: no hand-off (results are garbage)
        SNERF_CUDA_OK(cudaMemsetAsync(ring.ready, 0, (size_t)(1 + kRingConsumers) * kDySlots * ring.cap * sizeof(uint32_t), st));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * n_pairs + wg_ctas));
        cfg.blockDim = dim3((unsigned)kBwdThreads);
        cfg.dynamicSmemBytes = kFusedSmem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        SNERF_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_backward_kernel, bp, wp));
        return unmerge();
    }
    if (bp.vis_extra) SNERF_CUDA_OK(launch_clustered(tc_dgrad_kernel<true>, 2 * n_pairs, kBwdThreads, kBwdSmem, st, bp));
    else SNERF_CUDA_OK(launch_clustered(tc_dgrad_kernel<false>, 2 * n_pairs, kBwdThreads, kBwdSmem, st, bp));
    if (g_split_event) cudaEventRecord(g_split_event, st);
    tc_wgrad_kernel<<<num_sms(), kWgThreads, kWgSmem, st>>>(wp);
    SNERF_LAUNCH_OK("tc_wgrad_kernel");
    return unmerge();
}

}  // namespace snerf

// measurement hook (include/simplenerf_b200.h): a cudaEvent_t recorded on the launch stream between the dgrad and the wgrad launch
extern "C" void snerf_set_backward_split_event(void* event) { snerf::g_split_event = (cudaEvent_t)event; }
#ifdef SNERF_DEBUG
extern "C" void snerfdbg_set_wgrad_trace(long long* device_buffer) { snerf::g_wg_trace = device_buffer; }
extern "C" void snerfdbg_set_wgrad_debug(int bits) { snerf::g_wg_debug = bits; }
#endif  // SNERF_DEBUG
