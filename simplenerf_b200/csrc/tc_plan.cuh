// Plan shared by the tensor-path kernels: the sequence of GEMM steps of one MLP (forward chain and
// dgrad chain), the layout of the packed bf16 weight image, the activation / gradient stash and the workspace.
#pragma once
#include "common.cuh"
#include <stdlib.h>

#include "tc_common.cuh"

namespace snerf {
using namespace tc;

// ------------------------------------------------------------------------------------------------
// plan: GEMM steps, packed-weight layout
// ------------------------------------------------------------------------------------------------
constexpr int kCluster = 2;   // CTAs sharing one weight stream: every weight chunk is fetched once per cluster and multicast
constexpr int kStages = 3;
constexpr int kStageBytes = 32768;
constexpr int kMaxSteps = 10;
constexpr int kMaxChunks = 5;
constexpr int kPanelE = 4;   // panel id of the encoding buffer
constexpr int kPanelP = 5;   // (backward) first panel of the prologue buffer

enum EpiKind : uint8_t { EPI_RELU = 0, EPI_RELU_HEAD4, EPI_LINEAR, EPI_VIEW, BWD_LINEAR, BWD_MASK };

struct TcStep {
    uint32_t w_off;               // byte offset of the first weight chunk in the packed image
    uint16_t n_rows;              // N of the MMA (256 or 128)
    uint8_t n_chunks;
    uint8_t kind;
    uint8_t panel[kMaxChunks];    // A panel per chunk
    uint8_t ksteps[kMaxChunks];   // 16-wide K steps per chunk
    uint8_t bias_row;             // row of the smem bias table
    uint8_t slot;                 // stash slot of the panels this step writes
    uint8_t last_e_use;           // this step is the last reader of the aux panels (E / P) within a tile
    uint8_t mask_slot;            // (backward) activation-stash slot whose sign masks this step's output
};

struct PackChunk {
    const float* src;
    int32_t ld;
    int16_t n_rows, k_lo, k_hi, col0, row0, transposed;
    int16_t src_rows;             // rows the source really has from row0 on (the rest of n_rows is zero-filled)
    int16_t dst_row0;             // first row inside the destination chunk image (a multiple of 8: swizzle phase)
    uint32_t dst_off;
};
constexpr int kMaxPack = 96;

struct TcPlan {
    int n_fwd, n_bwd;
    TcStep fwd[kMaxSteps], bwd[kMaxSteps];
    int n_pack;
    PackChunk pack[kMaxPack];
    uint32_t packed_bytes;
    uint32_t tile_stash_bytes;    // activation (and dY) stash bytes per 128-point tile
};

static inline void add_chunk(TcPlan& pl, TcStep& st, int panel, int ksteps, const float* src, int ld, int n_rows, int k_lo, int k_hi,
                      int col0, int row0, bool transposed) {
    PackChunk& c = pl.pack[pl.n_pack++];
    c.src = src; c.ld = ld; c.n_rows = (int16_t)n_rows; c.k_lo = (int16_t)k_lo; c.k_hi = (int16_t)k_hi;
    c.col0 = (int16_t)col0; c.row0 = (int16_t)row0; c.transposed = transposed ? 1 : 0;
    c.src_rows = (int16_t)n_rows; c.dst_row0 = 0;
    c.dst_off = pl.packed_bytes;
    if (st.n_chunks == 0) st.w_off = pl.packed_bytes;
    st.panel[st.n_chunks] = (uint8_t)panel;
    st.ksteps[st.n_chunks] = (uint8_t)ksteps;
    st.n_chunks++;
    pl.packed_bytes += (uint32_t)n_rows * kRowBytes;
}

// Merged view branch.  feature_linear has no activation (src/models/SimpleNeRF01.py:691-697), so the view layer's feature
// part is one linear map of the last trunk activation:  W_view[:, :256] (W_feat h + b_feat) = W_vf h + W_view[:, :256] b_feat
// with W_vf = W_view[:, :256] W_feat  [128 x 256].  The tensor path multiplies by W_vf directly (formed in fp32 at pack
// time, rounded to bf16 once): the feature step, its dgrad step, its activation / gradient panels (1 KB of the 10.2 KB a
// point moved through HBM) and its weight-gradient job are gone.  The gradients of the two original matrices follow from
// G = dY_v^T h (accumulated by the view layer's weight-gradient job) by two small fp32 products (tc_unmerge_grads_kernel).
// Sigma head on the tensor core (MLPs with a view branch).  sigma_pre = w_head . h8 + b_head reads the same operand as the view
// step (the last trunk activation), so the head row rides that step as one more B row: N = 128 view columns + 16 (row 128 =
// pts_output_linear.weight[0] in bf16, rows 129..143 zero; N % 16 == 0 for M = 256), and sigma_pre appears as accumulator
// column 128.  An M256 pair MMA is paced by its A-operand fetch, so N = 144 costs what N = 128 did, while the last trunk
// layer's epilogue becomes the plain packed one (it was the slowest step of the chain: fp32 bias / ReLU / dot product per
// value, ~3 900 cycles against ~2 000).  The backward pass always treated the head that way (d h8 += d sigma_pre w_head as a
// bf16 chunk of the first dgrad step, dW_head from the bf16 stash).
constexpr int kViewN = 144;                                  // B rows of the view step
constexpr int kSigmaCol = 128;                               // accumulator column of sigma_pre
constexpr int kSlotHv = 8;                                   // activation-stash slot of the view layer's output
constexpr int kDySlotView = 9;                               // gradient-ring slot of dY_v (2 panels)
constexpr size_t kMergedWeightBytes = (128 * 256 + 128) * sizeof(float);   // W_vf and W_view[:, :256] b_feat, fp32, appended to the packed image

// prm may be null (layout only); wvf = the fp32 W_vf appended to the packed image (null: layout only)
static inline TcPlan build_plan(const snerf_mlp_desc& d, const float* const* prm, const float* wvf = nullptr) {
    const MlpDims m(d);
    TcPlan pl{};
    auto P = [&](int i) -> const float* { return prm ? prm[i] : nullptr; };
    // ---- forward steps ----
    for (int l = 0; l < m.depth; ++l) {
        TcStep& st = pl.fwd[pl.n_fwd++];
        st.n_rows = 256;
        st.kind = EPI_RELU;
        st.bias_row = (uint8_t)l;
        st.slot = (uint8_t)l;
        const int fan_in = m.trunk_fan_in(l);
        if (l == 0) {
            add_chunk(pl, st, kPanelE, ceil_div(m.trunk_in, 16), P(0), fan_in, 256, 0, m.trunk_in, 0, 0, false);
        } else {
            int col = 0;
            if (l - 1 == m.skip_layer) {
                add_chunk(pl, st, kPanelE, ceil_div(m.trunk_in, 16), P(2 * l), fan_in, 256, 0, m.trunk_in, 0, 0, false);
                col = m.trunk_in;
            }
            for (int j = 0; j < 4; ++j) add_chunk(pl, st, j, 4, P(2 * l), fan_in, 256, 0, 64, col + 64 * j, 0, false);
        }
    }
    pl.fwd[m.depth - 1].kind = m.has_view ? EPI_RELU : EPI_RELU_HEAD4;      // with a view branch the sigma head rides the view step
    pl.fwd[m.skip_layer + 1].last_e_use = 1;
    if (m.has_view) {
        TcStep& vw = pl.fwd[pl.n_fwd++];       // hv_pre = W_vf h + [per-ray bias] (+ W_view[:, enc part] E)
        vw.n_rows = kViewN; vw.kind = EPI_VIEW; vw.slot = kSlotHv;
        // a chunk image of kViewN rows: rows 0..127 from `src`, rows 128..143 by a second entry (the sigma row or zeros)
        auto add_view_chunk = [&](int panel, const float* src, int ld, int k_lo, int k_hi, int col0, const float* row128, int col128) {
            add_chunk(pl, vw, panel, 4, src, ld, kViewN, k_lo, k_hi, col0, 0, false);
            pl.pack[pl.n_pack - 1].n_rows = 128;
            pl.pack[pl.n_pack - 1].src_rows = 128;
            PackChunk& ex = pl.pack[pl.n_pack];
            ex = pl.pack[pl.n_pack - 1];
            ex.src = row128; ex.ld = m.width; ex.col0 = (int16_t)col128; ex.row0 = 0; ex.k_lo = 0; ex.k_hi = 64;
            ex.n_rows = kViewN - 128; ex.src_rows = row128 ? 1 : 0; ex.dst_row0 = kSigmaCol;
            ++pl.n_pack;
        };
        // row 128 = pts_output_linear.weight[0, 64 j .. 64 j + 63]  (layout-only plans pass a null table: nothing is read)
        for (int j = 0; j < 4; ++j) add_view_chunk(j, wvf, m.width, 0, 64, 64 * j, prm ? P(SNERF_P_HEAD_W) : (const float*)nullptr, 64 * j);
        if (m.enc_hi > 0) {   // points-augmentation: encoding bands trunk_degree.. feed the view layer (:633); no share in sigma
            add_view_chunk(kPanelE, P(SNERF_P_VIEW_W), m.view_in, m.trunk_in, m.enc, m.width, nullptr, 0);
            pl.fwd[m.skip_layer + 1].last_e_use = 0;
            vw.last_e_use = 1;
        }
    }
    pl.tile_stash_bytes = (uint32_t)(m.has_view ? 8 * 65536 + 32768 : 8 * 65536);

    // ---- backward (dgrad) steps: B operand = W^T chunks [256 in-features x 64 out-features] ----
    // dY_l = gradient w.r.t. the pre-activation of trunk layer l; gradient slot l.  Slot 9 = d hv_pre (slot 8 is unused).
    {
        TcStep& fg = pl.bwd[pl.n_bwd++];    // d h8 = dY_v W_vf + d head_pre W_head, masked by h8 > 0
        fg.n_rows = 256; fg.kind = BWD_MASK; fg.mask_slot = 7; fg.slot = 7; fg.last_e_use = 1;
        if (m.has_view)
            for (int c = 0; c < 2; ++c) add_chunk(pl, fg, kPanelP + c, 4, wvf, m.width, 256, 0, 64, 0, 64 * c, true);
        add_chunk(pl, fg, kPanelP + 2, 1, P(SNERF_P_HEAD_W), m.width, 256, 0, m.head_out, 0, 0, true);
    }
    for (int l = m.depth - 1; l >= 1; --l) {   // d h_l = dY_l W_l[:, hidden part], masked by h_l > 0
        TcStep& st = pl.bwd[pl.n_bwd++];
        st.n_rows = 256; st.kind = BWD_MASK; st.mask_slot = (uint8_t)(l - 1); st.slot = (uint8_t)(l - 1);
        const int hcol0 = (l - 1 == m.skip_layer) ? m.trunk_in : 0;
        for (int c = 0; c < 4; ++c) add_chunk(pl, st, c, 4, P(2 * l), m.trunk_fan_in(l), 256, 0, 64, hcol0, 64 * c, true);
    }
    return pl;
}


// ------------------------------------------------------------------------------------------------
// positional encoding helpers
// ------------------------------------------------------------------------------------------------
// enc[0..63]: x(3), then per band sin(3), cos(3); enc[63] = 0 (pad).  :537-551
// Tensor-path version: the argument is reduced exactly (x/2pi scaled by the power of two, minus its nearest integer)
// and evaluated with the MUFU units; absolute error < 1e-4 at band 9, far below the bf16 rounding of the operand.
__device__ __forceinline__ void encode_point(const float x[3], int degree, float* enc) {
    enc[0] = x[0]; enc[1] = x[1]; enc[2] = x[2];
    const float turns[3] = {x[0] * 0.15915494309189535f, x[1] * 0.15915494309189535f, x[2] * 0.15915494309189535f};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (k < degree) {
            const float scale = (float)(1 << k);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float u = turns[c] * scale;
                const float f = u - rintf(u);
                float sn, cs;
                __sincosf(f * 6.283185307179586f, &sn, &cs);
                enc[3 + 6 * k + c] = sn;
                enc[6 + 6 * k + c] = cs;
            }
        }
    }
}

// one element of encode_point (same arithmetic, so a recomputed operand equals the forward one bit for bit)
__device__ __forceinline__ float encode_element(const float x[3], int degree, int idx) {
    if (idx < 3) return idx == 0 ? x[0] : (idx == 1 ? x[1] : x[2]);
    const int j = idx - 3, k = j / 6, rem = j - 6 * k;
    const int c = rem >= 3 ? rem - 3 : rem;
    const float xc = c == 0 ? x[0] : (c == 1 ? x[1] : x[2]);
    const float u = (xc * 0.15915494309189535f) * (float)(1 << (k & 15));
    const float f = u - rintf(u);
    const float v = rem >= 3 ? __cosf(f * 6.283185307179586f) : __sinf(f * 6.283185307179586f);
    return k < degree ? v : 0.f;     // a select, not a branch: the elements of a row interleave freely
}

// accurate variant (fp32 consumers: the per-ray view-direction bias)
__device__ __forceinline__ void encode_point_accurate(const float x[3], int degree, float* enc) {
    enc[0] = x[0]; enc[1] = x[1]; enc[2] = x[2];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (k < degree) {
            const float freq = (float)(1 << k);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sn, cs;
                sincosf(x[c] * freq, &sn, &cs);
                enc[3 + 6 * k + c] = sn;
                enc[6 + 6 * k + c] = cs;
            }
        }
    }
}

constexpr uint32_t kBitsSlotBytes = 4096;                 // sign bits of one 128 x 256 activation tile: [row][panel][2 words]
constexpr uint32_t kBitsTileBytes = 8 * kBitsSlotBytes;   // trunk layers 0..7

struct TcWorkspace {
    size_t view_bias, act, dy, bits, flags, merged_grad, total;
    size_t vis_pre, view_enc, vis_extra;   // SNERF_FLAG_VIS_HEAD only (vis_tc.cu)
    int n_tiles;
    uint32_t ring_cap;
};

// ---- gradient ring (backward): dY panels travel from the dgrad CTA pairs to the wgrad CTAs of the SAME launch ----
// One ring per dY stash slot (trunk layers 0..7, feature = 8, view layer = 9), `cap` entries each; tile t of a slot lives
// in entry t % cap.  Slot 9 holds 2 panels (32 KB) per tile, the others 4 (64 KB).
// SNERF_BWD_RING=0 (default): two launches, dgrad then wgrad, ring = all tiles (the gradients travel through HBM).
// SNERF_BWD_RING=n > 0: ONE launch with both roles co-resident and an n-entry ring that stays in L2 -- correct and
// tested, but measured slower (profiles/r2_fused_backward.md: the weight-gradient CTAs are bound by their own
// load -> MMA -> release latency per 64-point stage, not by HBM, so halving their number costs more than the saved traffic).
constexpr int kDySlots = 10;
constexpr int kRingConsumers = 2;     // at most two wgrad jobs read one dY slot (skip layer: hidden part and encoding part)
struct RingCtl {
    uint8_t* base;
    uint32_t* ready;       // [kDySlots][cap]: tile + 1 once the tile's panels of that slot are in the ring
    uint32_t* consumed;    // [kRingConsumers][kDySlots][cap]: tile + 1 once that consumer has copied the entry out
    uint32_t cap;
    int use_flags;         // 0: two-kernel form, no hand-off inside the launch
    uint8_t n_consumers[kDySlots];
};
__host__ __device__ __forceinline__ uint32_t ring_entry_bytes(int slot) { return slot == 9 ? 2u * kPanelBytes : 4u * kPanelBytes; }
__host__ __device__ __forceinline__ size_t ring_slot_off(int slot, uint32_t cap) { return (size_t)slot * cap * (4u * kPanelBytes); }

static inline int bwd_ring_tiles() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SNERF_BWD_RING");
        v = e ? atoi(e) : 0;
        if (v < 0) v = 0;
        if (v > 0 && v < 8) v = 8;     // >= 4 entries are needed for progress; keep a margin
    }
    return v;
}
static inline bool bwd_fused() { return bwd_ring_tiles() != 0; }
static inline uint32_t bwd_ring_cap(int n_tiles) {
    const int r = bwd_ring_tiles();
    return (uint32_t)((r == 0 || r >= n_tiles) ? (n_tiles > 0 ? n_tiles : 1) : r);
}
// dgrad CTA pairs of the fused launch (the remaining SMs run the wgrad jobs)
static inline int bwd_dgrad_pairs() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SNERF_BWD_DGRAD_PAIRS");
        v = e ? atoi(e) : 43;
        if (v < 1) v = 1;
    }
    return v;
}

// Sign bits of post-ReLU bf16 activations (the ReLU masks of the backward pass).  A 16-column unit is 8 packed pairs;
// pair p.lo / p.hi nonzero is bit 15 / 31 of (pair + 0x7FFF7FFF) -- the values are non-negative, so no carry crosses the
// halves.  Byte permutes gather the four flag bytes of two pairs and a shift files them at bit (8 * byte + kk) of the
// word, kk = 4 * (unit & 1) + pair / 2: 32 flags per word, two words per row and 64-column panel.
// generic-mode byte permute; bit 3 of a selector nibble replicates the msb of the selected byte (the __byte_perm intrinsic masks it off)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t relu_bits_unit(const uint32_t (&pk)[8], int unit) {
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t z = prmt(pk[2 * k] + 0x7FFF7FFFu, pk[2 * k + 1] + 0x7FFF7FFFu, 0x7531u);
        w |= (z & 0x80808080u) >> (7 - (4 * (unit & 1) + k));
    }
    return w;
}
// inverse: all-ones / zero half-word masks for the 8 pairs of a unit
__device__ __forceinline__ void relu_mask_unit(uint32_t word, int unit, uint32_t (&pk)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = word << (7 - (4 * (unit & 1) + k));
        pk[2 * k] &= prmt(t, 0u, 0x9988u);        // sign-replicate bytes 0 / 1
        pk[2 * k + 1] &= prmt(t, 0u, 0xBBAAu);    // sign-replicate bytes 2 / 3
    }
}

static inline TcWorkspace tc_ws_layout(const MlpDims& m, const TcPlan& pl, int n_rays, int n_samples, uint32_t flags) {
    TcWorkspace w{};
    const long long P = (long long)n_rays * n_samples;
    w.n_tiles = (int)((P + kTileRows - 1) / kTileRows);
    size_t off = 0;
    w.view_bias = off;
    off += align_up(m.has_view ? (size_t)n_rays * (128 + 4) * sizeof(float) : 0, 1024);     // per-ray view bias, then 4 compositing constants per ray
    if (flags & SNERF_FLAG_SAVE_FOR_BWD) {
        w.act = off;
        off += (size_t)w.n_tiles * pl.tile_stash_bytes;
        w.ring_cap = bwd_ring_cap(w.n_tiles);
        w.dy = off;
        off += (size_t)w.ring_cap * (m.has_view ? 9 * 65536 + 32768 : 8 * 65536);
        w.bits = off;
        off += (size_t)w.n_tiles * kBitsTileBytes;
        w.flags = off;
        off += align_up((size_t)(1 + kRingConsumers) * kDySlots * w.ring_cap * sizeof(uint32_t), 1024);
        w.merged_grad = off;          // fp32 [128 x view_in] + [128]: dY_v^T [h | encodings] and the column sums of dY_v
        off += align_up(m.has_view ? (size_t)(128 * m.view_in + 128) * sizeof(float) : 0, 1024);
    }
    if (flags & SNERF_FLAG_VIS_HEAD) {
        // visibility head on the tensor path: the view layer's point part per point (bf16 [P,128], written by the forward
        // kernel's view-step epilogue), PE(view_dir) per ray (fp32 [n_rays,32], written by tc_view_bias_kernel) and, for the
        // backward pass, the head's contribution to dY_v (fp32 [P,128], written by tc_vis_bwd_kernel, added by the dgrad prologue)
        w.vis_pre = off;
        off += align_up((size_t)P * 128 * 2, 1024);
        w.view_enc = off;
        off += align_up((size_t)n_rays * 32 * sizeof(float), 1024);
        if (flags & SNERF_FLAG_SAVE_FOR_BWD) {
            w.vis_extra = off;
            off += align_up((size_t)P * 128 * sizeof(float), 1024);
        }
    }
    w.total = off + 1024;
    return w;
}

static inline int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}


// persistent chain kernels: one CTA per SM, whole clusters only
static inline int chain_grid(int n_tiles) {
    int g = n_tiles < num_sms() ? n_tiles : num_sms();
    g = (g + kCluster - 1) / kCluster * kCluster;
    if (g > num_sms()) g -= kCluster;
    return g < kCluster ? kCluster : g;
}

// Every role walks the same job list: tiles 2g and 2g+1 of this pair occupy slots 0 and 1 and their steps interleave.
// jx = g * n_steps + s is the job's index within its slot (barrier phases count per slot).
#define SNERF_FOR_EACH_JOB(my_super, n_steps)                   \
    for (int g = 0; 2 * g < (my_super); ++g)                    \
        for (int s = 0; s < (n_steps); ++s)                     \
            for (int x = 0; x < 2; ++x)                         \
                if (2 * g + x < (my_super))

// pair kernels: one CTA pair per two SMs; a pair works on 256-point super tiles
static inline int pair_grid(int n_tiles) {
    const int n_super = (n_tiles + 1) / 2, pairs = num_sms() / 2;
    return 2 * (n_super < pairs ? n_super : pairs);
}

template <typename Kernel, typename Params>
static inline cudaError_t launch_clustered(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t st, const Params& p) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

}  // namespace snerf
