// Debug probe (not part of the public ABI): one cta_group::2 GEMM  D[256 x N] = A[256 x 64] B[N x 64]^T  on a CTA pair.
// Pins the operand split the chain kernels rely on: CTA r holds A rows [128 r, 128 r + 128) and B rows
// [N/2 r, N/2 r + N/2) at the same smem offsets and receives D rows [128 r, 128 r + 128) in its own TMEM.
// mode 1 additionally completes the peer's weight copy directly on the leader's mbarrier (remote complete_tx).
#include "common.cuh"
#include "tc_common.cuh"

namespace snerf {
using namespace tc;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_probe_kernel(const uint8_t* __restrict__ a_img,
                                                                                  const uint8_t* __restrict__ b_img,
                                                                                  float* __restrict__ d_out, int n, int mode,
                                                                                  long long* timing) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_load, bar_peer, bar_done, bar_aux;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + 16384;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = cluster_rank();
    const uint32_t b_half = (uint32_t)(n / 2) * kRowBytes;
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_peer, 1);
        mbar_init(&bar_done, 1);
        mbar_init(&bar_aux, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc2<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        if (mode == 1 && rank == 1) {
            // peer: expect the bytes on the LEADER's barrier and let the copy complete there
            const uint32_t remote = cluster_addr(&bar_peer, 0);
            asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(16384u + b_half) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sa)),
                         "l"(a_img + 16384), "r"(16384u), "r"(remote) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sb)),
                         "l"(b_img + b_half), "r"(b_half), "r"(remote) : "memory");
        } else {
            mbar_arrive_expect_tx(&bar_load, 16384u + b_half);
            bulk_g2s(sa, a_img + rank * 16384, 16384u, &bar_load);
            bulk_g2s(sb, b_img + rank * b_half, b_half, &bar_load);
            mbar_wait(&bar_load, 0);
            if (rank == 1) mbar_arrive_cluster(cluster_addr(&bar_peer, 0));   // relay: my half has landed
        }
        if (rank == 0) {
            mbar_wait_cl(&bar_peer, 0);
            tc_fence_after();
            const uint32_t idesc = umma_idesc(256, n, false, false);
            const long long t0 = clock64();
            // mode bits (timing runs): 2 = commit to a spare barrier per 4 MMAs, 4 = cluster-scope wait on a completed
            // barrier per 4 MMAs, 8 = cta-scope wait per 4 MMAs, 16 = non-blocking cluster-scope probe per 4 MMAs
            for (int rep = 0; rep < (timing ? 64 : 1); ++rep) {
                if (mode & 4) { mbar_wait_cl(&bar_peer, 0); tc_fence_after(); }
                if (mode & 8) { mbar_wait(&bar_peer, 0); tc_fence_after(); }
                if (mode & 16) { (void)mbar_test_wait_cl(&bar_peer, 0); tc_fence_after(); }
                for (int k = 0; k < 4; ++k)
                    umma2(tmem, umma_desc_kmajor(smem_u32(sa), k), umma_desc_kmajor(smem_u32(sb), k), idesc, k != 0);
                if (mode & 2) umma_commit2(&bar_aux, 3);
            }
            const long long t1 = clock64();
            umma_commit2(&bar_done, 3);
            mbar_wait(&bar_done, 0);
            if (timing) { timing[0] = t1 - t0; timing[1] = clock64() - t0; }
        }
    }
    __syncwarp();
    mbar_wait(&bar_done, 0);
    tc_fence_after();
    for (int c = 0; c < n; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        float* row = d_out + (size_t)(rank * 128 + warp * 32 + lane) * n + c;
#pragma unroll
        for (int i = 0; i < 32; ++i) row[i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2<512>(tmem);
}

}  // namespace snerf

extern "C" int snerfdbg_pair_probe(const void* a_img, const void* b_img, float* d_out, int n, int mode, long long* timing, void* stream) {
    using namespace snerf;
    SNERF_REQUIRE(n == 256 || n == 128, "pair probe: N must be 128 or 256");
    SNERF_CUDA_OK(cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
    pair_probe_kernel<<<2, 128, 49152 + 1024, (cudaStream_t)stream>>>((const uint8_t*)a_img, (const uint8_t*)b_img, d_out, n, mode, timing);
    SNERF_LAUNCH_OK("pair_probe_kernel");
    return SNERF_OK;
}
