// Developer probe (libsimplenerf_b200_dbg.so only): how fast can ONE SM move a 64 KB shared-memory panel set to global memory?
//   mode 0: one thread, cp.async.bulk shared -> global (what the stash writers of the chain kernels do), `depth` groups in flight
//   mode 1: every thread, 16-byte ld.shared + st.global (coalesced 512 B per warp instruction)
//   mode 2: half the bytes each way, concurrently
//   mode 3: bulk copies of 16 KB (four per set) instead of one 64 KB copy
// Each CTA cycles through `window` bytes of its own global region (small window: stays in L2; large: goes to HBM).
#include "common.cuh"
#include "tc_common.cuh"

namespace snerf {
using namespace tc;

__global__ void __launch_bounds__(512, 1) store_probe_kernel(uint8_t* __restrict__ dst, size_t window, int reps, int mode, int depth,
                                                             long long* __restrict__ cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, blockIdx.x, 3u, 4u);
    fence_async_smem();
    __syncthreads();
    uint8_t* base = dst + (size_t)blockIdx.x * window;
    const size_t sets = window / 65536;
    const long long t0 = clock64();
    const bool bulk_thread = threadIdx.x == 0 && mode != 1;
    const int first_direct = mode == 2 ? 32 : 0;       // mode 2: warp 0 issues the bulk half, the other warps store the second half
    for (int r = 0; r < reps; ++r) {
        uint8_t* out = base + (size_t)(r % sets) * 65536;
        if (bulk_thread) {
            if (mode == 0) bulk_s2g(out, smem, 65536);
            else if (mode == 2) bulk_s2g(out, smem, 32768);
            else for (int j = 0; j < 4; ++j) bulk_s2g(out + j * 16384, smem + j * 16384, 16384);
            bulk_commit();
            if (depth <= 1) bulk_wait_read<0>(); else if (depth == 2) bulk_wait_read<1>(); else bulk_wait_read<3>();
        }
        if (mode == 1 || (mode == 2 && (int)threadIdx.x >= first_direct)) {
            const int lo = mode == 2 ? 32768 / 16 : 0, n = 65536 / 16;
            for (int i = lo + (int)threadIdx.x - first_direct; i < n; i += (int)blockDim.x - first_direct)
                reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(smem)[i];
        }
        if (mode == 2) __syncthreads();
    }
    if (bulk_thread) bulk_wait_all<0>();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

}  // namespace snerf

extern "C" int snerfdbg_store_probe(void* dst, size_t window, int reps, int mode, int depth, int grid, long long* cycles, void* stream) {
    using namespace snerf;
    SNERF_REQUIRE(window >= 65536 && window % 65536 == 0 && reps > 0 && grid > 0, "store probe: bad sizes");
    SNERF_CUDA_OK(cudaFuncSetAttribute(store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024));
    store_probe_kernel<<<grid, 512, 65536 + 1024, (cudaStream_t)stream>>>((uint8_t*)dst, window, reps, mode, depth, cycles);
    SNERF_LAUNCH_OK("store_probe_kernel");
    return SNERF_OK;
}
