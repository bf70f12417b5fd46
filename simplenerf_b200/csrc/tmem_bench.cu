// Debug microbenchmark (not part of the public ABI): TMEM -> register read throughput of tcgen05.ld for several
// instruction shapes and warp counts, optionally while the tensor pipe runs back-to-back MMAs into other columns.
// tools/tmem_bench.py prints the table; the numbers size the epilogue of the chain kernels.
#include "common.cuh"
#include "tc_common.cuh"

namespace snerf {
using namespace tc;

template <int X>
__device__ __forceinline__ void ld_shape(uint32_t taddr, uint32_t& sink) {
    if constexpr (X == 32) {
        uint32_t r[32];
        tmem_ld32_issue(taddr, r);
        tmem_ld_wait(r);
        sink ^= r[0] ^ r[31];
    } else if constexpr (X == 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink ^= r[0] ^ r[15];
    } else {   // 16x256b.x8: 16 lanes x 256 bit, 8 repeats = 32 registers per thread (64 columns of 16 lanes... per half warp)
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink ^= r[0] ^ r[31];
    }
}

// blockDim = 32 * (1 + n_ld_warps): warp 0 = MMA issuer, warps 1.. = readers.  out[warp] = cycles for `reps` loads.
__global__ void __launch_bounds__(544) tmem_bench_kernel(long long* out, int reps, int shape, int with_mma, int pipelined) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int stop;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_fence_init(); stop = 0; }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp == 0) {
        if (lane == 0 && with_mma) {
            const uint32_t idesc = umma_idesc(128, 256, false, false);
            const uint32_t a = smem_u32(smem), b = smem_u32(smem + 65536);
            int n = 0;
            const long long tm0 = clock64();
            while (*(volatile int*)&stop == 0 && n < 200000) {
                for (int k = 0; k < 16; ++k) umma(tmem + 256, umma_desc_kmajor(a + (k >> 2) * 16384, k & 3), umma_desc_kmajor(b, k & 3), idesc, true);
                n += 16;
            }
            umma_commit(&bar_done);
            mbar_wait(&bar_done, 0);
            out[63] = n;
            out[61] = clock64() - tm0;
        }
    } else {
        const int q = warp & 3;
        const uint32_t base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t sink = 0;
        // warm-up
        ld_shape<32>(base, sink);
        asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 32) : "memory");
        const long long t0 = clock64();
        if (shape == 32) {
            if (pipelined) {
                uint32_t ra[32], rb[32];
                tmem_ld32_issue(base, ra);
                for (int i = 0; i < reps; i += 2) {
                    tmem_ld_wait(ra);
                    tmem_ld32_issue(base + ((i + 1) * 32 & 255), rb);
                    sink ^= ra[0] ^ ra[31];
                    tmem_ld_wait(rb);
                    tmem_ld32_issue(base + ((i + 2) * 32 & 255), ra);
                    sink ^= rb[0] ^ rb[31];
                }
                tmem_ld_wait(ra);
                sink ^= ra[0];
            } else {
                for (int i = 0; i < reps; ++i) ld_shape<32>(base + (i * 32 & 255), sink);
            }
        } else if (shape == 16) {
            for (int i = 0; i < reps; ++i) ld_shape<16>(base + (i * 16 & 255), sink);
        } else {
            for (int i = 0; i < reps; ++i) ld_shape<256>(base + (i * 64 & 255), sink);
        }
        const long long t1 = clock64();
        if (lane == 0) out[warp] = t1 - t0;
        if (sink == 0x12345u) out[62] = sink;
        asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 32) : "memory");
        if (threadIdx.x == 32) *(volatile int*)&stop = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace snerf

extern "C" int snerfdbg_tmem_bench(long long* out, int n_ld_warps, int reps, int shape, int with_mma, int pipelined, void* stream) {
    using namespace snerf;
    SNERF_CUDA_OK(cudaFuncSetAttribute(tmem_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024));
    tmem_bench_kernel<<<1, 32 * (1 + n_ld_warps), 98304 + 1024, (cudaStream_t)stream>>>(out, reps, shape, with_mma, pipelined);
    SNERF_LAUNCH_OK("tmem_bench_kernel");
    return SNERF_OK;
}
