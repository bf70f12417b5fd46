// Counter-based random numbers for the kernels that consume them (stratified sampling :299, sample_pdf :341, sigma noise :670):
// Philox4x32-10 keyed by a 64-bit seed, counter = (index of a block of four elements, 64-bit draw offset).  Element e of draw
// `offset` is component e & 3 of block e >> 2, whatever kernel asks for it: snerf_fill_random writes exactly the numbers the
// *_rng entry points consume in place, which is how the tests pin them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace snerf {

struct RngKey {
    uint64_t seed, offset;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

__device__ __forceinline__ uint4 rng_block(const RngKey key, unsigned long long block) {
    return philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)key.offset, (uint32_t)(key.offset >> 32)),
                         make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32)));
}

// 24-bit uniform in [0, 1) (the resolution of torch.rand in fp32)
__device__ __forceinline__ float rng_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float4 rng_uniform4(const RngKey key, unsigned long long block) {
    const uint4 r = rng_block(key, block);
    return make_float4(rng_u01(r.x), rng_u01(r.y), rng_u01(r.z), rng_u01(r.w));
}

// four standard normals: two Box-Muller pairs; the radius takes a uniform in (0, 1]
__device__ __forceinline__ float4 rng_normal4(const RngKey key, unsigned long long block) {
    const uint4 r = rng_block(key, block);
    const float u0 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u1 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u1));
    float s0, c0, s1, c1;
    sincospif(2.0f * rng_u01(r.y), &s0, &c0);
    sincospif(2.0f * rng_u01(r.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__device__ __forceinline__ float rng_pick(const float4 v, unsigned long long element) {
    const int c = (int)(element & 3ull);
    return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
}

}  // namespace snerf
