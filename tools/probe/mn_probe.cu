// Probe (not part of libsatrans_b200.so): tcgen05.mma.kind::tf32 with BOTH operands MN-major in shared memory -- the form a
// WEIGHT gradient needs (dW = A^T G reduces over tokens, so the token index is the MMA's K and the natural [token][channel]
// rows are MN-major).  D[128 x 32] = A^T G with A: [128 tokens][128 channels], G: [128 tokens][32].
// Staging under test: per 32-channel block one atom [token][32 floats] with 128-byte rows, 16-byte chunk c of token t stored
// at chunk c ^ (t & 7) (SWIZZLE_128B), SBO = 1024 B (8 tokens), LBO = distance between the 32-channel blocks.
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {
__device__ __forceinline__ uint64_t make_desc_x(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)(layout & 7u) << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// flags: bit 0 apply the 16-byte XOR swizzle when staging, bit 1 swap LBO/SBO, bit 2 BASE32B staging (SBO = 512), bits 4..6 layout type
__global__ void __launch_bounds__(TC_NT) k_mn_probe(const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ out,
                                                    int flags) {
    extern __shared__ __align__(1024) float smem[];
    float* sA = smem;                       // 4 atoms x [128][32]
    float* sG = sA + 4 * 128 * 32;          // 1 atom
    uint64_t* bar = reinterpret_cast<uint64_t*>(sG + 128 * 32);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool swz = flags & 1;
    const uint32_t layout = (flags >> 4) & 7;
    {
        const int t = tid;
#pragma unroll 1
        for (int blk = 0; blk < 5; ++blk) {
            const float* src = blk < 4 ? A + t * 128 + blk * 32 : G + t * 32;
            float* dst = (blk < 4 ? sA + blk * 128 * 32 : sG) + t * 32;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(src + c * 4);
                int cc = swz ? (c ^ (t & 7)) : c;
                if (flags & 4) cc = ((((c >> 1) ^ (t & 3)) << 1) | (c & 1));       // SWIZZLE_128B_BASE32B: 32-byte chunks ^ (row & 3)
                *reinterpret_cast<float4*>(dst + cc * 4) = v;
            }
        }
    }
    if (warp == 0) tmem_alloc(s_tmem, 64);
    if (tid == 0) mbar_init(bar, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t lbo = 128 * 32 * 4, sbo = (flags & 4) ? 512 : 1024;
        if (flags & 2) { const uint32_t x = lbo; lbo = sbo; sbo = x; }
        for (int ks = 0; ks < 16; ++ks) {
            const uint64_t da = make_desc_x(smem_u32(sA) + ks * 1024, lbo, sbo, layout);
            const uint64_t db = make_desc_x(smem_u32(sG) + ks * 1024, lbo, sbo, layout);
            mma_ss(tbase, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    float row[D];
    tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16), row);
    tc_wait_ld();
    store_row(out + tid * D, row);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}
}  // namespace sat

extern "C" int sat_mn_probe(const float* A, const float* G, float* out, int32_t flags, void* stream) {
    const size_t smem = (5 * 128 * 32 + 16) * sizeof(float) + 1024;
    if (cudaFuncSetAttribute(sat::k_mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_mn_probe: smem");
    sat::k_mn_probe<<<1, sat::TC_NT, smem, (cudaStream_t)stream>>>(A, G, out, flags);
    return sat_host::check_launch("k_mn_probe");
}
