// tcgen05 / TMEM / mbarrier PTX wrappers and the packed-weight (UMMA B operand) layout shared by the tensor-core kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace sat {

constexpr int TC_NT = 128;            // threads per CTA = TMEM lanes = tokens per tile
constexpr int TC_COLS = 256;          // TMEM columns allocated per CTA (2 CTAs / SM)
constexpr int TC_A = 0;               // [0,64)   A operand, K = 32: hi at +0, lo at +32
constexpr int TC_H = 64;              // [64,128) A operand for one half of the MetaNet hidden layer
constexpr int TC_D = 128;             // [128,224) accumulators: QKV 96 | hidden 64 + out 32 | out-proj 32
constexpr int PACK_FLOATS = 8192;     // floats per packed layer block / per packed key block
// packed layer block: [QKV hi 3072][QKV lo 3072][WO hi 1024][WO lo 1024]
constexpr int PK_QKV_HI = 0, PK_QKV_LO = 3072, PK_WO_HI = 6144, PK_WO_LO = 7168;
// packed key block:   [W1 hi 2048][W1 lo 2048][W2a hi 1024][W2a lo 1024][W2b hi 1024][W2b lo 1024]
constexpr int PK_W1_HI = 0, PK_W1_LO = 2048, PK_W2A_HI = 4096, PK_W2A_LO = 5120, PK_W2B_HI = 6144, PK_W2B_LO = 7168;
// wpack = [L forward-layer blocks][n_keys forward-key blocks][L backward-layer blocks][n_keys backward-key blocks]
// backward layer block: B[n=i][k=j] = W[i][j] for W_Q, W_K, W_V (dX += dP @ W^T) and B[n=i][k=j] = W_out[j][i] (dR = dO @ W_out)
constexpr int PKB_WQN_HI = 0, PKB_WQN_LO = 1024, PKB_WKN_HI = 2048, PKB_WKN_LO = 3072, PKB_WVN_HI = 4096, PKB_WVN_LO = 5120,
              PKB_WOT_HI = 6144, PKB_WOT_LO = 7168;
// backward key block: W2T: B[n=c][k=j] = W2[c][j] (N = 64);  W1N half h: B[n=i][k=c] = W1[i][h*32 + c] (N = 32)
constexpr int PKB_W2T_HI = 0, PKB_W2T_LO = 2048, PKB_W1N0_HI = 4096, PKB_W1N0_LO = 5120, PKB_W1N1_HI = 6144, PKB_W1N1_LO = 7168;

// offset (floats) of B[n][k] in the canonical K-major, no-swizzle UMMA layout for K = 32 fp32/tf32 elements:
// core matrix = 8 rows (n) x 16 B (4 k); core matrices contiguous along k (LBO = 128 B), 8-row groups SBO = 1024 B apart
__host__ __device__ __forceinline__ int b_off32(int n, int k) { return (n >> 3) * 256 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3); }

// the same for K = 32 bf16 elements (kind::f16): core matrix = 8 rows x 16 B (8 k); LBO = 128 B, SBO = 512 B; offset in ELEMENTS
__host__ __device__ __forceinline__ int b_off16(int n, int k) { return (n >> 3) * 256 + (k >> 3) * 64 + (n & 7) * 8 + (k & 7); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers (tcgen05 / TMEM / mbarrier)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Spin on the barrier's phase.  A wait that has not completed after ~4 s of SM clock (a lost TMA transaction / MMA commit: a bug)
// traps, so that the launch fails loudly instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok, spins = 0;
    long long t0 = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && (++spins & 0x3FFu) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) __trap();
        }
    } while (!ok);
}
// ---- TMA: bulk asynchronous copies global -> shared, completion counted in bytes on an mbarrier -------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one box of the 2-D tensor map (cp.async.bulk.tensor -> SASS UTMALDG): x = element column, y = row of the box's first element
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// contiguous bytes (multiple of 16, both sides 16 B aligned) without a tensor map (cp.async.bulk -> SASS UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// the packed weights as a 2-D tensor [blocks * 32 rows][256 floats]; one box = TMA_BOX_ROWS rows = 8 KB
constexpr int TMA_ROW_FLOATS = 256, TMA_BOX_ROWS = 8, TMA_BOX_FLOATS = TMA_ROW_FLOATS * TMA_BOX_ROWS;
constexpr int PACK_ROWS = PACK_FLOATS / TMA_ROW_FLOATS;      // 32 rows per packed block

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, M = 128
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], kind::f16 (bf16 operands, fp32 accumulate), M = 128
__device__ __forceinline__ void mma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
                   "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
                   "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                   "r"(__float_as_uint(v[15])) : "memory");
}
// 32 floats of this thread's lane: columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[D]) {
    tmem_ld16(taddr, &v[0]);
    tmem_ld16(taddr + 16, &v[16]);
}
// PASSES selects the arithmetic of the projections: 1 = TF32, 3 = error-compensated 3xTF32 (default, fp32-level parity),
// 2 = bf16 operands with fp32 accumulation (kind::f16; reported separately, never the headline).
// write a 32-vector as A operand: hi at col, lo at col + 32 (PASSES == 3), as is (PASSES == 1), or as 16 columns of packed bf16
// pairs (PASSES == 2: element 2c in the low half of column c)
template <int PASSES>
__device__ __forceinline__ void tmem_st_operand(uint32_t taddr, const float (&v)[D]) {
    if (PASSES == 2) {
        float pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            uint32_t w;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(v[2 * i + 1]), "f"(v[2 * i]));      // {hi half, lo half}
            pk[i] = __uint_as_float(w);
        }
        tmem_st16(taddr, &pk[0]);
    } else if (PASSES == 1) {
        tmem_st16(taddr, &v[0]);
        tmem_st16(taddr + 16, &v[16]);
    } else {
        float hi[D], lo[D];
#pragma unroll
        for (int i = 0; i < D; i += 2) {
            hi[i] = tf32_hi(v[i]); hi[i + 1] = tf32_hi(v[i + 1]);
            add2(lo[i], lo[i + 1], v[i], v[i + 1], -hi[i], -hi[i + 1]);      // lo = v - hi (exact)
        }
        tmem_st16(taddr, &hi[0]);
        tmem_st16(taddr + 16, &hi[16]);
        tmem_st16(taddr + 32, &lo[0]);
        tmem_st16(taddr + 48, &lo[16]);
    }
}

// issue D[128 x N] (+)= A[128 x 32] @ B[N x 32]^T as 4 k-steps of K = 8 (x3 products with the hi/lo split)
template <int PASSES>
__device__ __forceinline__ void issue_gemm_k32(uint32_t d_tmem, uint32_t a_tmem, const float* b_hi, const float* b_lo, int N,
                                               bool accumulate) {
    const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
    uint32_t acc = accumulate ? 1u : 0u;
    if (PASSES == 2) {
        // bf16: the operand is stored as independent groups of 32 rows, group g at float offset 1024 g of the "hi" region (so
        // that the 32-row slices the backward fetches sit where the fp32 layout has them); per group two k-steps of K = 16
        const uint32_t idesc16 = make_idesc_bf16(32);
        for (int g = 0; g < (N >> 5); ++g) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
                mma_ts_bf16(d_tmem + 32 * g, a_tmem + kk * 8, make_b_desc(bh + g * 4096 + kk * 256, 128, 512), idesc16,
                            (kk > 0 || accumulate) ? 1u : 0u);
        }
        return;
    }
    const uint32_t idesc = make_idesc(N);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const uint64_t dh = make_b_desc(bh + kk * 256, 128, 1024);
        mma_ts_tf32(d_tmem, a_tmem + kk * 8, dh, idesc, acc);
        acc = 1u;
        if (PASSES == 3) {
            const uint64_t dl = make_b_desc(bl + kk * 256, 128, 1024);
            mma_ts_tf32(d_tmem, a_tmem + 32 + kk * 8, dh, idesc, 1u);
            mma_ts_tf32(d_tmem, a_tmem + kk * 8, dl, idesc, 1u);
        }
    }
}


}  // namespace sat
