// Probes of the tcgen05 shared-memory-operand (SS) MMA path used while designing tensor-core WEIGHT gradients
// (dW = A^T G with the token index as the MMA's K dimension).  Findings on B200 (tests/wgrad_probe.py):
//   * kind::tf32 SS MMAs with K-major SWIZZLE_NONE operands and an arbitrary LBO work (sat_debug_ss 0/0, sat_debug_wgrad);
//   * MN-major tf32 operands (a_major / b_major = 1) return zeros with SWIZZLE_NONE (and with every other layout type) EXCEPT
//     SWIZZLE_128B_BASE32B, which works: see mn_probe.cu (round 2) and csrc/backward_tc_wg.cu;
//   * a full k_bwd_meta_tc with these weight-gradient GEMMs (commit "Tensor-core weight gradients in k_bwd_meta_tc") was
//     parity-green but ~7 % SLOWER than the CUDA-core wgrad32 version (scalar transposing stores, 48 MMAs per 32x32 block,
//     per-GEMM drain latency), so the shipped kernel keeps weight gradients on the CUDA cores.
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {

// ---- probe-only helpers: SS-form MMAs and the K-major transposed staging of the (not shipped) tensor-core weight gradient ----
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, M = 128 (both operands through shared-memory descriptors)
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor with both operands MN-major (weight-gradient GEMMs: the reduction dimension K is the token index)
__device__ __forceinline__ constexpr uint32_t make_idesc_mn(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- weight-gradient operand staging: [32 channels][128 tokens] fp32, K-major (K = token), SWIZZLE_NONE -----------------
// (MN-major tf32 operands return zeros on this part -- tests/wgrad_probe.py -- so the token-major register rows are
// transposed on the way into shared memory.)  Core matrix = 8 channels x 4 tokens (16 B per channel row, 128 B total).
// Token chunks (4 tokens) are the OUTER index and the 4 channel groups the inner one:
//     byte offset(ch, tok) = (tok/4)*528 + (ch/8)*128 + (ch%8)*16 + (tok%4)*4        LBO = 528 B, SBO = 128 B
// 528 = 4*128 + 16: the 16 B skew makes the per-token scalar stores of a warp hit 32 distinct banks.  An M = 128 MMA reads
// 16 channel groups (2 KB) per chunk: groups 4..15 alias the following chunks (finite don't-care rows of D).
constexpr int WG_LBO = 528, WG_SBO = 128;
constexpr int WG_TILE_FLOATS = (32 * WG_LBO + 2048) / 4;        // 32 chunks + overread room = 4736 floats
__device__ __forceinline__ void wg_stage_row(float* __restrict__ tile_hi, float* __restrict__ tile_lo, int tok, const float (&v)[D]) {
    const int base = (tok >> 2) * (WG_LBO / 4) + (tok & 3);      // floats
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float h = tf32_hi(v[c]);
        const int off = base + (c >> 3) * 32 + (c & 7) * 4;
        tile_hi[off] = h;
        tile_lo[off] = v[c] - h;
    }
}
// first MMA of a chain (overwrites D) / accumulating MMA: compile-time predicate, no setp
__device__ __forceinline__ void mma_ss_tf32_first(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ss_tf32_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
// D_w[128 x 32] = A^T G over the 128 tokens of the tile: 16 k-steps (8 tokens) x (hi*hi + lo*hi + hi*lo).
// Fully unrolled; the descriptor of k-step ks is the base descriptor + ks * (2*LBO/16) in the start-address field.
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* g_hi, const float* g_lo) {
    const uint32_t idesc = make_idesc(32);
    const uint64_t dah = make_b_desc(smem_u32(a_hi), WG_LBO, WG_SBO), dal = make_b_desc(smem_u32(a_lo), WG_LBO, WG_SBO);
    const uint64_t dgh = make_b_desc(smem_u32(g_hi), WG_LBO, WG_SBO), dgl = make_b_desc(smem_u32(g_lo), WG_LBO, WG_SBO);
    constexpr uint64_t STEP = (2 * WG_LBO) >> 4;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
        const uint64_t o = ks * STEP;
        if (ks == 0) mma_ss_tf32_first(d_tmem, dah, dgh, idesc);
        else mma_ss_tf32_acc(d_tmem, dah + o, dgh + o, idesc);
        mma_ss_tf32_acc(d_tmem, dal + o, dgh + o, idesc);
        mma_ss_tf32_acc(d_tmem, dah + o, dgl + o, idesc);
    }
}


constexpr int TCB_W = 192;
}

// ---- debugging aid: one weight-gradient GEMM  out[32][32] = A^T G  (A, G: [128][32]) through the staging + SS-MMA path ----
namespace sat {
__global__ void __launch_bounds__(TC_NT) k_debug_wgrad(const float* __restrict__ A, const float* __restrict__ G, float* __restrict__ out,
                                                       float* __restrict__ raw) {
    extern __shared__ __align__(128) float smem[];
    float* sAh = smem;
    float* sAl = sAh + WG_TILE_FLOATS;
    float* sGh = sAl + WG_TILE_FLOATS;
    float* sGl = sGh + WG_TILE_FLOATS;
    uint64_t* wbar = reinterpret_cast<uint64_t*>(sGl + WG_TILE_FLOATS + 1024);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(wbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * WG_TILE_FLOATS + 1024; i += TC_NT) sAh[i] = 0.f;
    if (warp == 0) tmem_alloc(s_tmem, TC_COLS);
    if (tid == 0) mbar_init(wbar, 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    float a[D], g[D];
    load_row(A + tid * D, a);
    load_row(G + tid * D, g);
    wg_stage_row(sAh, sAl, tid, a);
    wg_stage_row(sGh, sGl, tid, g);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        issue_wgrad(tbase + TCB_W, sAh, sAl, sGh, sGl);
        mma_commit(wbar);
    }
    mbar_wait(wbar, 0);
    tc_fence_after();
    float blk[D];
    tmem_ld32(tlane + TCB_W, blk);
    tc_wait_ld();
    store_row(raw + tid * D, blk);                 // all 128 lanes (rows >= 32 are don't-care)
    if (warp == 0) store_row(out + tid * D, blk);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}
}  // namespace sat

extern "C" int sat_debug_wgrad(const float* A, const float* G, float* out, float* raw, void* stream) {
    const size_t smem = (4 * sat::WG_TILE_FLOATS + 1024 + 16) * sizeof(float);
    if (cudaFuncSetAttribute(sat::k_debug_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_debug_wgrad: smem");
    sat::k_debug_wgrad<<<1, sat::TC_NT, smem, (cudaStream_t)stream>>>(A, G, out, raw);
    return sat_host::check_launch("k_debug_wgrad");
}

// ---- debugging aid 2: D[128][32] = P[128][128] Q[32][128]^T (K = 128) via SS MMAs with K-major or MN-major staging -----------
namespace sat {
__global__ void __launch_bounds__(TC_NT) k_debug_ss(const float* __restrict__ P, const float* __restrict__ Q, float* __restrict__ raw,
                                                    int a_mn, int b_mn, int swap_lbo_sbo) {
    extern __shared__ __align__(128) float smem[];
    float* sP = smem;                 // 128 x 128
    float* sQ = sP + 128 * 128;       // 32 x 128
    uint64_t* wbar = reinterpret_cast<uint64_t*>(sQ + 32 * 128 + 64);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(wbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // stage: K-major: off(r,k) = (r>>3)*1024 + (k>>2)*32 + (r&7)*4 + (k&3);  MN-major: off(r,k) = (k>>3)*(rows*8) + (r>>2)*32 + (k&7)*4 + (r&3)
    for (int e = tid; e < 128 * 128; e += TC_NT) {
        const int r = e >> 7, k = e & 127;
        const int off = a_mn ? ((k >> 3) * (128 * 8) + (r >> 2) * 32 + (k & 7) * 4 + (r & 3)) : ((r >> 3) * 1024 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3));
        sP[off] = P[e];
    }
    for (int e = tid; e < 32 * 128; e += TC_NT) {
        const int r = e >> 7, k = e & 127;
        const int off = b_mn ? ((k >> 3) * (32 * 8) + (r >> 2) * 32 + (k & 7) * 4 + (r & 3)) : ((r >> 3) * 1024 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3));
        sQ[off] = Q[e];
    }
    if (warp == 0) tmem_alloc(s_tmem, TC_COLS);
    if (tid == 0) mbar_init(wbar, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) |
                               ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t pa = smem_u32(sP), qa = smem_u32(sQ);
        for (int ks = 0; ks < 16; ++ks) {
            // K-major: k-step = 2 core matrices along K = 256 B; LBO = 128 (K dir), SBO = 4096 (8-row groups)
            // MN-major: k-step = one 8-token group: A: 128*8*4 = 4096 B, B: 32*8*4 = 1024 B; SBO = 128 (MN groups), LBO = group stride
            uint32_t a_lbo = a_mn ? 4096u : 128u, a_sbo = a_mn ? 128u : 4096u;
            uint32_t b_lbo = b_mn ? 1024u : 128u, b_sbo = b_mn ? 128u : 4096u;
            if (swap_lbo_sbo) {
                if (a_mn) { const uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; }
                if (b_mn) { const uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            }
            const uint64_t da = make_b_desc(pa + ks * (a_mn ? 4096u : 256u), a_lbo, a_sbo);
            const uint64_t db = make_b_desc(qa + ks * (b_mn ? 1024u : 256u), b_lbo, b_sbo);
            mma_ss_tf32(tbase + TCB_W, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        mma_commit(wbar);
    }
    mbar_wait(wbar, 0);
    tc_fence_after();
    float blk[D];
    tmem_ld32(tlane + TCB_W, blk);
    tc_wait_ld();
    store_row(raw + tid * D, blk);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}
}  // namespace sat

extern "C" int sat_debug_ss(const float* P, const float* Q, float* raw, int32_t a_mn, int32_t b_mn, int32_t swap, void* stream) {
    const size_t smem = (128 * 128 + 32 * 128 + 64 + 16) * sizeof(float);
    if (cudaFuncSetAttribute(sat::k_debug_ss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_debug_ss: smem");
    sat::k_debug_ss<<<1, sat::TC_NT, smem, (cudaStream_t)stream>>>(P, Q, raw, a_mn, b_mn, swap);
    return sat_host::check_launch("k_debug_ss");
}
