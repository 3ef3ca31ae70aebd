// Variant of backward_tc.cu with the WEIGHT gradients on the tensor core as well (selected with SAT_WGRAD_TC=1; not the default).
// Parity-green (tests/test_gpu_parity.py::test_tensor_core_weight_gradients) and measured on B200: 1.42 ms of backward per step
// against 1.37 ms for the CUDA-core weight gradients of backward_tc.cu.  A 32x32 (x 2..4) product over the 120 tokens of a tile
// needs 15 k-steps x 3 TF32 terms = 45 MMA instructions whatever its size, six products per tile add ~270 issues to the ~150
// of the data-gradient rounds, and with the FFMA work gone the two resident CTAs per SM no longer fill each other's round-trip
// waits (TMEM: 2 x 256 columns, shared memory: 2 x 110 KB -- a third CTA does not fit).  Kept as the worked answer to "can
// tf32 weight gradients run MN-major on tcgen05": yes, with SWIZZLE_128B_BASE32B only.
//
// Tensor-core version of k_bwd_meta (MetaNet-LayerNorm bwd + MetaNet bwd + QKV projection bwd), sm_100a.
//
// Same thread-per-token / 128-lane mapping as forward_tc.cu.  Every DATA-gradient / recompute projection is a
// tcgen05.mma.kind::tf32 grouped GEMM with the tokens' vectors as A operand in TMEM and pre-packed weights (the
// backward blocks of k_pack_weights) as B operand in shared memory:
//     u = x W                (recompute)        pre = u W1            (recompute, ReLU decisions come from the forward)
//     dh = dm W2^T                              du += dh W1^T                              dX += [du_q|du_k|dv] [W_Q|W_K|W_V]^T
// WEIGHT gradients (reductions over tokens) run on the tensor core too: both operands are the natural [token][channel] rows,
// i.e. MN-major shared-memory operands (see "weight gradients on the tensor core" below).
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {

constexpr int TCB_A = 0, TCB_H = 64, TCB_D = 128, TCB_H2 = 192;          // TMEM column map (256 columns)
constexpr int TCB_SB = 4096;                                             // floats of the B-operand staging buffer (16 KB)

// ---- weight gradients on the tensor core --------------------------------------------------------------------------------------
// dW = A^T G reduces over TOKENS, so the token index is the MMA's K and the natural [token][32 channels] rows are MN-major
// operands.  tcgen05.mma.kind::tf32 takes MN-major shared-memory operands in ONE layout only (probe: tools/probe/mn_probe.cu;
// every other layout type silently yields zeros): SWIZZLE_128B_BASE32B -- rows of 128 bytes, the 32-byte chunk q of row t
// stored at chunk q ^ (t & 3), 4 rows per swizzle atom (SBO = 512 B), one MMA (K = 8 tokens) per 1024 bytes, the 32-channel
// blocks of an operand LBO bytes apart.  A thread stores ITS token's row with eight 16-byte stores per tile (hi and lo tiles of
// the 3xTF32 split), no transposition.  Every product is issued as M = 128 (four 32-channel blocks; the blocks past the real
// ones are whatever finite values the neighbouring tiles hold: their rows of D are never read) and rides in the SAME commit
// as the next data-gradient round, so it costs no extra round trip.  The result (D^T rows = output channels, one per lane) is
// added to this CTA's row of the per-CTA partial buffers in global memory (L2-resident, one owner per element: deterministic).
constexpr int WG_LAYOUT_128B_BASE32B = 1;
__host__ __device__ __forceinline__ int wg_rows(int F) { return ((TC_NT / F) * F + 7) & ~7; }       // tile rows (K = tokens), multiple of 8
__device__ __forceinline__ uint64_t make_wg_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((512u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)WG_LAYOUT_128B_BASE32B << 61);
}
// instruction descriptor: D = F32, A = B = TF32, both MN-major, M = 128, N = 32
__device__ __forceinline__ constexpr uint32_t make_idesc_wg() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// this thread's row (token t) of an operand -> its hi / lo tiles (t < tile rows; rows past the last token carry zeros)
__device__ __forceinline__ void wg_stage(float* __restrict__ tile_hi, float* __restrict__ tile_lo, int t, const float (&v)[D]) {
    float* rh = tile_hi + t * D;
    float* rl = tile_lo + t * D;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int pos = (q ^ (t & 3)) * 8;
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            hi[i] = tf32_hi(v[q * 8 + i]); hi[i + 1] = tf32_hi(v[q * 8 + i + 1]);
            add2(lo[i], lo[i + 1], v[q * 8 + i], v[q * 8 + i + 1], -hi[i], -hi[i + 1]);
        }
        *reinterpret_cast<float4*>(rh + pos) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(rh + pos + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<float4*>(rl + pos) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(rl + pos + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    }
}
// D^T[128 x 32] = [A blocks]^T G over the tile's tokens: ksteps x (hi*hi + lo*hi + hi*lo); a_* = first 32-channel block of A
// (the following blocks LBO = tile bytes apart), g_* = the 32-channel G tile
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* g_hi, const float* g_lo,
                                            int ksteps, uint32_t tile_bytes) {
    const uint32_t idesc = make_idesc_wg();
    const uint64_t dah = make_wg_desc(smem_u32(a_hi), tile_bytes), dal = make_wg_desc(smem_u32(a_lo), tile_bytes);
    const uint64_t dgh = make_wg_desc(smem_u32(g_hi), tile_bytes), dgl = make_wg_desc(smem_u32(g_lo), tile_bytes);
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t o = (uint64_t)ks * (1024u >> 4);           // 8 tokens = 1024 bytes further in the start-address field
        mma_ss_tf32(d_tmem, dah + o, dgh + o, idesc, ks > 0 ? 1u : 0u);
        mma_ss_tf32(d_tmem, dal + o, dgh + o, idesc, 1u);
        mma_ss_tf32(d_tmem, dah + o, dgl + o, idesc, 1u);
    }
}

// Shared memory: B staging (16 KB) | six operand tiles [rows][32] (A blocks hi0 hi1 lo0 lo1, G hi lo) | LN params | dLN slots |
// barriers.  F = 19: 120 rows -> 15 KB per tile, 110 KB in total -> 2 CTAs per SM.  (+1 KB: the tiles start 1024-byte aligned.)
static size_t tcb_smem_bytes(int F) {
    return (size_t)(TCB_SB + 6 * wg_rows(F) * D + 192 + (TC_NT / 32) * 128 + 8) * sizeof(float) + 1024;
}

// B operands of the next round: n floats of packed weights -> sB by TMA (cp.async.bulk, SASS UBLKCP), issued by the MMA-issuing
// thread; it waits for the bytes (mbarrier transaction count) right before issuing the round's MMAs, so the copy rides under
// the CUDA-core work in between.  Callers make sure every MMA / thread that read the destination has finished (a round has
// been waited for, or a __syncthreads); the proxy fence orders earlier generic-proxy accesses of sB before the async write.
__device__ __forceinline__ void load_b(float* __restrict__ sB, const float* __restrict__ src, int n, int tid, uint64_t* bar) {
    if (tid == 0) {
        fence_async_smem();
        mbar_expect_tx(bar, (uint32_t)n * 4u);
        tma_load_1d(sB, src, (uint32_t)n * 4u, bar);
    }
}
// two pieces, one transaction
__device__ __forceinline__ void load_b2(float* __restrict__ d0, const float* __restrict__ s0, float* __restrict__ d1,
                                        const float* __restrict__ s1, int n, int tid, uint64_t* bar) {
    if (tid == 0) {
        fence_async_smem();
        mbar_expect_tx(bar, 2u * (uint32_t)n * 4u);
        tma_load_1d(d0, s0, (uint32_t)n * 4u, bar);
        tma_load_1d(d1, s1, (uint32_t)n * 4u, bar);
    }
}

// this lane's row of a weight-gradient result (TMEM columns [col, col + 32)) += into the CTA's partial row: elements
// dst[j * ld] for j < 32 (ld = 1: a contiguous row; otherwise a column of a row-major matrix, coalesced across the lanes)
__device__ __forceinline__ void wg_flush(uint32_t taddr, float* __restrict__ dst, int ld, bool active) {
    float w[D];
    tmem_ld32(taddr, w);
    tc_wait_ld();
    if (!active) return;
    if (ld == 1) {
#pragma unroll
        for (int j = 0; j < D; j += 4) {
            float4 c = *reinterpret_cast<float4*>(dst + j);
            c.x += w[j]; c.y += w[j + 1]; c.z += w[j + 2]; c.w += w[j + 3];
            *reinterpret_cast<float4*>(dst + j) = c;
        }
    } else {
        float c[D];
#pragma unroll
        for (int j = 0; j < D; ++j) c[j] = dst[(long long)j * ld];
#pragma unroll
        for (int j = 0; j < D; ++j) dst[(long long)j * ld] = c[j] + w[j];
    }
}

template <int PASSES>
__global__ void __launch_bounds__(TC_NT, 2) k_bwd_meta_tc_wg(const BwdArgs a, const float* __restrict__ wpack, int n_keys) {
    extern __shared__ __align__(128) float smem_raw[];
    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = m.F, L = m.L, l = a.layer;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const int rows = wg_rows(F), ksteps = rows >> 3;
    const int TILE = rows * D;                                    // floats per operand tile
    const uint32_t tile_bytes = (uint32_t)TILE * 4u;
    float* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;      // 1024-byte aligned (swizzle atoms)
    float* sB = smem;                         // B operands of the current data round
    float* sT = sB + TCB_SB;                  // operand tiles: A hi0 hi1 lo0 lo1 | G hi lo
    float* sLN = sT + 6 * TILE;               // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;                  // per warp: dQLN_G dQLN_B dKLN_G dKLN_B
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sdLN + (TC_NT / 32) * 128);   // [0] MMA completion, [1] B operands landed (TMA)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar + 2);
    float* const tAh = sT, * const tAl = sT + 2 * TILE, * const tGh = sT + 4 * TILE, * const tGl = sT + 5 * TILE;

    const int s = tid / F, f = tid - s * F;
    // every CTA stays inside ONE scenario: one set of generated-weight gradients per CTA, reduced over CTAs in a fixed order
    int t0, t1, g_cta, c0_, c1_;
    scenario_partition(a.counters + CNT_HIST, m.n_dom, TC_NT / F, gridDim.x, blockIdx.x, -1, t0, t1, g_cta, c0_, c1_);
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const int nblk = L + n_keys;
    const float* pk_lf = wpack + (long long)l * PACK_FLOATS;                    // forward layer block
    const float* pk_lb = wpack + (long long)(nblk + l) * PACK_FLOATS;           // backward layer block
    // this CTA's rows of the partial buffers (k_dense_reduce adds the rows in CTA order): zeroed here, accumulated tile by tile
    const GPart gp = gpart_layout(m);
    float* part = a.gpart + gp.meta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * GP_META;
    float* tpart = a.gpart + gp.theta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * gp.nside * PSZ;

    copy_to_smem(sLN, m.dense + m.off_layers + (long long)l * LAYER_FLOATS + OFF_LN_G, 192, tid, TC_NT);
    for (int i = tid * 4; i < 3072; i += TC_NT * 4) *reinterpret_cast<float4*>(part + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid * 4; i < gp.nside * PSZ; i += TC_NT * 4) *reinterpret_cast<float4*>(tpart + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 6 * TILE; i += TC_NT) sT[i] = 0.f;      // (the don't-care blocks of the M = 128 products must be finite)
    for (int i = tid; i < (TC_NT / 32) * 128; i += TC_NT) sdLN[i] = 0.f;
    if (warp == 0) tmem_alloc(s_tmem, TC_COLS);
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0, bphase = 0;           // bphase: parity of the B-operand barrier (every round waits for exactly one load)
    float* const sdLNw = sdLN + warp * 128;   // this warp's LayerNorm-gradient slots
    const bool stage = tid < rows;            // owns a row of the operand tiles (zeros past the last token)

    // one round: [TMEM operands + operand tiles stored by all] -> sync -> thread 0 waits for the B operands, issues the data GEMM(s)
    // and the weight-gradient GEMM riding with them, commits -> all wait
#define SAT_TC_ROUND(ISSUE)                                  \
    do {                                                     \
        tc_wait_st();                                        \
        fence_async_smem();                                  \
        tc_fence_before();                                   \
        __syncthreads();                                     \
        if (tid == 0) {                                      \
            mbar_wait(mbar + 1, bphase);                     \
            tc_fence_after();                                \
            ISSUE;                                           \
            mma_commit(mbar);                                \
        }                                                    \
        bphase ^= 1u;                                        \
        mbar_wait(mbar, phase);                              \
        phase ^= 1u;                                         \
        tc_fence_after();                                    \
    } while (0)

    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int g = td.x, start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        const int kq = theta_key(g, l, 0, L, pos), kk = theta_key(g, l, 1, L, pos);

        // Rolled over the Q / K sides (half the code of an unrolled body: the kernel was instruction-fetch bound).  The
        // side's d(x W) goes back to its dqkv slot in place and is re-read for the projection phase below.
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const bool has_meta = side ? meta_k : meta_q;
            float* dqkv_s = a.dqkv + (long long)side * a.act_stride + tokpos;
            if (!has_meta) continue;                                                        // dqkv slot already holds d(x W)
            const int key = side ? kk : kq;
            const float* pk_kf = wpack + (long long)(L + key) * PACK_FLOATS;             // forward key block
            const float* pk_kb = wpack + (long long)(nblk + L + key) * PACK_FLOATS;      // backward key block
            float* dTh = tpart + ((side && pos) ? PSZ : 0);                              // this side's d theta: [W1 32x64 | W2 64x32]
            const float* lng = (side && pos) ? sLN + 128 : sLN + 64;
            float* sdLNs = (side && pos) ? sdLNw + 64 : sdLNw;

            float du[D], zh[D];
            if (valid) {
                load_row(dqkv_s, du);                                                       // dLoss/dQ' (post-LN)
                load_row(act_l + (side ? ACT_ZK : ACT_ZQ) * a.act_stride + tokpos, zh);     // requested before the first round
            } else { zero(du); zero(zh); }
            float u[D];                                                                    // u = x @ W (recomputed)
            {
                float xin[D];
                if (valid) load_row(act_l + ACT_X * a.act_stride + tokpos, xin); else zero(xin);
                load_b2(sB, pk_lf + PK_QKV_HI + side * 1024, sB + 1024, pk_lf + PK_QKV_LO + side * 1024, 1024, tid, mbar + 1);
                tmem_st_operand<PASSES>(tlane + TCB_A, xin);
                SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB, sB + 1024, 32, false));
                tmem_ld32(tlane + TCB_D, u);
                tc_wait_ld();
            }
            float dm[D];
            {   // LayerNorm backward of the MetaNet output
                float rstd = 0.f;
                if (valid) rstd = normalize32(zh);
                layernorm_bwd32(du, zh, rstd, lng, sdLNs, sdLNs + 32, lane, valid);         // du := dLoss/dz
                const DropCtx dc = make_drop(seed32, token_gid, l, side ? SITE_K : SITE_Q, a.p_drop, true);
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float f0, f1;
                    drop_pair(dc, i >> 1, f0, f1);
                    dm[i] = valid ? du[i] * f0 : 0.f; dm[i + 1] = valid ? du[i + 1] * f1 : 0.f;
                }
            }
            uint2 mk = make_uint2(0u, 0u);
            if (valid) mk = *reinterpret_cast<const uint2*>(relu_mask_ptr(const_cast<float*>(a.acts), a.act_stride, L, l, side, tokpos / D));

            // ---- pre = u @ W1 (N = 64): hidden activations h (ReLU decisions from the forward) ----------------------------------
            load_b(sB, pk_kf + PK_W1_HI, 4096, tid, mbar + 1);
            tmem_st_operand<PASSES>(tlane + TCB_A, u);
            SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PK_W1_HI, sB + PK_W1_LO, 64, false));
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float h[D];
                tmem_ld32(tlane + TCB_D + hh * 32, h);
                tc_wait_ld();
                const uint32_t mb = hh ? mk.y : mk.x;
#pragma unroll
                for (int i = 0; i < D; ++i) h[i] = ((mb >> i) & 1u) ? h[i] : 0.f;
                if (stage) wg_stage(tAh + hh * TILE, tAl + hh * TILE, tid, h);
            }
            if (stage) wg_stage(tGh, tGl, tid, dm);
            // ---- dh = dm @ W2^T (N = 64), with dW2^T... rows: dW2[c][j] = sum_t h[t][c] dm[t][j] riding in the same commit -------
            load_b(sB, pk_kb + PKB_W2T_HI, 4096, tid, mbar + 1);
            tmem_st_operand<PASSES>(tlane + TCB_A, dm);
            SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_W2T_HI, sB + PKB_W2T_LO, 64, false),
                          issue_wgrad(tbase + TCB_H, tAh, tAl, tGh, tGl, ksteps, tile_bytes)));
            wg_flush(tlane + TCB_H, dTh + D * HID + tid * D, 1, tid < HID);      // dW2 row c = lane (before H is overwritten below)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float dh[D];
                tmem_ld32(tlane + TCB_D + hh * 32, dh);
                tc_wait_ld();
                const uint32_t mb = hh ? mk.y : mk.x;
#pragma unroll
                for (int i = 0; i < D; ++i) dh[i] = ((mb >> i) & 1u) ? dh[i] : 0.f;
                tmem_st_operand<PASSES>(tlane + (hh ? TCB_H2 : TCB_H), dh);
                if (stage) wg_stage(tAh + hh * TILE, tAl + hh * TILE, tid, dh);
            }
            if (stage) wg_stage(tGh, tGl, tid, u);
            // ---- du += dh @ W1^T, with dW1^T[c][i] = sum_t dh[t][c] u[t][i] riding in the same commit --------------------------
            load_b(sB, pk_kb + PKB_W1N0_HI, 4096, tid, mbar + 1);
            SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB, sB + 1024, 32, false),
                          issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H2, sB + 2048, sB + 3072, 32, true),
                          issue_wgrad(tbase + TCB_D + 32, tAh, tAl, tGh, tGl, ksteps, tile_bytes)));
            {
                float add[D];
                tmem_ld32(tlane + TCB_D, add);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < D; ++i) du[i] += add[i];
            }
            if (valid) store_row(dqkv_s, du);
            wg_flush(tlane + TCB_D + 32, dTh + tid, HID, tid < HID);             // dW1[i][c]: column c = lane of the 32 x 64 matrix
        }
        // ---- projections: dX = dX_res + du_q W_Q^T + du_k W_K^T + dv W_V^T; dW_Q / dW_K / dW_V (transposed products, as above) ----
        float dv[D], res[D];                                     // requested now, used later
        {
            float v[D];
            // (written back by this thread at the end of each side: coherent loads, see load_row_cg)
            if (valid) load_row_cg(a.dqkv + tokpos, v); else zero(v);
            if (stage) wg_stage(tAh, tAl, tid, v);
            tmem_st_operand<PASSES>(tlane + TCB_A, v);
            if (valid) load_row_cg(a.dqkv + a.act_stride + tokpos, v); else zero(v);
            if (stage) wg_stage(tAh + TILE, tAl + TILE, tid, v);
            tmem_st_operand<PASSES>(tlane + TCB_H, v);
            if (valid) load_row(act_l + ACT_X * a.act_stride + tokpos, v); else zero(v);
            if (stage) wg_stage(tGh, tGl, tid, v);
            if (valid) { load_row(a.dqkv + 2 * a.act_stride + tokpos, dv); load_row(a.dx + tokpos, res); }
            else { zero(dv); zero(res); }
        }
        load_b(sB, pk_lb + PKB_WQN_HI, 4096, tid, mbar + 1);
        SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_WQN_HI, sB + PKB_WQN_LO, 32, false),
                      issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB + PKB_WKN_HI, sB + PKB_WKN_LO, 32, true),
                      issue_wgrad(tbase + TCB_D + 32, tAh, tAl, tGh, tGl, ksteps, tile_bytes)));
        // rows 0..31 = dW_Q^T, 32..63 = dW_K^T: dW[i][o] with o = lane & 31 -> a column of the 32 x 32 matrix
        wg_flush(tlane + TCB_D + 32, part + (tid >> 5) * 1024 + (tid & 31), D, tid < 2 * D);
        if (stage) wg_stage(tAh, tAl, tid, dv);                  // (block 1 keeps du_k: its rows of D are not read)
        tmem_st_operand<PASSES>(tlane + TCB_H2, dv);
        load_b(sB, pk_lb + PKB_WVN_HI, 2048, tid, mbar + 1);
        SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H2, sB, sB + 1024, 32, true),
                      issue_wgrad(tbase + TCB_D + 32, tAh, tAl, tGh, tGl, ksteps, tile_bytes)));
        float dxa[D];
        tmem_ld32(tlane + TCB_D, dxa);
        tc_wait_ld();
        if (valid) {
#pragma unroll
            for (int i = 0; i < D; ++i) dxa[i] += res[i];
            // layer 0: this is dLoss/dE0 of the token; sat_rows_reduce sums it per touched table row in sorted (row, token) order
            store_row(a.dx + tokpos, dxa);
        }
        wg_flush(tlane + TCB_D + 32, part + OFF_WV + tid, D, tid < D);
    }
#undef SAT_TC_ROUND
    __syncthreads();
    store_partial_warps<TC_NT>(part + 3072, sdLN, 128, tid);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}

}  // namespace sat

using namespace sat;

int sat_launch_bwd_meta_tc_wg(const BwdArgs& a, const float* wpack, int n_keys, int grid, int passes, cudaStream_t s) {
    const size_t smem = tcb_smem_bytes(a.m.F);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_bwd_meta_tc_wg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_bwd_meta_tc_wg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_bwd_meta_tc_wg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    if (passes == 3) k_bwd_meta_tc_wg<3><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    else if (passes == 2) k_bwd_meta_tc_wg<2><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    else k_bwd_meta_tc_wg<1><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    return sat_host::check_launch("k_bwd_meta_tc_wg");
}
