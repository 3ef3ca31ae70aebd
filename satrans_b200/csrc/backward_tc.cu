// Tensor-core version of k_bwd_meta (MetaNet-LayerNorm bwd + MetaNet bwd + QKV projection bwd), sm_100a.
//
// Same thread-per-token / 128-lane mapping as forward_tc.cu; BOTH kinds of GEMM run as tcgen05.mma.kind::tf32
// (3xTF32 error-compensated with PASSES = 3):
//  * DATA gradients / recomputes: A = the tokens' vectors written by their owning threads into TMEM, B = pre-packed
//    weights (backward blocks of k_pack_weights) in shared memory:
//        u = x W (recompute)   pre = u W1 (recompute; the ReLU decisions come from the forward)
//        dh = dm W2^T          du += dh W1^T          dX += [du_q | du_k | dv] [W_Q | W_K | W_V]^T
//  * WEIGHT gradients dW = A^T G (reduction over the tile's 128 tokens = the MMA's K dimension): both operands are
//    transposed by their owning threads into shared memory in a K-major UMMA layout (hi / lo split), one MMA
//    chain of 16 k-steps x 3 products lands the 32 x 32 block in TMEM, warp 0 drains it into a CTA-private scratch
//    accumulator in L2 that is flushed with one atomicAdd per element per CTA.
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {

constexpr int TCB_A = 0, TCB_H = 64, TCB_D = 128, TCB_W = 192;           // TMEM column map (256 columns)
constexpr int TCB_SB = 4096;                                             // floats of the B-operand staging buffer (16 KB)
constexpr int TCB_SMEM_FLOATS = TCB_SB + 4 * WG_TILE_FLOATS + 192 + 128 + 16;
constexpr int WACC_DW = 0, WACC_THQ = 3072, WACC_THK = 3072 + PSZ, WACC_FLOATS = 3072 + 2 * PSZ;    // per-CTA scratch

template <int PASSES>
__global__ void __launch_bounds__(TC_NT, 2) k_bwd_meta_tc(const BwdArgs a, const float* __restrict__ wpack, int n_keys,
                                                          float* __restrict__ wacc_all) {
    extern __shared__ __align__(128) float smem[];
    float* sB = smem;                                  // B operands of the current data round (<= 16 KB)
    float* sAh = sB + TCB_SB;                          // weight-gradient operand tiles (MN-major, hi / lo)
    float* sAl = sAh + WG_TILE_FLOATS;
    float* sGh = sAl + WG_TILE_FLOATS;
    float* sGl = sGh + WG_TILE_FLOATS;
    float* sLN = sGl + WG_TILE_FLOATS;                 // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;                           // dQLN_G dQLN_B dKLN_G dKLN_B
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sdLN + 128);
    uint64_t* wbar = mbar + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(wbar + 1);

    const SatModel& m = a.m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = m.F, L = m.L, l = a.layer;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const int n_tiles = a.counters[CNT_NTILES2];
    const int s = tid / F, f = tid - s * F;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const int nblk = L + n_keys;
    const float* pk_lf = wpack + (long long)l * PACK_FLOATS;                    // forward layer block
    const float* pk_lb = wpack + (long long)(nblk + l) * PACK_FLOATS;           // backward layer block
    float* wacc = wacc_all + (long long)blockIdx.x * WACC_FLOATS;               // CTA-private weight-gradient scratch (L2)

    copy_to_smem(sLN, m.dense + m.off_layers + (long long)l * LAYER_FLOATS + OFF_LN_G, 192, tid, TC_NT);
    for (int i = tid; i < WACC_FLOATS; i += TC_NT) wacc[i] = 0.f;
    for (int i = tid; i < 4 * WG_TILE_FLOATS; i += TC_NT) sAh[i] = 0.f;         // rows of absent tokens must be finite
    if (tid < 128) sdLN[tid] = 0.f;
    if (warp == 0) tmem_alloc(s_tmem, TC_COLS);
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(wbar, 1); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0, wphase = 0;
    int cur_q = -1, cur_k = -1;
    float* wg_dst = nullptr;        // where the pending weight-gradient block goes (row 0), nullptr = none pending
    int wg_ld = 0, wg_count = 0;

#define SAT_TC_ROUND(ISSUE)                                  \
    do {                                                     \
        fence_async_smem();                                  \
        tc_wait_st();                                        \
        tc_fence_before();                                   \
        __syncthreads();                                     \
        if (tid == 0) {                                      \
            tc_fence_after();                                \
            ISSUE;                                           \
            mma_commit(mbar);                                \
        }                                                    \
        mbar_wait(mbar, phase);                              \
        phase ^= 1u;                                         \
        tc_fence_after();                                    \
    } while (0)

    // wait for the pending weight-gradient MMA chain and add its 32 x 32 block (TMEM lanes 0..31) into the scratch
#define SAT_WG_DRAIN()                                                                        \
    do {                                                                                      \
        if (wg_dst != nullptr) {                                                              \
            mbar_wait(wbar, wphase);                                                          \
            wphase ^= 1u;                                                                     \
            tc_fence_after();                                                                 \
            if (warp == 0) {                                                                  \
                float blk[D];                                                                 \
                tmem_ld32(tlane + TCB_W, blk);                                                \
                tc_wait_ld();                                                                 \
                float* row = wg_dst + lane * wg_ld;                                           \
                _Pragma("unroll") for (int i = 0; i < D; i += 4) {                            \
                    float4 o = *reinterpret_cast<float4*>(row + i);                           \
                    o.x += blk[i]; o.y += blk[i + 1]; o.z += blk[i + 2]; o.w += blk[i + 3];   \
                    *reinterpret_cast<float4*>(row + i) = o;                                  \
                }                                                                             \
                tc_fence_before();                                                            \
            }                                                                                 \
            wg_dst = nullptr;                                                                 \
        }                                                                                     \
    } while (0)

    // operands are staged; one elected thread (rotating over the warps) issues the chain
#define SAT_WG_ISSUE(DST, LD)                                                                 \
    do {                                                                                      \
        fence_async_smem();                                                                   \
        tc_fence_before();                                                                    \
        __syncthreads();                                                                      \
        if (tid == (wg_count & 3) * 32) {                                                     \
            tc_fence_after();                                                                 \
            issue_wgrad(tbase + TCB_W, sAh, sAl, sGh, sGl);                                   \
            mma_commit(wbar);                                                                 \
        }                                                                                     \
        ++wg_count;                                                                           \
        wg_dst = (DST);                                                                       \
        wg_ld = (LD);                                                                         \
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
        if (kq != cur_q || (pos && kk != cur_k)) {          // scenario changed: hand the generated-weight gradients over
            SAT_WG_DRAIN();
            __syncthreads();
            if (kq != cur_q) {
                if (cur_q >= 0) flush_add<TC_NT>(a.dtheta + (long long)cur_q * PSZ, wacc + WACC_THQ, PSZ, tid);
                cur_q = kq;
            }
            if (pos && kk != cur_k) {
                if (cur_k >= 0) flush_add<TC_NT>(a.dtheta + (long long)cur_k * PSZ, wacc + WACC_THK, PSZ, tid);
                cur_k = kk;
            }
            __syncthreads();
        }

        float keep_q[D], keep_k[D];          // d(x W_Q), d(x W_K): A operands of the final dX rounds
        zero(keep_q); zero(keep_k);
        float xin[D];
        if (valid) load_row(act_l + ACT_X * a.act_stride + tokpos, xin); else zero(xin);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const bool has_meta = side ? meta_k : meta_q;
            const int key = side ? kk : kq;
            const float* pk_kf = wpack + (long long)(L + key) * PACK_FLOATS;             // forward key block
            const float* pk_kb = wpack + (long long)(nblk + L + key) * PACK_FLOATS;      // backward key block
            float* wTh = wacc + ((side && pos) ? WACC_THK : WACC_THQ);
            const float* lng = (side && pos) ? sLN + 128 : sLN + 64;
            float* sdLNs = (side && pos) ? sdLN + 64 : sdLN;
            float (&du)[D] = side ? keep_k : keep_q;

            if (valid) load_row(a.dqkv + (long long)side * a.act_stride + tokpos, du);     // dLoss/dQ' (post-LN)
            if (has_meta) {
                float u[D];                                                                // u = x @ W (recomputed)
                copy_to_smem(sB, pk_lf + PK_QKV_HI + side * 1024, 1024, tid, TC_NT);
                copy_to_smem(sB + 1024, pk_lf + PK_QKV_LO + side * 1024, 1024, tid, TC_NT);
                tmem_st_operand<PASSES>(tlane + TCB_A, xin);
                SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB, sB + 1024, 32, false));
                tmem_ld32(tlane + TCB_D, u);
                tc_wait_ld();
                float dm[D];
                {   // LayerNorm backward of the MetaNet output
                    float zh[D];
                    float rstd = 0.f;
                    if (valid) { load_row(act_l + (side ? ACT_ZK : ACT_ZQ) * a.act_stride + tokpos, zh); rstd = normalize32(zh); }
                    else zero(zh);
                    layernorm_bwd32(du, zh, rstd, lng, sdLNs, sdLNs + 32, lane, valid);     // du := dLoss/dz
                    const DropCtx dc = make_drop(a.seed, token_gid, l, side ? SITE_K : SITE_Q, a.p_drop, true);
#pragma unroll
                    for (int i = 0; i < D; ++i) dm[i] = valid ? du[i] * drop_factor(dc, i) : 0.f;
                }
                uint2 mk = make_uint2(0u, 0u);
                if (valid) mk = *reinterpret_cast<const uint2*>(relu_mask_ptr(const_cast<float*>(a.acts), a.act_stride, L, l, side, tokpos / D));

                // ---- pre = u @ W1 (N = 64); dW2[half*32 + c][j] += h[c] dm[j] ------------------------------------------
                copy_to_smem(sB, pk_kf + PK_W1_HI, 4096, tid, TC_NT);
                tmem_st_operand<PASSES>(tlane + TCB_A, u);
                SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PK_W1_HI, sB + PK_W1_LO, 64, false));
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float h[D];
                    tmem_ld32(tlane + TCB_D + half * 32, h);
                    tc_wait_ld();
                    const uint32_t mb = half ? mk.y : mk.x;
#pragma unroll
                    for (int i = 0; i < D; ++i) h[i] = ((mb >> i) & 1u) ? h[i] : 0.f;
                    SAT_WG_DRAIN();
                    wg_stage_row(sAh, sAl, tid, h);
                    if (half == 0) wg_stage_row(sGh, sGl, tid, dm);
                    SAT_WG_ISSUE(wTh + D * HID + half * D * D, D);
                }
                // ---- dh = dm @ W2^T (N = 64), masked; dW1[i][half*32 + c] += u[i] dh[c]; du += dh @ W1^T -----------------
                copy_to_smem(sB, pk_kb + PKB_W2T_HI, 4096, tid, TC_NT);
                tmem_st_operand<PASSES>(tlane + TCB_A, dm);
                SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_W2T_HI, sB + PKB_W2T_LO, 64, false));
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float dh[D];
                    tmem_ld32(tlane + TCB_D + half * 32, dh);
                    tc_wait_ld();
                    const uint32_t mb = half ? mk.y : mk.x;
#pragma unroll
                    for (int i = 0; i < D; ++i) dh[i] = ((mb >> i) & 1u) ? dh[i] : 0.f;
                    tmem_st_operand<PASSES>(tlane + (half ? TCB_A : TCB_H), dh);            // A operands of the du round
                    SAT_WG_DRAIN();
                    if (half == 0) wg_stage_row(sAh, sAl, tid, u);
                    wg_stage_row(sGh, sGl, tid, dh);
                    SAT_WG_ISSUE(wTh + half * D, HID);
                }
                copy_to_smem(sB, pk_kb + PKB_W1N0_HI, 4096, tid, TC_NT);
                SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB, sB + 1024, 32, false),
                              issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + 2048, sB + 3072, 32, true)));
                {
                    float add[D];
                    tmem_ld32(tlane + TCB_D, add);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < D; ++i) du[i] += add[i];
                }
            }
            // ---- projection weight gradient: dW[i][j] += x[i] du[j] ----------------------------------------------------
            SAT_WG_DRAIN();
            if (side == 0) wg_stage_row(sAh, sAl, tid, xin);       // x stays staged for the K side and for V
            else if (has_meta) wg_stage_row(sAh, sAl, tid, xin);   // (the MetaNet GEMMs of this side overwrote it)
            wg_stage_row(sGh, sGl, tid, du);
            SAT_WG_ISSUE(wacc + WACC_DW + (side ? OFF_WK : OFF_WQ), D);
        }
        // ---- V: dW_V += x^T dv ; dX = dX_res + du_q W_Q^T + du_k W_K^T + dv W_V^T (two rounds, 16 KB of operands each) ---
        float dv[D];
        if (valid) load_row(a.dqkv + 2 * a.act_stride + tokpos, dv); else zero(dv);
        SAT_WG_DRAIN();
        wg_stage_row(sGh, sGl, tid, dv);                           // sAh / sAl still hold x
        SAT_WG_ISSUE(wacc + WACC_DW + OFF_WV, D);
        copy_to_smem(sB, pk_lb, 4096, tid, TC_NT);                 // W_Q^N, W_K^N (hi | lo each)
        tmem_st_operand<PASSES>(tlane + TCB_A, keep_q);
        tmem_st_operand<PASSES>(tlane + TCB_H, keep_k);
        SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_WQN_HI, sB + PKB_WQN_LO, 32, false),
                      issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB + PKB_WKN_HI, sB + PKB_WKN_LO, 32, true)));
        copy_to_smem(sB, pk_lb + PKB_WVN_HI, 2048, tid, TC_NT);
        tmem_st_operand<PASSES>(tlane + TCB_A, dv);
        SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB, sB + 1024, 32, true));
        float dxa[D];
        tmem_ld32(tlane + TCB_D, dxa);
        tc_wait_ld();
        if (valid) {
            float res[D];
            load_row(a.dx + tokpos, res);
#pragma unroll
            for (int i = 0; i < D; ++i) dxa[i] += res[i];
            if (l > 0) {
                store_row(a.dx + tokpos, dxa);
            } else {
                long long id = load_id(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.sparse_cols + f));
                const long long lo = __ldg(m.row_off + f), hi = __ldg(m.row_off + f + 1);
                if (id < 0 || id >= hi - lo) id = 0;
                const long long row = lo + id;
                const int owner = (int)(row % m.n_shards);
                const long long lrow = row / m.n_shards;
                const uint32_t wbits = *reinterpret_cast<const volatile uint32_t*>(m.shard_bitmap[owner] + (lrow >> 5));
                const int slot = *reinterpret_cast<const volatile int32_t*>(m.shard_word_rank[owner] + (lrow >> 5)) +
                                 __popc(wbits & ((1u << (lrow & 31)) - 1u));
                float* dst = m.shard_row_grad[owner] + (long long)slot * D;
#pragma unroll
                for (int i = 0; i < D; i += 4)
                    atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(dxa[i], dxa[i + 1], dxa[i + 2], dxa[i + 3]));
            }
        }
    }
    SAT_WG_DRAIN();
#undef SAT_TC_ROUND
#undef SAT_WG_DRAIN
#undef SAT_WG_ISSUE
    __syncthreads();
    float* dlay = a.dgrad + m.off_layers + (long long)l * LAYER_FLOATS;
    flush_add<TC_NT>(dlay, wacc + WACC_DW, 3072, tid);
    flush_add<TC_NT>(dlay + OFF_QLN_G, sdLN, 128, tid);
    if (cur_q >= 0) flush_add<TC_NT>(a.dtheta + (long long)cur_q * PSZ, wacc + WACC_THQ, PSZ, tid);
    if (pos && cur_k >= 0) flush_add<TC_NT>(a.dtheta + (long long)cur_k * PSZ, wacc + WACC_THK, PSZ, tid);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}

}  // namespace sat

using namespace sat;

int sat_launch_bwd_meta_tc(const BwdArgs& a, const float* wpack, int n_keys, float* wacc, int grid, int passes, cudaStream_t s) {
    const size_t smem = TCB_SMEM_FLOATS * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_bwd_meta_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_bwd_meta_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = true;
    }
    if (passes == 3) k_bwd_meta_tc<3><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys, wacc);
    else k_bwd_meta_tc<1><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys, wacc);
    return sat_host::check_launch("k_bwd_meta_tc");
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
