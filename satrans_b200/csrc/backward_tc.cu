// Tensor-core version of k_bwd_meta (MetaNet-LayerNorm bwd + MetaNet bwd + QKV projection bwd), sm_100a.
//
// Same thread-per-token / 128-lane mapping as forward_tc.cu.  Every DATA-gradient / recompute projection is a
// tcgen05.mma.kind::tf32 grouped GEMM with the tokens' vectors as A operand in TMEM and pre-packed weights (the
// backward blocks of k_pack_weights) as B operand in shared memory:
//     u = x W                (recompute)        pre = u W1            (recompute, ReLU decisions come from the forward)
//     dh = dm W2^T                              du += dh W1^T                              dX += [du_q|du_k|dv] [W_Q|W_K|W_V]^T
// WEIGHT gradients (reductions over tokens) stay on the CUDA cores: operands staged as [token][36] tiles, 4x4 register
// blocks, shared-memory accumulators flushed once per CTA (wgrad32, as in backward.cu).
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {

constexpr int TCB_A = 0, TCB_H = 64, TCB_D = 128, TCB_H2 = 192;          // TMEM column map (256 columns)
constexpr int TCB_SB = 4096;                                             // floats of the B-operand staging buffer (16 KB)

// Shared memory (floats): B staging | three [tile_rows][36] weight-gradient operand tiles | LN params | dW_Q/K/V | dLN | d theta
// (one key block, two with per-position keys).  tile_rows = (128 / F) * F rounded up to 4: the threads past a tile's last
// token store zero rows, so the weight-gradient loops run over whole groups of 4 tokens.
// F = 19: 97.9 KB (EN) / 114.4 KB (ENP) -> 2 CTAs per SM either way.
static size_t tcb_smem_bytes(int F, bool pos) {
    const int rows = ((TC_NT / F) * F + 3) & ~3;
    return (size_t)(TCB_SB + 3 * rows * KV_LD + 192 + 3072 + (TC_NT / 32) * 128 + (pos ? 2 : 1) * PSZ + 8) * sizeof(float);   // (+8: two mbarriers, TMEM address)
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

template <int PASSES>
__global__ void __launch_bounds__(TC_NT, 2) k_bwd_meta_tc(const BwdArgs a, const float* __restrict__ wpack, int n_keys) {
    extern __shared__ __align__(128) float smem[];
    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hsel = tid >> 6;
    const int F = m.F, L = m.L, l = a.layer;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const int tile_rows = ((TC_NT / F) * F + 3) & ~3;
    float* sB = smem;                         // B operands of the current round
    float* sT0 = sB + TCB_SB;                 // weight-gradient operand tiles
    float* sT1 = sT0 + tile_rows * KV_LD;
    float* sT2 = sT1 + tile_rows * KV_LD;
    float* sLN = sT2 + tile_rows * KV_LD;     // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdW = sLN + 192;                   // dWQ dWK dWV
    float* sdLN = sdW + 3072;                 // per warp: dQLN_G dQLN_B dKLN_G dKLN_B
    float* sdThQ = sdLN + (TC_NT / 32) * 128;
    float* sdThK = sdThQ + PSZ;               // only with per-position keys
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sdThQ + (pos ? 2 : 1) * PSZ);   // [0] MMA completion, [1] B operands landed (TMA)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar + 2);

    const int s = tid / F, f = tid - s * F;
    // every CTA stays inside ONE scenario: one set of generated-weight gradients per CTA, reduced over CTAs in a fixed order
    int t0, t1, g_cta, c0_, c1_;
    scenario_partition(a.counters + CNT_HIST, m.n_dom, TC_NT / F, gridDim.x, blockIdx.x, -1, t0, t1, g_cta, c0_, c1_);
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const int nblk = L + n_keys;
    const float* pk_lf = wpack + (long long)l * PACK_FLOATS;                    // forward layer block
    const float* pk_lb = wpack + (long long)(nblk + l) * PACK_FLOATS;           // backward layer block

    copy_to_smem(sLN, m.dense + m.off_layers + (long long)l * LAYER_FLOATS + OFF_LN_G, 192, tid, TC_NT);
    for (int i = tid; i < 3072; i += TC_NT) sdW[i] = 0.f;
    for (int i = tid; i < (pos ? 2 : 1) * PSZ; i += TC_NT) sdThQ[i] = 0.f;
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
    float* const sTh = hsel ? sT1 : sT0;      // this 64-thread half's tile in the paired weight-gradient phases

#define SAT_TC_ROUND(ISSUE)                                  \
    do {                                                     \
        tc_wait_st();                                        \
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
        const int ntok4 = (cnt * F + 3) & ~3;
        const bool stage = tid < tile_rows;       // owns a row of the operand tiles (zeros past the last token)
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        const int kq = theta_key(g, l, 0, L, pos), kk = theta_key(g, l, 1, L, pos);
        __syncthreads();                       // previous tile's weight-gradient phases are done with the tiles / accumulators

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
            float* sdTh = (side && pos) ? sdThK : sdThQ;
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

            // ---- pre = u @ W1 (N = 64): hidden activations; dW2[half*32 + c][j] += h[c] dm[j], both halves at once ------------
            load_b(sB, pk_kf + PK_W1_HI, 4096, tid, mbar + 1);
            tmem_st_operand<PASSES>(tlane + TCB_A, u);
            SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PK_W1_HI, sB + PK_W1_LO, 64, false));
            if (stage) sts_row(sT2 + tid * KV_LD, dm);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float h[D];
                tmem_ld32(tlane + TCB_D + hh * 32, h);
                tc_wait_ld();
                const uint32_t mb = hh ? mk.y : mk.x;
#pragma unroll
                for (int i = 0; i < D; ++i) h[i] = ((mb >> i) & 1u) ? h[i] : 0.f;
                if (stage) sts_row((hh ? sT1 : sT0) + tid * KV_LD, h);
            }
            __syncthreads();
            wgrad_half(sTh, sT2, ntok4, sdTh + D * HID + hsel * D * D, D, tid);
            // ---- dh = dm @ W2^T (N = 64), masked; dW1[i][half*32 + c] += u[i] dh[c]; du += dh @ W1^T ---------------------------
            load_b(sB, pk_kb + PKB_W2T_HI, 4096, tid, mbar + 1);
            tmem_st_operand<PASSES>(tlane + TCB_A, dm);
            SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_W2T_HI, sB + PKB_W2T_LO, 64, false));
            if (stage) sts_row(sT2 + tid * KV_LD, u);            // (the round's barrier: dW2 phase is done with the tiles)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float dh[D];
                tmem_ld32(tlane + TCB_D + hh * 32, dh);
                tc_wait_ld();
                const uint32_t mb = hh ? mk.y : mk.x;
#pragma unroll
                for (int i = 0; i < D; ++i) dh[i] = ((mb >> i) & 1u) ? dh[i] : 0.f;
                tmem_st_operand<PASSES>(tlane + (hh ? TCB_H2 : TCB_H), dh);
                if (stage) sts_row((hh ? sT1 : sT0) + tid * KV_LD, dh);
            }
            __syncthreads();
            wgrad_half(sT2, sTh, ntok4, sdTh + hsel * D, HID, tid);
            load_b(sB, pk_kb + PKB_W1N0_HI, 4096, tid, mbar + 1);
            SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB, sB + 1024, 32, false),
                          issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H2, sB + 2048, sB + 3072, 32, true)));
            {
                float add[D];
                tmem_ld32(tlane + TCB_D, add);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < D; ++i) du[i] += add[i];
            }
            if (valid) store_row(dqkv_s, du);
        }
        // ---- projections: dW_Q += x^T du_q and dW_K += x^T du_k (paired), dW_V += x^T dv;
        //      dX = dX_res + du_q W_Q^T + du_k W_K^T + dv W_V^T -----------------------------------------------------------------
        __syncthreads();                                         // tiles free (a side without MetaNet has no round barrier)
        float dv[D], res[D];                                     // requested now, used after the paired phase / at the end
        {
            float v[D];
            // (written back by this thread at the end of each side: coherent loads, see load_row_cg)
            if (valid) load_row_cg(a.dqkv + tokpos, v); else zero(v);
            if (stage) sts_row(sT0 + tid * KV_LD, v);
            tmem_st_operand<PASSES>(tlane + TCB_A, v);
            if (valid) load_row_cg(a.dqkv + a.act_stride + tokpos, v); else zero(v);
            if (stage) sts_row(sT1 + tid * KV_LD, v);
            tmem_st_operand<PASSES>(tlane + TCB_H, v);
            if (valid) load_row(act_l + ACT_X * a.act_stride + tokpos, v); else zero(v);
            if (stage) sts_row(sT2 + tid * KV_LD, v);
            if (valid) { load_row(a.dqkv + 2 * a.act_stride + tokpos, dv); load_row(a.dx + tokpos, res); }
            else { zero(dv); zero(res); }
        }
        load_b(sB, pk_lb + PKB_WQN_HI, 4096, tid, mbar + 1);
        __syncthreads();
        wgrad_half(sT2, sTh, ntok4, sdW + (hsel ? OFF_WK : OFF_WQ), D, tid);
        SAT_TC_ROUND((issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_A, sB + PKB_WQN_HI, sB + PKB_WQN_LO, 32, false),
                      issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H, sB + PKB_WKN_HI, sB + PKB_WKN_LO, 32, true)));
        if (stage) sts_row(sT0 + tid * KV_LD, dv);               // (the round's barrier: paired phase is done with sT0)
        tmem_st_operand<PASSES>(tlane + TCB_H2, dv);
        load_b(sB, pk_lb + PKB_WVN_HI, 2048, tid, mbar + 1);
        __syncthreads();
        // dW_V: one product for 128 threads -> each half takes half of the token groups; the upper half parks its partial
        // block in the free part of sB and the lower half adds both to the accumulator after the round's barrier
        float cv[4][4];
        {
            const int mid = ((ntok4 >> 1) + 3) & ~3;
            wgrad_block(sT2, sT0, hsel ? mid : 0, hsel ? ntok4 : mid, tid, cv);
            if (hsel) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    *reinterpret_cast<float4*>(sB + 2048 + (tid & 63) * 16 + r * 4) = make_float4(cv[r][0], cv[r][1], cv[r][2], cv[r][3]);
                fence_async_smem();        // these generic-proxy stores are ordered before the next TMA write of this part of sB
            }
        }
        SAT_TC_ROUND(issue_gemm_k32<PASSES>(tbase + TCB_D, tbase + TCB_H2, sB, sB + 1024, 32, true));
        if (!hsel) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 o = *reinterpret_cast<const float4*>(sB + 2048 + (tid & 63) * 16 + r * 4);
                cv[r][0] += o.x; cv[r][1] += o.y; cv[r][2] += o.z; cv[r][3] += o.w;
            }
            wgrad_store(sdW + OFF_WV, D, tid, cv);
        }
        float dxa[D];
        tmem_ld32(tlane + TCB_D, dxa);
        tc_wait_ld();
        if (valid) {
#pragma unroll
            for (int i = 0; i < D; ++i) dxa[i] += res[i];
            // layer 0: this is dLoss/dE0 of the token; sat_rows_reduce sums it per touched table row in sorted (row, token) order
            store_row(a.dx + tokpos, dxa);
        }
    }
#undef SAT_TC_ROUND
    __syncthreads();
    {   // this CTA's rows of the partial buffers (k_dense_reduce adds the rows in CTA order)
        const GPart gp = gpart_layout(m);
        float* part = a.gpart + gp.meta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * GP_META;
        store_partial<TC_NT>(part, sdW, 3072, tid);
        store_partial_warps<TC_NT>(part + 3072, sdLN, 128, tid);
        float* tpart = a.gpart + gp.theta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * gp.nside * PSZ;
        store_partial<TC_NT>(tpart, sdThQ, gp.nside * PSZ, tid);          // sdThQ | sdThK are contiguous
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}

}  // namespace sat

using namespace sat;

int sat_launch_bwd_meta_tc(const BwdArgs& a, const float* wpack, int n_keys, int grid, int passes, cudaStream_t s) {
    const size_t smem = tcb_smem_bytes(a.m.F, a.m.flags & SAT_FLAG_POS);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_bwd_meta_tc<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_bwd_meta_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_bwd_meta_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    if (passes == 3) k_bwd_meta_tc<3><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    else if (passes == 2) k_bwd_meta_tc<2><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    else k_bwd_meta_tc<1><<<grid, TC_NT, smem, s>>>(a, wpack, n_keys);
    return sat_host::check_launch("k_bwd_meta_tc");
}
