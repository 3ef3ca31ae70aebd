// Fused SATrans forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same mapping as forward.cu -- one thread = one field token, a CTA = whole samples of ONE scenario -- but a
// CTA is 128 threads = the 128 TMEM lanes of a cta_group::1 MMA with M = 128: lane i IS token i.  Every
// projection of the layer (QKV, the scenario-generated MetaNet W1/W2 of the Q and K sides, the output
// projection) is a grouped GEMM  [128 tokens x K] @ [K x N]  issued by ONE thread as tcgen05.mma.kind::tf32:
//   * A operand: the tokens' activations, written by their owning threads straight into TMEM
//     (tcgen05.st 32x32b: thread t writes lane t) -- no shared-memory staging, no bank conflicts;
//   * B operand: the CTA-uniform weights in shared memory in the canonical K-major no-swizzle UMMA layout,
//     pre-packed once per step by k_pack_weights (and, for the MetaNet, once per scenario key) and fetched by TMA
//     (cp.async.bulk.tensor through a 2-D tensor map over the packed buffer, 8 KB boxes, mbarrier transaction counts):
//     the MMA-issuing thread prefetches the next layer's QKV block as soon as this layer's QKV MMAs have completed and
//     the next layer's W_out block after this layer's output projection, so the copies ride under the attention math
//     and no thread ever touches a weight byte;
//   * D accumulator: TMEM; tcgen05.ld 32x32b hands each thread ITS token's output row back in registers,
//     where LayerNorm / ReLU / residual / dropout / softmax stay thread-local exactly as in forward.cu.
// Precision: PASSES = 3 is the error-compensated 3xTF32 scheme (x = hi + lo, D = A_hi B_hi + A_lo B_hi + A_hi B_lo),
// which keeps fp32-level parity with the reference; PASSES = 1 is plain TF32 (fast mode, reported separately).
// Attention over the F <= 32 tokens of a sample (9 % of the FLOPs, F x F x 8 per head) stays on the CUDA cores.
#include <cuda_bf16.h>

#include "fwd_common.cuh"
#include "bwd_common.cuh"
#include "tc_common.cuh"

namespace sat {


// ---------------------------------------------------------------------------------------------------------
// weight packing: layer weights + generated MetaNet weights -> UMMA B operands (hi / lo split)
// ---------------------------------------------------------------------------------------------------------
#if SAT_INST_H == 4      // one copy of the packing kernel (this file is compiled once per head count)
__global__ void __launch_bounds__(256) k_pack_weights(const SatModel m, int n_keys, const float* __restrict__ theta,
                                                      float* __restrict__ wpack, int bf16) {
    const int nblk = m.L + n_keys;
    const int total = 2 * nblk * 4096;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const int blk = e >> 12, i = e & 4095;
        float v;
        int base_hi, base_lo, pn, pk_;
        if (blk < m.L) {
            const float* lay = m.dense + m.off_layers + (long long)blk * LAYER_FLOATS;
            if (i < 3072) {                      // B[n][k] = W_{Q|K|V}[k][n']   (x @ W)
                const int n = i >> 5, k = i & 31;
                v = lay[(n >> 5) * 1024 + k * 32 + (n & 31)];
                base_hi = PK_QKV_HI; base_lo = PK_QKV_LO; pn = n; pk_ = k;
            } else {                             // B[n][k] = Out_linear.weight[n][k]   (r @ W^T)
                const int j = i - 3072, n = j >> 5, k = j & 31;
                v = lay[OFF_WO + n * 32 + k];
                base_hi = PK_WO_HI; base_lo = PK_WO_LO; pn = n; pk_ = k;
            }
        } else if (blk < nblk) {
            const float* th = theta + (long long)(blk - m.L) * PSZ;
            if (i < 2048) {                      // W1: B[n][k] = theta[k*64 + n], N = 64
                const int n = i >> 5, k = i & 31;
                v = th[k * HID + n];
                base_hi = PK_W1_HI; base_lo = PK_W1_LO; pn = n; pk_ = k;
            } else {                             // W2 halves: B[n][k] = theta[2048 + (half*32 + k)*32 + n], N = 32
                const int j = i - 2048, half = j >> 10, jj = j & 1023, n = jj >> 5, k = jj & 31;
                v = th[D * HID + (half * 32 + k) * D + n];
                base_hi = (half ? PK_W2B_HI : PK_W2A_HI); base_lo = (half ? PK_W2B_LO : PK_W2A_LO); pn = n; pk_ = k;
            }
        } else if (blk < nblk + m.L) {           // backward layer block
            const float* lay = m.dense + m.off_layers + (long long)(blk - nblk) * LAYER_FLOATS;
            const int which = i >> 10, j = i & 1023, n = j >> 5, k = j & 31;
            if (which < 3) {                     // B[n=i][k=j] = W[i][j]
                v = lay[which * 1024 + n * 32 + k];
                base_hi = which * 2048; base_lo = which * 2048 + 1024; pn = n; pk_ = k;
            } else {                             // B[n=i][k=j] = W_out[j][i]
                v = lay[OFF_WO + k * 32 + n];
                base_hi = PKB_WOT_HI; base_lo = PKB_WOT_LO; pn = n; pk_ = k;
            }
        } else {                                 // backward key block
            const float* th = theta + (long long)(blk - nblk - m.L) * PSZ;
            if (i < 2048) {                      // W2T: B[n=c][k=j] = W2[c][j], N = 64
                const int n = i >> 5, k = i & 31;
                v = th[D * HID + n * D + k];
                base_hi = PKB_W2T_HI; base_lo = PKB_W2T_LO; pn = n; pk_ = k;
            } else {                             // W1N halves: B[n=i][k=c] = W1[i][half*32 + c]
                const int j = i - 2048, half = j >> 10, jj = j & 1023, n = jj >> 5, k = jj & 31;
                v = th[n * HID + half * 32 + k];
                base_hi = (half ? PKB_W1N1_HI : PKB_W1N0_HI); base_lo = (half ? PKB_W1N1_LO : PKB_W1N0_LO); pn = n; pk_ = k;
            }
        }
        float* dst = wpack + (long long)blk * PACK_FLOATS;
        if (bf16) {                 // bf16: groups of 32 rows, group g in the first half of floats [1024 g, 1024 g + 1024) of the "hi" region
            reinterpret_cast<__nv_bfloat16*>(dst + base_hi + (pn >> 5) * 1024)[b_off16(pn & 31, pk_)] = __float2bfloat16_rn(v);
        } else {
            const float hi = tf32_hi(v);
            dst[base_hi + b_off32(pn, pk_)] = hi;
            dst[base_lo + b_off32(pn, pk_)] = v - hi;
        }
    }
}
#endif


constexpr int TC_HEAD_FLOATS = SAT_MAXF * D + 64 + 8;     // d dnn_linear accumulators: F*32 weights | n_dense weights | bias
constexpr int TC_SMEM_FLOATS = 2 * PACK_FLOATS + 2 * TC_NT * KV_LD + 192 + 2 * TC_NT + TC_HEAD_FLOATS + 16;
// mbarriers: [0] MMA completion (tcgen05.commit), [1] QKV block landed, [2] W_out block landed, [3] key block landed
constexpr uint32_t TS_QKV = 1u, TS_WO = 2u, TS_KEY = 4u, TS_KEY_PENDING = 8u;    // phase bits / flags of the issuing thread

template <int H, bool TRAIN, int PASSES>
__global__ void __launch_bounds__(TC_NT, 2) k_forward_tc(const FwdArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smem[];
    float* sBL = smem;                            // packed layer block   (32 KB)
    float* sBT = sBL + PACK_FLOATS;               // packed key block     (32 KB)
    float* sK = sBT + PACK_FLOATS;                // [128][36]
    float* sV = sK + TC_NT * KV_LD;               // [128][36]
    float* sLN = sV + TC_NT * KV_LD;              // 6 x 32 LayerNorm vectors
    float* sRed = sLN + 192;                      // [128]
    float* sDz = sRed + TC_NT;                    // [128] dLoss/dlogit of the tile's samples (head backward)
    float* sHead = sDz + TC_NT;                   // this CTA's sums of d dnn_linear over its tiles
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sHead + TC_HEAD_FLOATS);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar + 4);
    __shared__ double s_lossv[TC_NT];      // per-sample loss terms of the tile, added in sample order (no atomics)

    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int F = m.F, L = m.L;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool relu_out = m.flags & SAT_FLAG_RELU, use_res = m.flags & SAT_FLAG_RES;
    const int n_tiles = *a.ntiles_ptr;
    const int s = tid / F, f = tid - s * F;

    const bool head_bwd = TRAIN && a.y != nullptr && a.dx != nullptr;     // the logit head's backward is folded into this kernel
    for (int i = tid; i < TC_HEAD_FLOATS; i += TC_NT) sHead[i] = 0.f;
    if (warp == 0) tmem_alloc(s_tmem, TC_COLS);
    if (tid == 0) {
        mbar_init(mbar + 0, 1); mbar_init(mbar + 1, 1); mbar_init(mbar + 2, 1); mbar_init(mbar + 3, 1);
        tma_prefetch_desc(&tmap);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *s_tmem;
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes
    uint32_t phase = 0;
    uint32_t ts = 0u;             // issuing thread: phase bits of the three TMA barriers + "key block in flight"
    int cur_key = -1;

    // TMA fetches (issuing thread only).  Packed layer block l = rows [32 l, 32 l + 32) of the tensor map: QKV hi|lo in rows
    // 0..23, W_out hi|lo in rows 24..31; packed key block k = rows [32 (L + k), ...).
    auto fetch_qkv = [&](int l) {
        mbar_expect_tx(mbar + 1, 3 * TMA_BOX_FLOATS * 4);
#pragma unroll
        for (int j = 0; j < 3; ++j) tma_load_2d(sBL + j * TMA_BOX_FLOATS, &tmap, 0, l * PACK_ROWS + j * TMA_BOX_ROWS, mbar + 1);
    };
    auto fetch_wo = [&](int l) {
        mbar_expect_tx(mbar + 2, TMA_BOX_FLOATS * 4);
        tma_load_2d(sBL + PK_WO_HI, &tmap, 0, l * PACK_ROWS + 3 * TMA_BOX_ROWS, mbar + 2);
    };
    auto fetch_key = [&](int k) {
        mbar_expect_tx(mbar + 3, 4 * TMA_BOX_FLOATS * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(sBT + j * TMA_BOX_FLOATS, &tmap, 0, (L + k) * PACK_ROWS + j * TMA_BOX_ROWS, mbar + 3);
        ts |= TS_KEY_PENDING;
    };
    // the issuing thread waits for a weight block right before the first MMA that reads it
    auto wait_block = [&](int which, uint32_t bit) {
        mbar_wait(mbar + which, (ts & bit) ? 1u : 0u);
        ts ^= bit;
    };
    auto wait_key = [&]() {
        if (ts & TS_KEY_PENDING) { wait_block(3, TS_KEY); ts &= ~TS_KEY_PENDING; }
    };
    if (tid == 0 && (int)blockIdx.x < n_tiles) { fetch_qkv(0); fetch_wo(0); }

    // one MMA round trip: [operands stored by all] -> sync -> thread 0 (waits for the weights,) issues + commits -> all wait
#define SAT_TC_ROUND(PRE, ISSUE)                             \
    do {                                                     \
        tc_wait_st();                                        \
        tc_fence_before();                                   \
        __syncthreads();                                     \
        if (tid == 0) {                                      \
            PRE;                                             \
            tc_fence_after();                                \
            ISSUE;                                           \
            mma_commit(mbar);                                \
        }                                                    \
        mbar_wait(mbar, phase);                              \
        phase ^= 1u;                                         \
        tc_fence_after();                                    \
    } while (0)

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int g = td.x, start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;

        float x[D];
        zero(x);
        if (valid) {
            long long id = load_id(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.sparse_cols + f));
            const long long lo = __ldg(m.row_off + f), hi = __ldg(m.row_off + f + 1);
            if (id < 0 || id >= hi - lo) { atomicOr(&a.counters[CNT_ERR], 1); id = 0; }
            load_row(table_row(m, lo + id), x);
        }

        const bool last_tile = t + (int)gridDim.x >= n_tiles;
        for (int l = 0; l < L; ++l) {
            __syncthreads();                      // previous users of sK / sV / sLN are done (every MMA round has been waited for)
            copy_to_smem(sLN, m.dense + m.off_layers + (long long)l * LAYER_FLOATS + OFF_LN_G, 192, tid, TC_NT);
            const int kq = theta_key(g, l, 0, L, pos), kk = theta_key(g, l, 1, L, pos);
            const bool has_next = !(last_tile && l == L - 1);         // another (tile, layer) follows in this CTA
            const int l_next = l + 1 < L ? l + 1 : 0;
            // K side first: its generated weights go into sBT now, the Q side's (ENP) are swapped in later.  A key block is only
            // fetched for a side that has a MetaNet, so that every fetch is waited for (by that side's W1 round) before the next
            if (meta_k && kk != cur_key) { if (tid == 0) fetch_key(kk); cur_key = kk; }
            float* act_l = TRAIN ? a.acts + (long long)l * ACT_PER_LAYER * a.act_stride : nullptr;
            if (TRAIN && valid) store_row(act_l + ACT_X * a.act_stride + tokpos, x);

            // ---- round 1: [Q | K | V] = x @ [W_Q | W_K | W_V] ---------------------------------------------------
            tmem_st_operand<PASSES>(tlane + TC_A, x);
            SAT_TC_ROUND(wait_block(1, TS_QKV),
                         issue_gemm_k32<PASSES>(tbase + TC_D, tbase + TC_A, sBL + PK_QKV_HI, sBL + PK_QKV_LO, 96, false));
            if (tid == 0 && has_next) fetch_qkv(l_next);           // the QKV block is free again: prefetch the next layer's
            float q[D], k[D], v[D];
            tmem_ld32(tlane + TC_D, q);
            tmem_ld32(tlane + TC_D + 32, k);
            tmem_ld32(tlane + TC_D + 64, v);
            tc_wait_ld();
            if (TRAIN && valid) store_row(act_l + ACT_V * a.act_stride + tokpos, v);
            sts_row(sV + tid * KV_LD, v);

            // ---- MetaNet on K then Q: z = dropout(relu(u @ W1) @ W2) + u ; LayerNorm ------------------------------
#pragma unroll
            for (int side = 1; side >= 0; --side) {
                float (&u)[D] = side ? k : q;
                const bool has_meta = side ? meta_k : meta_q;
                if (side == 0 && meta_q && kq != cur_key) {       // ENP (or no K-side MetaNet): the Q side's generated weights
                    if (tid == 0) fetch_key(kq);                  // (every MMA that read sBT has completed)
                    cur_key = kq;
                }
                if (has_meta) {
                    tmem_st_operand<PASSES>(tlane + TC_A, u);
                    SAT_TC_ROUND(wait_key(),
                                 issue_gemm_k32<PASSES>(tbase + TC_D, tbase + TC_A, sBT + PK_W1_HI, sBT + PK_W1_LO, 64, false));
                    float hbuf[D];
                    uint32_t m0 = 0u, m1 = 0u;
                    tmem_ld32(tlane + TC_D, hbuf);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < D; ++i) { if (hbuf[i] > 0.f) m0 |= 1u << i; hbuf[i] = fmaxf(hbuf[i], 0.f); }
                    tmem_st_operand<PASSES>(tlane + TC_H, hbuf);
                    tmem_ld32(tlane + TC_D + 32, hbuf);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < D; ++i) { if (hbuf[i] > 0.f) m1 |= 1u << i; hbuf[i] = fmaxf(hbuf[i], 0.f); }
                    if (TRAIN && valid)
                        *reinterpret_cast<uint2*>(relu_mask_ptr(a.acts, a.act_stride, L, l, side, tokpos / D)) = make_uint2(m0, m1);
                    // the hidden accumulator columns are free once both halves are in registers: the second half's operand
                    // goes there, and both W2 halves run in ONE round
                    tmem_st_operand<PASSES>(tlane + TC_D, hbuf);
                    SAT_TC_ROUND((void)0,
                                 (issue_gemm_k32<PASSES>(tbase + TC_D + 64, tbase + TC_H, sBT + PK_W2A_HI, sBT + PK_W2A_LO, 32, false),
                                  issue_gemm_k32<PASSES>(tbase + TC_D + 64, tbase + TC_D, sBT + PK_W2B_HI, sBT + PK_W2B_LO, 32, true)));
                    tmem_ld32(tlane + TC_D + 64, hbuf);
                    tc_wait_ld();
                    const DropCtx dc = make_drop(seed32, token_gid, l, side ? SITE_K : SITE_Q, a.p_drop, TRAIN);
#pragma unroll
                    for (int i = 0; i < D; i += 2) {
                        float f0, f1;
                        drop_pair(dc, i >> 1, f0, f1);
                        u[i] = fmaf(hbuf[i], f0, u[i]); u[i + 1] = fmaf(hbuf[i + 1], f1, u[i + 1]);
                    }
                }
                if (TRAIN && valid) store_row(act_l + (side ? ACT_ZK : ACT_ZQ) * a.act_stride + tokpos, u);
                if (has_meta) {
                    const float* ln = (side && pos) ? sLN + 128 : sLN + 64;
                    float mean, rstd;
                    layernorm32(u, ln, ln + 32, mean, rstd);
                }
                if (side) sts_row(sK + tid * KV_LD, k);
            }
            __syncthreads();                      // K', V tiles complete

            // ---- attention on the CUDA cores, then the output projection on the tensor core ------------------------
            float r[D];
            zero(r);
            if (valid)
                attention_fwd<H>(q, sK, sV, s * F, F, make_drop(seed32, token_gid, l, SITE_ATT, a.p_drop, TRAIN), r);
            if (TRAIN && valid) store_row(act_l + ACT_R * a.act_stride + tokpos, r);
            tmem_st_operand<PASSES>(tlane + TC_A, r);
            SAT_TC_ROUND(wait_block(2, TS_WO),
                         issue_gemm_k32<PASSES>(tbase + TC_D, tbase + TC_A, sBL + PK_WO_HI, sBL + PK_WO_LO, 32, false));
            if (tid == 0 && has_next) fetch_wo(l_next);
            float o[D];
            tmem_ld32(tlane + TC_D, o);
            tc_wait_ld();
            const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, TRAIN);
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                float f[2];
                drop_pair(dco, i >> 1, f[0], f[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float vv = relu_out ? fmaxf(o[i + e], 0.f) : o[i + e];
                    vv *= f[e];
                    x[i + e] = use_res ? vv + x[i + e] : vv;
                }
            }
            if (TRAIN && valid) store_row(act_l + ACT_Y * a.act_stride + tokpos, x);      // y (pre-LayerNorm): k_bwd_attn starts from it
            float mean, rstd;
            layernorm32(x, sLN, sLN + 32, mean, rstd);
        }

        // ---- logit head + sigmoid (+ BCE): identical to forward.cu ------------------------------------------------
        float part = 0.f;
        if (valid) {
            if (TRAIN && !head_bwd) store_row(a.acts + (long long)L * ACT_PER_LAYER * a.act_stride + tokpos, x);
            const float* hw = m.dense + m.off_headw + f * D;
#pragma unroll
            for (int i = 0; i < D; i += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(hw + i));
                part = fmaf(x[i], w4.x, part); part = fmaf(x[i + 1], w4.y, part);
                part = fmaf(x[i + 2], w4.z, part); part = fmaf(x[i + 3], w4.w, part);
            }
        }
        sRed[tid] = part;
        __syncthreads();
        if (valid && f == 0) {
            float z = 0.f;
            for (int j = 0; j < F; ++j) z += sRed[tid + j];
            for (int kd = 0; kd < m.n_dense; ++kd)
                z = fmaf(load_dense(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.dense_cols + kd)),
                         __ldg(m.dense + m.off_headw + F * D + kd), z);
            z += __ldg(m.dense + m.off_headb);
            const float yp = 1.f / (1.f + expf(-z));
            a.y_pred[b] = yp;
            double lossv = 0.0;
            float dzval = 0.f;
            if (a.y) {
                const float yt = __ldg(a.y + b);
                const float pq = (1.f - yp) * yp;            // d sigmoid / d logit
                if (m.flags & SAT_FLAG_LOSS_MSE) {           // F.mse_loss(reduction='sum')
                    const float d = yp - yt;
                    lossv = (double)(d * d);
                    dzval = 2.f * d * pq;
                } else if (m.flags & SAT_FLAG_LOSS_MAE) {    // F.l1_loss(reduction='sum')
                    const float d = yp - yt;
                    lossv = (double)fabsf(d);
                    dzval = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * pq;
                } else {
                    // F.binary_cross_entropy(reduction='sum'): logs clamped at -100; autograd of BCE then sigmoid:
                    // (p - y) / max((1-p) p, 1e-12) * (1-p) p
                    const float lp = fmaxf(logf(yp), -100.f), l1p = fmaxf(logf(1.f - yp), -100.f);
                    lossv = (double)(-(yt * lp + (1.f - yt) * l1p));
                    dzval = (yp - yt) / fmaxf(pq, 1e-12f) * pq;
                }
                a.dz[p] = dzval;
            }
            sDz[s] = dzval;
            s_lossv[s] = lossv;
        }
        __syncthreads();
        if (tid == 0 && a.loss_part) {
            double acc = 0.0;
            for (int ss = 0; ss < cnt; ++ss) acc += s_lossv[ss];
            a.loss_part[t] = acc;
        }
        if (head_bwd) {
            // ---- backward of the logit head (reference satrans.py:244-255 under autograd), folded into the forward's epilogue:
            //      dX_L[tok] = dz[b] * w_head[f];  d w_head[f] += dz[b] * x_L[tok];  d bias += dz[b]  (sums over the CTA's tiles in
            //      a fixed order; the CTA's totals go to its row of the head partial buffer, k_dense_reduce adds the rows)
            const float dzv = valid ? sDz[s] : 0.f;
            float xl[D];
            if (valid) {
                const float* hw = m.dense + m.off_headw + f * D;
                float dout[D];
#pragma unroll
                for (int i = 0; i < D; i += 4) {
                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(hw + i));
                    dout[i] = dzv * w4.x; dout[i + 1] = dzv * w4.y; dout[i + 2] = dzv * w4.z; dout[i + 3] = dzv * w4.w;
                }
                store_row(a.dx + tokpos, dout);
            }
#pragma unroll
            for (int i = 0; i < D; ++i) xl[i] = valid ? x[i] * dzv : 0.f;
            sts_row(sK + tid * KV_LD, xl);                    // (the K' tile is free: the last layer's attention is done)
            __syncthreads();
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int ff = (tid >> 3) + pass * (TC_NT / 8), ch = (tid & 7) * 4;
                if (ff < F) {
                    float4 acc = *reinterpret_cast<const float4*>(sHead + ff * D + ch);
                    for (int ss = 0; ss < cnt; ++ss) {
                        const float4 v = *reinterpret_cast<const float4*>(sK + (ss * F + ff) * KV_LD + ch);
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    }
                    *reinterpret_cast<float4*>(sHead + ff * D + ch) = acc;
                }
            }
            if (tid == 0) {
                float accb = sHead[F * D + m.n_dense];
                for (int ss = 0; ss < cnt; ++ss) accb += sDz[ss];
                sHead[F * D + m.n_dense] = accb;
            }
            if (tid >= 32 && tid < 32 + m.n_dense) {
                const int kd = tid - 32;
                float accd = sHead[F * D + kd];
                for (int ss = 0; ss < cnt; ++ss)
                    accd = fmaf(sDz[ss], load_dense(a.x, a.id_kind, (long long)a.perm[start + ss] * a.ld + __ldg(m.dense_cols + kd)), accd);
                sHead[F * D + kd] = accd;
            }
        }
    }
#undef SAT_TC_ROUND
    tc_fence_before();
    __syncthreads();
    if (head_bwd) {
        const GPart gp = gpart_layout(m);
        float* part = a.gpart + gp.head + (long long)blockIdx.x * gp.head_ld;
        for (int i = tid; i < F * D + m.n_dense + 1; i += TC_NT) part[i] = sHead[i];
    }
    if (warp == 0) tmem_dealloc(tbase, TC_COLS);
}

}  // namespace sat

using namespace sat;

#ifndef SAT_INST_H
#error "compile forward_tc.cu with -DSAT_INST_H={1,2,4,8}"
#endif
#define SAT_CAT_(a, b) a##b
#define SAT_CAT(a, b) SAT_CAT_(a, b)

// 2-D tensor map over the packed weights: [blocks * 32 rows][256 floats] fp32, box = 8 rows x 256 floats (8 KB), no swizzle
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_wpack_tmap(const float* wpack, long long n_blocks, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return sat_host::fail(SAT_ECUDA, "sat_forward_tc: cuTensorMapEncodeTiled is not available from the driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)TMA_ROW_FLOATS, (cuuint64_t)(n_blocks * PACK_ROWS)};
    const cuuint64_t gstride[1] = {(cuuint64_t)TMA_ROW_FLOATS * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TMA_ROW_FLOATS, (cuuint32_t)TMA_BOX_ROWS};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wpack), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return sat_host::fail(SAT_ECUDA, "sat_forward_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SAT_OK;
}

template <bool TRAIN, int PASSES>
static int launch_tc(const FwdArgs& a, const CUtensorMap& tmap, int grid, cudaStream_t s) {
    constexpr int H = SAT_INST_H;
    const size_t smem = TC_SMEM_FLOATS * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_forward_tc<H, TRAIN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_forward_tc: cannot raise dynamic shared memory to %zu", smem);
        configured = true;
    }
    k_forward_tc<H, TRAIN, PASSES><<<grid, TC_NT, smem, s>>>(a, tmap);
    return sat_host::check_launch("k_forward_tc");
}

int SAT_CAT(sat_launch_forward_tc_h, SAT_INST_H)(const FwdArgs& a, const float* wpack, int grid, int training, int passes,
                                                 cudaStream_t s) {
    const int n_keys = (a.m.flags & SAT_FLAG_POS) ? a.m.n_dom * a.m.L * 2 : a.m.n_dom;
    CUtensorMap tmap;
    const int rc = make_wpack_tmap(wpack, 2ll * (a.m.L + n_keys), &tmap);
    if (rc) return rc;
    if (passes == 3) return training ? launch_tc<true, 3>(a, tmap, grid, s) : launch_tc<false, 3>(a, tmap, grid, s);
    if (passes == 1) return training ? launch_tc<true, 1>(a, tmap, grid, s) : launch_tc<false, 1>(a, tmap, grid, s);
    if (passes == 2) return training ? launch_tc<true, 2>(a, tmap, grid, s) : launch_tc<false, 2>(a, tmap, grid, s);
    return sat_host::fail(SAT_EINVAL, "sat_forward_tc: passes must be 1 (tf32), 2 (bf16) or 3 (3xtf32), got %d", passes);
}

#if SAT_INST_H == 4
// one copy of the packing entry point (any H translation unit would do)
int sat_launch_pack_weights(const SatModel* m, int n_keys, const float* theta, float* wpack, int bf16, cudaStream_t s) {
    const int total = 2 * (m->L + n_keys) * 4096;
    k_pack_weights<<<(total + 255) / 256, 256, 0, s>>>(*m, n_keys, theta, wpack, bf16);
    return sat_host::check_launch("k_pack_weights");
}
#endif
