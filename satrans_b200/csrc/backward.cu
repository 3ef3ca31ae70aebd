// Fused SATrans backward (replaces torch autograd over the graph of reference satrans.py:197-256).
//
// Same mapping as the forward: one thread = one field token, a CTA = whole samples of one scenario.
// Per layer, two kernels:
//   k_bwd_attn : LayerNorm-out bwd, out-projection bwd (+dW_out), attention bwd -> dQ', dK', dV
//   k_bwd_meta : MetaNet-LayerNorm bwd, MetaNet bwd (+d theta), QKV projection bwd (+dW_q/k/v) -> dX
// Data gradients are thread-local GEMVs against weights in shared memory (as in the forward).  Weight
// gradients are reductions over tokens: the two operands are staged as [token][36] shared tiles and every
// thread owns a 4x4 block of the 32x32 product (64 blocks x 3 token groups = 192 threads), accumulated
// into shared-memory accumulators.  No floating-point atomics: every CTA stores its accumulators into its own
// row of SatWork.gpart and k_dense_reduce adds the rows in CTA order, so the gradients are bit-reproducible.
// Attention uses the FlashAttention identity  sum_j P_ij dP_ij = dR_i . R_i  so no (F x F) tensor is
// stored; probabilities are recomputed from Q'/K' for the query-role pass and again for the key-role pass.
#include <stdlib.h>

#include "bwd_common.cuh"

namespace sat {


// =====================================================================================================
// head backward: dOut_L[tok][c] = dz[b] * w_head[f*32+c];  dW_head, db_head.   (reference satrans.py:244-255)
// =====================================================================================================
__global__ void __launch_bounds__(NT) k_bwd_head(const BwdArgs a) {
    __shared__ __align__(16) float sT[NT * KV_LD];
    __shared__ float sDz[NT];
    const SatModel& m = a.m;
    const int tid = threadIdx.x, F = m.F, L = m.L;
    const int n_tiles = a.counters[CNT_NTILES];
    const int s = tid / F, f = tid - s * F;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    // thread (ff, chunk) owns dW_head[ff*32 + chunk*4 .. +4) for ff = tid/8 (+24 per pass)
    float4 accw[2];
    accw[0] = make_float4(0.f, 0.f, 0.f, 0.f); accw[1] = accw[0];
    float accb = 0.f, accd = 0.f;      // bias grad; dense-feature weight grad (thread k < n_dense)
    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int p = start + s;
        const long long tokpos = ((long long)p * F + f) * D;
        __syncthreads();
        if (valid) {
            const float dzv = a.dz[p];
            float xl[D], dout[D];
            load_row(a.acts + (long long)L * ACT_PER_LAYER * a.act_stride + tokpos, xl);
            const float* hw = m.dense + m.off_headw + f * D;
#pragma unroll
            for (int i = 0; i < D; ++i) { dout[i] = dzv * __ldg(hw + i); xl[i] *= dzv; }
            store_row(a.dx + tokpos, dout);
            sts_row(sT + tid * KV_LD, xl);
            if (f == 0) sDz[s] = dzv;
        }
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
            const int ff = (tid >> 3) + pass * (NT / 8), ch = (tid & 7) * 4;
            if (ff < F) {
                for (int ss = 0; ss < cnt; ++ss) {
                    const float4 v = *reinterpret_cast<const float4*>(sT + (ss * F + ff) * KV_LD + ch);
                    accw[pass].x += v.x; accw[pass].y += v.y; accw[pass].z += v.z; accw[pass].w += v.w;
                }
            }
        }
        if (tid == 0) for (int ss = 0; ss < cnt; ++ss) accb += sDz[ss];
        if (tid >= 32 && tid < 32 + m.n_dense) {
            const int k = tid - 32;
            for (int ss = 0; ss < cnt; ++ss) {
                const int b = a.perm[start + ss];
                accd = fmaf(sDz[ss], load_dense(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.dense_cols + k)), accd);
            }
        }
    }
    // this CTA's row of the head partial buffer: [F*32 weights | n_dense weights | bias]
    const GPart gp = gpart_layout(m);
    float* part = a.gpart + gp.head + (long long)blockIdx.x * gp.head_ld;
    for (int pass = 0; pass < 2; ++pass) {
        const int ff = (tid >> 3) + pass * (NT / 8), ch = (tid & 7) * 4;
        if (ff < F) *reinterpret_cast<float4*>(part + ff * D + ch) = accw[pass];
    }
    if (tid == 0) part[F * D + m.n_dense] = accb;
    if (tid >= 32 && tid < 32 + m.n_dense) part[F * D + (tid - 32)] = accd;
}

// =====================================================================================================
// kernel B: MetaNet LayerNorm bwd + MetaNet bwd + QKV projection bwd; layer 0 scatters dE0 to table rows.
// =====================================================================================================
constexpr int BWDB_SMEM_FLOATS = 3072 /*WQ WK WV*/ + 3072 /*dWQ dWK dWV*/ + 192 /*LN params*/ + (NT / 32) * 128 /*dLN q,k per warp*/ +
                                 2 * PSZ /*theta q,k*/ + 2 * PSZ /*dtheta q,k*/ + 2 * NT * KV_LD;

// One MetaNet side (Q or K).  On entry dqp holds dLoss/dQ' (post-LN); on exit dLoss/d(x @ W) .
// All threads of the CTA must call this (it contains __syncthreads()).
__device__ __forceinline__ void metanet_side_bwd(const BwdArgs& a, bool has_meta, bool valid, int ntok, int tid, int lane,
                                                 const float* __restrict__ sWp,      // projection weight [32][32]
                                                 const float* __restrict__ sTh,      // W1|W2
                                                 float* __restrict__ sdTh,           // their gradient accumulators
                                                 const float* __restrict__ lng,      // LN gamma (smem)
                                                 float* __restrict__ sdLNg,          // dgamma[32] | dbeta[32]
                                                 float* __restrict__ sdWp,           // projection weight gradient
                                                 float* __restrict__ sA, float* __restrict__ sG,
                                                 const float* __restrict__ act_x, const float* __restrict__ act_z,
                                                 const uint32_t* __restrict__ relu_mask,   // 2 words of this token/side
                                                 long long tokpos, const DropCtx& dc, float (&dqp)[D]) {
    float qp[D];
    zero(qp);
    if (has_meta) {
        float zh[D];
        float rstd = 0.f;
        if (valid) { load_row(act_z + tokpos, zh); rstd = normalize32(zh); } else zero(zh);
        layernorm_bwd32(dqp, zh, rstd, lng, sdLNg, sdLNg + 32, lane, valid);        // dqp := dLoss/dz
        float dm[D];
        if (valid) {
            float xin[D];
            load_row(act_x + tokpos, xin);
            vm_outer<D, D, D>(sWp, xin, qp);                                         // recompute qp = x @ W
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                float f0, f1;
                drop_pair(dc, i >> 1, f0, f1);
                dm[i] = dqp[i] * f0; dm[i + 1] = dqp[i + 1] * f1;
            }
        } else zero(dm);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            // the forward's ReLU decisions (saved), not a re-evaluation: identical function even if the forward
            // ran on the tensor cores
            const uint32_t mask = valid ? __ldg(relu_mask + half) : 0u;
            {
                float h[D];
                zero(h);
                if (valid) {
#pragma unroll
                    for (int c0 = 0; c0 < D; c0 += 16) {
                        float hc[16];
                        zero(hc);
                        vm_outer<D, 16, HID>(sTh + half * D + c0, qp, hc);
#pragma unroll
                        for (int c = 0; c < 16; ++c) h[c0 + c] = ((mask >> (c0 + c)) & 1u) ? hc[c] : 0.f;
                    }
                    sts_row(sA + tid * KV_LD, h);
                    sts_row(sG + tid * KV_LD, dm);
                }
            }
            __syncthreads();
            wgrad32(sA, sG, ntok, sdTh + D * HID + half * D * D, D, tid);       // dW2[half*32 + c][j] += h[c] dm[j]
            __syncthreads();
            float dh[D];
            zero(dh);
            if (valid) {
#pragma unroll
                for (int c0 = 0; c0 < D; c0 += 16) {
                    float dhc[16];
                    zero(dhc);
                    vm_dot<D, 16, D>(sTh + D * HID + (half * D + c0) * D, dm, dhc);   // dh[c] = sum_j dm[j] W2[c][j]
#pragma unroll
                    for (int c = 0; c < 16; ++c) dh[c0 + c] = ((mask >> (c0 + c)) & 1u) ? dhc[c] : 0.f;
                }
                sts_row(sA + tid * KV_LD, qp);
                sts_row(sG + tid * KV_LD, dh);
            }
            __syncthreads();
            wgrad32(sA, sG, ntok, sdTh + half * D, HID, tid);                     // dW1[i][half*32 + c] += qp[i] dh[c]
            __syncthreads();
            if (valid) vm_dot<D, D, HID>(sTh + half * D, dh, dqp);                 // dqp[i] += sum_c dh[c] W1[i][half*32+c]
        }
    }
    // projection weight gradient: dW[i][j] += x[i] dqp[j]
    if (valid) {
        float xin[D];
        load_row(act_x + tokpos, xin);
        sts_row(sA + tid * KV_LD, xin);
        sts_row(sG + tid * KV_LD, dqp);
    }
    __syncthreads();
    wgrad32(sA, sG, ntok, sdWp, D, tid);
    __syncthreads();
}

// Backward of the ablation variants of one side (reference satrans.py:61-64,68-69,79-81 under autograd; forward.cu):
//   bilinear (Q side, heads = H > 0):  Q'_h = qp_h @ B_h  ->  dB_h += qp_h^T dQ'_h  (the full 32x32 product qp^T dQ' is accumulated
//                                      at sdTh[1024 + i*32 + j]; k_bwd_meta keeps its diagonal blocks),  dqp_h = dQ'_h @ B_h^T
//   gate:                              Q' = qp * g * 2       ->  dg += sum_tokens 2 qp dQ' (tokens added in order), dqp = dQ' * 2 g
// with qp = x @ W recomputed.  On entry dqp = dLoss/dQ', on exit dLoss/d(x @ W).  All threads of the CTA must call this.
template <int H>
__device__ __forceinline__ void bilinear_dq(const float* __restrict__ sTh, float (&dqp)[D]) {
    constexpr int DH = D / H;
    float out[D];
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int i = 0; i < DH; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < DH; ++j) acc = fmaf(dqp[h * DH + j], sTh[(h * DH + i) * DH + j], acc);
            out[h * DH + i] = acc;
        }
#pragma unroll
    for (int i = 0; i < D; ++i) dqp[i] = out[i];
}
__device__ __forceinline__ void variant_side_bwd(bool gate, int heads, bool valid, int ntok, int tid,
                                                 const float* __restrict__ sWp, const float* __restrict__ sTh,
                                                 float* __restrict__ sdTh, float* __restrict__ sA, float* __restrict__ sG,
                                                 const float* __restrict__ act_x, long long tokpos, float (&dqp)[D]) {
    float qp[D];
    zero(qp);
    if (valid) {
        float xin[D];
        load_row(act_x + tokpos, xin);
        vm_outer<D, D, D>(sWp, xin, qp);
    }
    if (heads > 0) {
        if (valid) { sts_row(sA + tid * KV_LD, qp); sts_row(sG + tid * KV_LD, dqp); }
        __syncthreads();
        wgrad32(sA, sG, ntok, sdTh + 1024, D, tid);
        __syncthreads();
        if (valid) {
            switch (heads) {
                case 1: bilinear_dq<1>(sTh, dqp); break;
                case 2: bilinear_dq<2>(sTh, dqp); break;
                case 4: bilinear_dq<4>(sTh, dqp); break;
                default: bilinear_dq<8>(sTh, dqp); break;
            }
        }
    }
    if (gate) {
        if (valid) {
            float pr[D];
#pragma unroll
            for (int i = 0; i < D; ++i) { pr[i] = 2.f * qp[i] * dqp[i]; dqp[i] = dqp[i] * (2.f * sTh[i]); }
            sts_row(sA + tid * KV_LD, pr);
        }
        __syncthreads();
        if (tid < D) {
            float acc = 0.f;
            for (int t = 0; t < ntok; ++t) acc += sA[t * KV_LD + tid];
            sdTh[tid] += acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(NT, 1) k_bwd_meta(const BwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                    // WQ WK WV
    float* sdW = sW + 3072;
    float* sLN = sdW + 3072;             // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;             // per warp: dQLN_G dQLN_B dKLN_G dKLN_B
    float* sThQ = sdLN + (NT / 32) * 128;
    float* sThK = sThQ + PSZ;
    float* sdThQ = sThK + PSZ;
    float* sdThK = sdThQ + PSZ;
    float* sA = sdThK + PSZ;
    float* sG = sA + NT * KV_LD;

    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, lane = tid & 31, F = m.F, L = m.L, l = a.layer;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool gate_q = m.flags & SAT_FLAG_GATE_Q, gate_k = m.flags & SAT_FLAG_GATE_K;
    const int bil_heads = (m.flags & SAT_FLAG_BILINEAR) ? m.H : 0;
    const int s = tid / F, f = tid - s * F;
    // every CTA stays inside ONE scenario (one set of generated-weight gradients per CTA, see scenario_partition)
    int t0, t1, g_cta, c0_, c1_;
    scenario_partition(a.counters + CNT_HIST, m.n_dom, NT / F, gridDim.x, blockIdx.x, -1, t0, t1, g_cta, c0_, c1_);
    const float* lay = m.dense + m.off_layers + (long long)l * LAYER_FLOATS;
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;

    copy_to_smem(sW, lay, 3072, tid, NT);
    copy_to_smem(sLN, lay + OFF_LN_G, 192, tid, NT);
    for (int i = tid; i < 3072; i += NT) sdW[i] = 0.f;
    for (int i = tid; i < 2 * PSZ; i += NT) sdThQ[i] = 0.f;      // sdThQ and sdThK are contiguous
    for (int i = tid; i < (NT / 32) * 128; i += NT) sdLN[i] = 0.f;
    int cur_q = -1, cur_k = -1;
    float* sdLNq = sdLN + (tid >> 5) * 128;                    // this warp's slots
    float* sdLNk = pos ? sdLNq + 64 : sdLNq;                   // without 'pos' K shares Q's MetaNet LayerNorm
    const float* klng = pos ? sLN + 128 : sLN + 64;
    float* sdThKuse = pos ? sdThK : sdThQ;                     // ... and Q's generated weights

    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int g = td.x, start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int ntok = cnt * F;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        const int kq = theta_key(g, l, 0, L, pos), kk = theta_key(g, l, 1, L, pos);
        __syncthreads();
        if (kq != cur_q) { copy_to_smem(sThQ, a.theta + (long long)kq * PSZ, PSZ, tid, NT); cur_q = kq; }       // (once per CTA)
        if (pos && kk != cur_k) { copy_to_smem(sThK, a.theta + (long long)kk * PSZ, PSZ, tid, NT); cur_k = kk; }
        __syncthreads();
        const float* thK = pos ? sThK : sThQ;

        float dxa[D];                       // accumulates dLoss/dX of this token
        if (valid) load_row(a.dx + tokpos, dxa); else zero(dxa);
        {   // ---- Q side ---------------------------------------------------------------------------------
            float dqp[D];
            if (valid) load_row(a.dqkv + 0 * a.act_stride + tokpos, dqp); else zero(dqp);
            if (gate_q || bil_heads)
                variant_side_bwd(gate_q, bil_heads, valid, ntok, tid, sW + OFF_WQ, sThQ, sdThQ, sA, sG, act_l + ACT_X * a.act_stride,
                                 tokpos, dqp);
            metanet_side_bwd(a, meta_q, valid, ntok, tid, lane, sW + OFF_WQ, sThQ, sdThQ, sLN + 64, sdLNq, sdW + OFF_WQ, sA,
                             sG, act_l + ACT_X * a.act_stride, act_l + ACT_ZQ * a.act_stride,
                             relu_mask_ptr(const_cast<float*>(a.acts), a.act_stride, L, l, 0, tokpos / D), tokpos,
                             make_drop(seed32, token_gid, l, SITE_Q, a.p_drop, true), dqp);
            if (valid) vm_dot<D, D, D>(sW + OFF_WQ, dqp, dxa);          // dX[i] += sum_j dqp[j] WQ[i][j]
        }
        {   // ---- K side ---------------------------------------------------------------------------------
            float dkp[D];
            if (valid) load_row(a.dqkv + 1 * a.act_stride + tokpos, dkp); else zero(dkp);
            if (gate_k)
                variant_side_bwd(true, 0, valid, ntok, tid, sW + OFF_WK, thK, sdThKuse, sA, sG, act_l + ACT_X * a.act_stride, tokpos, dkp);
            metanet_side_bwd(a, meta_k, valid, ntok, tid, lane, sW + OFF_WK, thK, sdThKuse, klng, sdLNk, sdW + OFF_WK, sA,
                             sG, act_l + ACT_X * a.act_stride, act_l + ACT_ZK * a.act_stride,
                             relu_mask_ptr(const_cast<float*>(a.acts), a.act_stride, L, l, 1, tokpos / D), tokpos,
                             make_drop(seed32, token_gid, l, SITE_K, a.p_drop, true), dkp);
            if (valid) vm_dot<D, D, D>(sW + OFF_WK, dkp, dxa);
        }
        {   // ---- V side ---------------------------------------------------------------------------------
            float dv[D];
            if (valid) {
                float xin[D];
                load_row(a.dqkv + 2 * a.act_stride + tokpos, dv);
                load_row(act_l + ACT_X * a.act_stride + tokpos, xin);
                sts_row(sA + tid * KV_LD, xin);
                sts_row(sG + tid * KV_LD, dv);
            }
            __syncthreads();
            wgrad32(sA, sG, ntok, sdW + OFF_WV, D, tid);
            __syncthreads();
            if (valid) vm_dot<D, D, D>(sW + OFF_WV, dv, dxa);
        }
        // layer 0: this is dLoss/dE0 of the token; sat_rows_reduce sums it per touched table row in sorted (row, token) order
        if (valid) store_row(a.dx + tokpos, dxa);
    }
    __syncthreads();
    if (bil_heads) {        // d B_h = the h-th diagonal dh x dh block of qp^T dQ' (accumulated at sdThQ[1024 + i*32 + j])
        const int dh = D / bil_heads;
        for (int e = tid; e < D * dh; e += NT) {
            const int row = e / dh, j = e - row * dh;                     // row = h*dh + i
            sdThQ[e] = sdThQ[1024 + row * D + (row / dh) * dh + j];
        }
        __syncthreads();
    }
    const GPart gp = gpart_layout(m);
    float* part = a.gpart + gp.meta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * GP_META;
    store_partial(part, sdW, 3072, tid);
    store_partial_warps(part + 3072, sdLN, 128, tid);
    float* tpart = a.gpart + gp.theta + ((long long)l * SAT_GRID_MAX + blockIdx.x) * gp.nside * PSZ;
    store_partial(tpart, sdThQ, gp.nside * PSZ, tid);          // sdThQ | sdThK are contiguous
}

// =====================================================================================================
// Fixed-order reduction of the per-CTA partial buffers into the flat dense gradient and d theta.
// One thread per output element x RED_SPLIT CTA groups; the groups' sums are combined in group order.
// =====================================================================================================
constexpr int RED_SPLIT = 4;
struct RedArgs {
    SatModel m;
    const float* gpart;
    const int32_t* counters;
    float* dgrad;
    float* dtheta;
    int g_attn, g_meta, g_head;   // CTAs that wrote partials
    int spt_meta;                 // samples per tile of the tile set the MetaNet backward walked
};

__device__ __forceinline__ float sum_rows(const float* __restrict__ base, long long stride, int c0, int c1, int grp) {
    // rows [c0, c1) split into RED_SPLIT contiguous groups; this thread adds group `grp` in row order
    const int n = c1 - c0, per = (n + RED_SPLIT - 1) / RED_SPLIT;
    const int a0 = c0 + grp * per, a1 = min(c1, a0 + per);
    float acc = 0.f;
    int c = a0;
    for (; c + 4 <= a1; c += 4) {
        const float v0 = base[(long long)c * stride], v1 = base[(long long)(c + 1) * stride];
        const float v2 = base[(long long)(c + 2) * stride], v3 = base[(long long)(c + 3) * stride];
        acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; c < a1; ++c) acc += base[(long long)c * stride];
    return acc;
}

__global__ void __launch_bounds__(256) k_dense_reduce(const RedArgs a) {
    __shared__ float s_part[RED_SPLIT][64];
    const SatModel& m = a.m;
    const GPart gp = gpart_layout(m);
    const int L = m.L;
    const bool pos = m.flags & SAT_FLAG_POS;
    const int n_keys = pos ? m.n_dom * L * 2 : m.n_dom;
    const long long n_lay = (long long)L * LAYER_FLOATS, n_head = m.F * D + m.n_dense + 1, n_theta = (long long)n_keys * PSZ;
    const long long total = n_lay + n_head + n_theta;
    const int el = threadIdx.x & 63, grp = threadIdx.x >> 6;
    for (long long e0 = (long long)blockIdx.x * 64; e0 < total; e0 += (long long)gridDim.x * 64) {
        const long long e = e0 + el;
        float v = 0.f;
        float* out = nullptr;
        if (e < n_lay) {
            const int l = (int)(e / LAYER_FLOATS), i = (int)(e - (long long)l * LAYER_FLOATS);
            out = a.dgrad + m.off_layers + e;
            if (i < OFF_WO)             // dW_Q, dW_K, dW_V
                v = sum_rows(a.gpart + gp.meta + (long long)l * SAT_GRID_MAX * GP_META + i, GP_META, 0, a.g_meta, grp);
            else if (i < OFF_QLN_G)     // dW_out, d layer_norm
                v = sum_rows(a.gpart + gp.attn + (long long)l * SAT_GRID_MAX * GP_ATTN + (i - OFF_WO), GP_ATTN, 0, a.g_attn, grp);
            else                        // d MetaNet LayerNorms
                v = sum_rows(a.gpart + gp.meta + (long long)l * SAT_GRID_MAX * GP_META + 3072 + (i - OFF_QLN_G), GP_META, 0, a.g_meta, grp);
        } else if (e < n_lay + n_head) {
            const int i = (int)(e - n_lay);
            out = a.dgrad + m.off_headw + i;            // off_headb == off_headw + F*32 + n_dense
            v = sum_rows(a.gpart + gp.head + i, gp.head_ld, 0, a.g_head, grp);
        } else if (e < total) {
            const long long j = e - n_lay - n_head;
            const int key = (int)(j / PSZ), i = (int)(j - (long long)key * PSZ);
            out = a.dtheta + j;
            int g, t0, t1, g_of, c0, c1;
            if (pos) {                  // key = (g * L + l) * 2 + side: one layer, one side
                const int side = key & 1, l = (key >> 1) % L;
                g = (key >> 1) / L;
                scenario_partition(a.counters + CNT_HIST, m.n_dom, a.spt_meta, a.g_meta, -1, g, t0, t1, g_of, c0, c1);
                v = sum_rows(a.gpart + gp.theta + ((long long)l * SAT_GRID_MAX * 2 + side) * PSZ + i, 2 * PSZ, c0, c1, grp);
            } else {                    // key = g: the Q and K sides of all L layers share theta; layers added last to first
                g = key;
                scenario_partition(a.counters + CNT_HIST, m.n_dom, a.spt_meta, a.g_meta, -1, g, t0, t1, g_of, c0, c1);
                for (int l = L - 1; l >= 0; --l)
                    v += sum_rows(a.gpart + gp.theta + (long long)l * SAT_GRID_MAX * PSZ + i, PSZ, c0, c1, grp);
            }
        }
        s_part[grp][el] = v;
        __syncthreads();
        if (grp == 0 && out) *out = (s_part[0][el] + s_part[1][el]) + (s_part[2][el] + s_part[3][el]);
        __syncthreads();
    }
}

}  // namespace sat

using namespace sat;

int sat_launch_bwd_meta_tc(const sat::BwdArgs& a, const float* wpack, int n_keys, int grid, int passes, cudaStream_t s);
int sat_launch_bwd_meta_tc_wg(const sat::BwdArgs& a, const float* wpack, int n_keys, int grid, int passes, cudaStream_t s);
int sat_launch_bwd_attn2_h2(const sat::BwdArgs& a, int grid, cudaStream_t s);
int sat_launch_bwd_attn2_h4(const sat::BwdArgs& a, int grid, cudaStream_t s);
int sat_launch_bwd_attn2_h8(const sat::BwdArgs& a, int grid, cudaStream_t s);
#ifndef SAT_BWD_ATTN_SPLIT_DEFAULT
#define SAT_BWD_ATTN_SPLIT_DEFAULT 0
#endif
// SAT_BWD_ATTN_SPLIT=0 / 1: attention backward with one / two threads per token (backward_attn.cu / backward_attn2.cu)
static bool attn_token_split() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SAT_BWD_ATTN_SPLIT"); v = e ? (atoi(e) != 0) : SAT_BWD_ATTN_SPLIT_DEFAULT; }
    return v == 1;
}
// SAT_WGRAD_TC=1: weight gradients on the tensor core too (backward_tc_wg.cu; measured slower, see there)
static bool wgrad_on_tensor_core() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SAT_WGRAD_TC"); v = (e && atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}
extern "C" int sat_backward_finish(const SatModel* m, SatWork* w, int32_t passes, void* stream);

// passes = 0: CUDA-core k_bwd_meta; 1 / 3: tensor-core k_bwd_meta_tc (TF32 / 3xTF32), needs sat_pack_weights first
static int backward_impl(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, float dropout_p,
                         uint64_t seed, int passes, void* stream) {
    if (!m || !x || !w || !w->acts || !w->dx || !w->dqkv || !w->dgrad || !w->dtheta || !w->gpart || !w->tiles2)
        return sat_host::fail(SAT_EINVAL, "sat_backward: null argument / workspace not sized for training");
    if (w->B == 0) return SAT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int n_keys = (m->flags & SAT_FLAG_POS) ? m->n_dom * m->L * 2 : m->n_dom;
    // (k_dense_reduce writes every layer / head / d theta element; the encoder backward accumulates into the rest)
    if (cudaMemsetAsync(w->dgrad, 0, (size_t)m->n_dense_params * sizeof(float), s) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_backward: memset failed");
    BwdArgs a;
    a.m = *m; a.x = x; a.id_kind = id_kind; a.ld = ld; a.perm = w->perm; a.tiles = w->tiles; a.counters = w->counters;
    a.theta = w->theta; a.dtheta = w->dtheta; a.acts = w->acts; a.act_stride = (long long)w->cap_B * m->F * D;
    a.dz = w->dz; a.dx = w->dx; a.dqkv = w->dqkv; a.dgrad = w->dgrad; a.bitmap = w->bitmap; a.word_rank = w->word_rank;
    a.row_grad = w->row_grad; a.layer = 0; a.p_drop = dropout_p; a.seed = seed; a.gpart = w->gpart;
    a.seed_ptr = w->device_step ? reinterpret_cast<const unsigned long long*>(w->step_state) : nullptr;
    const int spt = NT / m->F;
    const int max_tiles = (w->B + spt - 1) / spt + m->n_dom;
    const int grid2 = max_tiles < 148 * 2 ? max_tiles : 148 * 2;
    const int grid1 = max_tiles < 148 ? max_tiles : 148;
    static bool configured_b = false;
    const size_t smem_b = BWDB_SMEM_FLOATS * sizeof(float);
    if (!configured_b) {
        if (cudaFuncSetAttribute(k_bwd_meta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem_b);
        configured_b = true;
    }
    int rc = SAT_OK;
    const bool head_done = w->head_bwd_done != 0;      // sat_forward_tc already ran the logit head's backward in its epilogue
    if (!head_done) {
        k_bwd_head<<<grid2, NT, 0, s>>>(a);
        rc = sat_host::check_launch("k_bwd_head");
        if (rc) return rc;
    }
    const int spt2 = NT2 / m->F;
    const int max_tiles2 = (w->B + spt2 - 1) / spt2 + m->n_dom;
    BwdArgs aa = a;                           // k_bwd_attn walks the 128-token tile set
    aa.tiles = w->tiles2;
    const int grid_a = max_tiles2 < 148 * 2 ? max_tiles2 : 148 * 2;
    for (int l = m->L - 1; l >= 0; --l) {
        a.layer = l;
        aa.layer = l;
        const bool split = attn_token_split() && m->H >= 2;      // two threads per token (backward_attn2.cu)
        switch (m->H) {
            case 1: rc = sat_launch_bwd_attn_h1(aa, grid_a, s); break;
            case 2: rc = split ? sat_launch_bwd_attn2_h2(aa, grid_a, s) : sat_launch_bwd_attn_h2(aa, grid_a, s); break;
            case 4: rc = split ? sat_launch_bwd_attn2_h4(aa, grid_a, s) : sat_launch_bwd_attn_h4(aa, grid_a, s); break;
            case 8: rc = split ? sat_launch_bwd_attn2_h8(aa, grid_a, s) : sat_launch_bwd_attn_h8(aa, grid_a, s); break;
            default: return sat_host::fail(SAT_EINVAL, "sat_backward: heads=%d not in {1,2,4,8}", m->H);
        }
        if (rc) return rc;
        if (passes == 0) {
            k_bwd_meta<<<grid1, NT, smem_b, s>>>(a);
            rc = sat_host::check_launch("k_bwd_meta");
        } else {
            BwdArgs a2 = a;
            a2.tiles = w->tiles2;
            rc = wgrad_on_tensor_core() ? sat_launch_bwd_meta_tc_wg(a2, w->wpack, n_keys, grid_a, passes, s)
                                        : sat_launch_bwd_meta_tc(a2, w->wpack, n_keys, grid_a, passes, s);
        }
        if (rc) return rc;
    }
    if (w->defer_finish) return SAT_OK;            // the caller runs sat_backward_finish (e.g. beside the table sweep)
    return sat_backward_finish(m, w, passes, stream);
}

extern "C" int sat_backward_finish(const SatModel* m, SatWork* w, int32_t passes, void* stream) {
    if (!m || !w || !w->gpart || !w->dgrad || !w->dtheta) return sat_host::fail(SAT_EINVAL, "sat_backward_finish: null argument");
    if (w->B == 0) return SAT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int n_keys = (m->flags & SAT_FLAG_POS) ? m->n_dom * m->L * 2 : m->n_dom;
    // the grids the backward kernels were launched with (same formulas as backward_impl / sat_forward_tc)
    const int spt = NT / m->F, spt2 = NT2 / m->F;
    const int max_tiles = (w->B + spt - 1) / spt + m->n_dom, max_tiles2 = (w->B + spt2 - 1) / spt2 + m->n_dom;
    const int grid2 = max_tiles < 148 * 2 ? max_tiles : 148 * 2, grid1 = max_tiles < 148 ? max_tiles : 148;
    const int grid_a = max_tiles2 < 148 * 2 ? max_tiles2 : 148 * 2;
    const bool head_done = w->head_bwd_done != 0;
    RedArgs ra;
    ra.m = *m; ra.gpart = w->gpart; ra.counters = w->counters; ra.dgrad = w->dgrad; ra.dtheta = w->dtheta;
    ra.g_attn = grid_a; ra.g_meta = passes == 0 ? grid1 : grid_a; ra.g_head = head_done ? grid_a : grid2; ra.spt_meta = passes == 0 ? spt : spt2;
    const long long total = (long long)m->L * LAYER_FLOATS + m->F * D + m->n_dense + 1 + (long long)n_keys * PSZ;
    long long blocks = (total + 63) / 64;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_dense_reduce<<<(int)blocks, 256, 0, s>>>(ra);
    const int rc = sat_host::check_launch("k_dense_reduce");
    if (rc) return rc;
    return sat_encoder_bwd(m, w, stream);
}

extern "C" int sat_backward(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, float dropout_p,
                            uint64_t seed, void* stream) {
    return backward_impl(m, x, id_kind, ld, w, dropout_p, seed, 0, stream);
}

extern "C" int sat_backward_tc(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, float dropout_p,
                               uint64_t seed, int32_t passes, void* stream) {
    if (passes < 1 || passes > 3) return sat_host::fail(SAT_EINVAL, "sat_backward_tc: passes must be 1, 2 or 3, got %d", passes);
    if (!w || !w->wpack || !w->tiles2) return sat_host::fail(SAT_EINVAL, "sat_backward_tc: workspace has no packed weights");
    if (m && (m->flags & (SAT_FLAG_GATE_Q | SAT_FLAG_GATE_K | SAT_FLAG_BILINEAR)))
        return sat_host::fail(SAT_EINVAL, "sat_backward_tc: the gate / bilinear variants run in sat_backward (CUDA-core kernels) only");
    return backward_impl(m, x, id_kind, ld, w, dropout_p, seed, passes, stream);
}
