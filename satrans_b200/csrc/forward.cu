// Fused SATrans forward: embedding gather + L x Meta_Transformer_Layer + logit head + sigmoid (+ BCE).
//
// Replaces, per batch, the ~1.2 k ATen dispatches of SATrans.forward (reference satrans.py:197-256),
// Meta_Transformer_Layer.forward (satrans.py:50-100) and MetaNet.forward (submodules.py:77-103).
//
// Mapping: one thread = one field token (sample b, field f); its 32-channel vector lives in registers
// from the table row to the logit.  A CTA holds floor(192/F) whole samples of ONE scenario (tiles come
// from sat_bucket_by_domain), so the layer weights and that scenario's generated MetaNet weights are
// CTA-uniform: they sit in shared memory and every weight read is a broadcast LDS.128 feeding 4 FFMAs.
// Only K' and V cross threads (attention over the F tokens of a sample): they are staged in padded
// [token][36] shared tiles.  LayerNorm / softmax are thread-local register reductions.
// Nothing of shape (B,4096), (H,B,F,F) or (B,F,32)-per-op is ever materialised in HBM (inference), and
// training saves only what the backward kernels re-use (x, pre-LN q/k, v, attention output).
#include "fwd_common.cuh"

namespace sat {


constexpr int FWD_SMEM_FLOATS = 4096 + 192 + 2 * PSZ + 2 * NT * KV_LD + NT;

// MetaNet (reference submodules.py:77-103) on one token: z = dropout(relu(q @ W1) @ W2) + q.
// th = [W1 (32x64) | W2 (64x32)] in shared memory.  Hidden units are produced and consumed 16 at a time.
__device__ __forceinline__ void metanet_fwd(float (&q)[D], const float* __restrict__ th, const DropCtx& dc,
                                            float (&z)[D], uint32_t (&mask)[2]) {
    float mo[D];
    zero(mo);
    mask[0] = mask[1] = 0u;
#pragma unroll 1
    for (int c0 = 0; c0 < HID; c0 += 16) {
        float h[16];
        zero(h);
        vm_outer<D, 16, HID>(th + c0, q, h);
        uint32_t bits = 0u;
#pragma unroll
        for (int c = 0; c < 16; ++c) { if (h[c] > 0.f) bits |= 1u << c; h[c] = fmaxf(h[c], 0.f); }
        if (c0 < 32) mask[0] |= bits << c0; else mask[1] |= bits << (c0 - 32);
        vm_outer<16, D, D>(th + D * HID + c0 * D, h, mo);
    }
#pragma unroll
    for (int i = 0; i < D; i += 2) {
        float f0, f1;
        drop_pair(dc, i >> 1, f0, f1);
        z[i] = fmaf(mo[i], f0, q[i]); z[i + 1] = fmaf(mo[i + 1], f1, q[i + 1]);
    }
}

template <int H, bool TRAIN>
__global__ void __launch_bounds__(NT, 2) k_forward(const FwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                    // WQ | WK | WV | WO           (4096)
    float* sLN = sW + 4096;              // 6 x 32 LayerNorm vectors    (192)  -- contiguous with sW in the layer block
    float* sThQ = sLN + 192;             // generated W1|W2 for the Q side
    float* sThK = sThQ + PSZ;            // ... K side (ENP only)
    float* sK = sThK + PSZ;              // [NT][36]
    float* sV = sK + NT * KV_LD;         // [NT][36]
    float* sRed = sV + NT * KV_LD;       // [NT]
    __shared__ double s_lossv[NT];      // per-sample loss terms of the tile, added in sample order (no atomics)

    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x;
    const int F = m.F, L = m.L;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool relu_out = m.flags & SAT_FLAG_RELU, use_res = m.flags & SAT_FLAG_RES;
    const int n_tiles = *a.ntiles_ptr;
    const int s = tid / F, f = tid - s * F;
    int cur_q = -1, cur_k = -1;

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int g = td.x, start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;

        // ---- gather: x = T_f[id]  (reference meta_basemodel.py:533-535) -------------------------------
        float x[D];
        zero(x);
        if (valid) {
            long long id = load_id(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.sparse_cols + f));
            const long long lo = __ldg(m.row_off + f), hi = __ldg(m.row_off + f + 1);
            if (id < 0 || id >= hi - lo) { atomicOr(&a.counters[CNT_ERR], 1); id = 0; }
            load_row(table_row(m, lo + id), x);
        }

        for (int l = 0; l < L; ++l) {
            __syncthreads();
            copy_to_smem(sW, m.dense + m.off_layers + (long long)l * LAYER_FLOATS, LAYER_FLOATS, tid, NT);
            const int kq = theta_key(g, l, 0, L, pos), kk = theta_key(g, l, 1, L, pos);
            if (kq != cur_q) { copy_to_smem(sThQ, a.theta + (long long)kq * PSZ, PSZ, tid, NT); cur_q = kq; }
            if (pos && kk != cur_k) { copy_to_smem(sThK, a.theta + (long long)kk * PSZ, PSZ, tid, NT); cur_k = kk; }
            __syncthreads();
            const float* thK = pos ? sThK : sThQ;
            const float* kln = pos ? sLN + 128 : sLN + 64;      // K-side MetaNet LayerNorm (aliases Q's without 'pos')
            float* act_l = TRAIN ? a.acts + (long long)l * ACT_PER_LAYER * a.act_stride : nullptr;
            float q[D];
            if (valid) {
                if (TRAIN) store_row(act_l + ACT_X * a.act_stride + tokpos, x);
                float mean, rstd;
                {   // K' = LN(MetaNet(x @ W_Key))            (satrans.py:56,67-73)
                    float kp[D];
                    zero(kp);
                    vm_outer<D, D, D>(sW + OFF_WK, x, kp);
                    if (meta_k) {
                        float z[D];
                        uint32_t mk[2];
                        metanet_fwd(kp, thK, make_drop(seed32, token_gid, l, SITE_K, a.p_drop, TRAIN), z, mk);
                        if (TRAIN) {
                            store_row(act_l + ACT_ZK * a.act_stride + tokpos, z);
                            *reinterpret_cast<uint2*>(relu_mask_ptr(a.acts, a.act_stride, L, l, 1, tokpos / D)) = make_uint2(mk[0], mk[1]);
                        }
                        layernorm32(z, kln, kln + 32, mean, rstd);
                        sts_row(sK + tid * KV_LD, z);
                    } else {
                        if (TRAIN) store_row(act_l + ACT_ZK * a.act_stride + tokpos, kp);
                        sts_row(sK + tid * KV_LD, kp);
                    }
                }
                {   // V = x @ W_Value                         (satrans.py:57)
                    float v[D];
                    zero(v);
                    vm_outer<D, D, D>(sW + OFF_WV, x, v);
                    if (TRAIN) store_row(act_l + ACT_V * a.act_stride + tokpos, v);
                    sts_row(sV + tid * KV_LD, v);
                }
                {   // Q' = LN(MetaNet(x @ W_Query))           (satrans.py:55,60-66)
                    float qp[D];
                    zero(qp);
                    vm_outer<D, D, D>(sW + OFF_WQ, x, qp);
                    if (meta_q) {
                        uint32_t mq[2];
                        metanet_fwd(qp, sThQ, make_drop(seed32, token_gid, l, SITE_Q, a.p_drop, TRAIN), q, mq);
                        if (TRAIN) {
                            store_row(act_l + ACT_ZQ * a.act_stride + tokpos, q);
                            *reinterpret_cast<uint2*>(relu_mask_ptr(a.acts, a.act_stride, L, l, 0, tokpos / D)) = make_uint2(mq[0], mq[1]);
                        }
                        layernorm32(q, sLN + 64, sLN + 96, mean, rstd);
                    } else {
#pragma unroll
                        for (int i = 0; i < D; ++i) q[i] = qp[i];
                        if (TRAIN) store_row(act_l + ACT_ZQ * a.act_stride + tokpos, q);
                    }
                }
            }
            __syncthreads();
            if (valid) {
                float r[D];
                attention_fwd<H>(q, sK, sV, s * F, F, make_drop(seed32, token_gid, l, SITE_ATT, a.p_drop, TRAIN), r);
                if (TRAIN) store_row(act_l + ACT_R * a.act_stride + tokpos, r);
                // O = dropout(act(R @ W_out^T)); x = LN(O + x)   (satrans.py:91-99)
                float o[D];
                zero(o);
                vm_dot<D, D, D>(sW + OFF_WO, r, o);
                const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, TRAIN);
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float fk[2];
                    drop_pair(dco, i >> 1, fk[0], fk[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float v = relu_out ? fmaxf(o[i + e], 0.f) : o[i + e];
                        v *= fk[e];
                        x[i + e] = use_res ? v + x[i + e] : v;
                    }
                }
                if (TRAIN) store_row(act_l + ACT_Y * a.act_stride + tokpos, x);      // y (pre-LayerNorm): k_bwd_attn starts from it
                float mean, rstd;
                layernorm32(x, sLN, sLN + 32, mean, rstd);
            }
        }

        // ---- logit head + sigmoid (+ BCE)  (satrans.py:244-255, meta_basemodel.py:317) ---------------
        float part = 0.f;
        if (valid) {
            if (TRAIN) store_row(a.acts + (long long)L * ACT_PER_LAYER * a.act_stride + tokpos, x);
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
            for (int k = 0; k < m.n_dense; ++k)
                z = fmaf(load_dense(a.x, a.id_kind, (long long)b * a.ld + __ldg(m.dense_cols + k)),
                         __ldg(m.dense + m.off_headw + F * D + k), z);
            z += __ldg(m.dense + m.off_headb);
            const float yp = 1.f / (1.f + expf(-z));
            a.y_pred[b] = yp;
            double lossv = 0.0;
            if (a.y) {
                const float yt = __ldg(a.y + b);
                const float pq = (1.f - yp) * yp;            // d sigmoid / d logit
                if (m.flags & SAT_FLAG_LOSS_MSE) {           // F.mse_loss(reduction='sum')
                    const float d = yp - yt;
                    lossv = (double)(d * d);
                    a.dz[p] = 2.f * d * pq;
                } else if (m.flags & SAT_FLAG_LOSS_MAE) {    // F.l1_loss(reduction='sum')
                    const float d = yp - yt;
                    lossv = (double)fabsf(d);
                    a.dz[p] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * pq;
                } else {
                    // F.binary_cross_entropy(reduction='sum'): logs clamped at -100; autograd of BCE then sigmoid:
                    // (p - y) / max((1-p) p, 1e-12) * (1-p) p
                    const float lp = fmaxf(logf(yp), -100.f), l1p = fmaxf(logf(1.f - yp), -100.f);
                    lossv = (double)(-(yt * lp + (1.f - yt) * l1p));
                    a.dz[p] = (yp - yt) / fmaxf(pq, 1e-12f) * pq;
                }
            }
            s_lossv[s] = lossv;
        }
        __syncthreads();
        if (tid == 0 && a.loss_part) {
            double acc = 0.0;
            for (int ss = 0; ss < cnt; ++ss) acc += s_lossv[ss];
            a.loss_part[t] = acc;
        }
    }
}

}  // namespace sat

using namespace sat;

#ifndef SAT_INST_H
#error "compile forward.cu with -DSAT_INST_H={1,2,4,8}"
#endif
#define SAT_CAT_(a, b) a##b
#define SAT_CAT(a, b) SAT_CAT_(a, b)

int SAT_CAT(sat_launch_forward_h, SAT_INST_H)(const FwdArgs& a, int grid, int training, cudaStream_t s) {
    constexpr int H = SAT_INST_H;
    const size_t smem = FWD_SMEM_FLOATS * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_forward<H, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_forward<H, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_forward: cannot raise dynamic shared memory to %zu", smem);
        configured = true;
    }
    if (training) k_forward<H, true><<<grid, NT, smem, s>>>(a);
    else k_forward<H, false><<<grid, NT, smem, s>>>(a);
    return sat_host::check_launch("k_forward");
}
