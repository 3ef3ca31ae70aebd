// Fused SATrans backward (replaces torch autograd over the graph of reference satrans.py:197-256).
//
// Same mapping as the forward: one thread = one field token, a CTA = whole samples of one scenario.
// Per layer, two kernels:
//   k_bwd_attn : LayerNorm-out bwd, out-projection bwd (+dW_out), attention bwd -> dQ', dK', dV
//   k_bwd_meta : MetaNet-LayerNorm bwd, MetaNet bwd (+d theta), QKV projection bwd (+dW_q/k/v) -> dX
// Data gradients are thread-local GEMVs against weights in shared memory (as in the forward).  Weight
// gradients are reductions over tokens: the two operands are staged as [token][36] shared tiles and every
// thread owns a 4x4 block of the 32x32 product (64 blocks x 3 token groups = 192 threads), accumulated
// into shared-memory accumulators that are flushed to HBM with one atomicAdd per element per CTA.
// Attention uses the FlashAttention identity  sum_j P_ij dP_ij = dR_i . R_i  so no (F x F) tensor is
// stored; probabilities are recomputed from Q'/K' for the query-role pass and again for the key-role pass.
#include "bwd_common.cuh"

#ifndef SAT_BWDA_JC
#define SAT_BWDA_JC 4          // keys per basic block in the query-role pass
#endif
#ifndef SAT_BWDA_PREFETCH
#define SAT_BWDA_PREFETCH 0    // 1: request the phase-2 rows before the dW_out phase (spills at 255 registers: slower)
#endif

namespace sat {

// =====================================================================================================
// kernel A: LayerNorm-out bwd, out-projection bwd, attention bwd.
// =====================================================================================================
constexpr int ST_LD = 25;              // row stride of the per-token softmax statistics (odd: conflict-free)
constexpr int bwda_smem_floats(int nth) {
    return 1024 /*WO*/ + 1024 /*dWO*/ + 192 /*LN params*/ + 64 /*dgamma,dbeta*/ + 2 * nth * KV_LD /*tiles*/ +
           nth * ST_LD /*mx,sum,Dh per head + dropout key*/;
}

// NTH threads (= tokens) per CTA: 192 walks the `tiles` set of k_bucket, 128 the `tiles2` set; MINB = CTAs per SM the register
// budget is sized for.
template <int H, int NTH, int MINB>
__global__ void __launch_bounds__(NTH, MINB) k_bwd_attn(const BwdArgs a) {
    constexpr int DH = D / H;
    extern __shared__ __align__(16) float smem[];
    float* sWO = smem;
    float* sdWO = sWO + 1024;
    float* sLN = sdWO + 1024;           // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;            // dgamma[32] dbeta[32] of the layer's output LayerNorm
    float* sT0 = sdLN + 64;             // [NT][36]  r / K' / Q'
    float* sT1 = sT0 + NTH * KV_LD;      // [NT][36]  dO / V / dR
    float* sSt = sT1 + NTH * KV_LD;      // [NT][25]  per head: row max, 1/row sum, D = dR.R; [24] = attention-dropout key

    const SatModel& m = a.m;
    const int tid = threadIdx.x, lane = tid & 31, F = m.F, l = a.layer;
    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool relu_out = m.flags & SAT_FLAG_RELU, use_res = m.flags & SAT_FLAG_RES;
    const int n_tiles = a.counters[NTH == NT ? CNT_NTILES : CNT_NTILES2];
    const int s = tid / F, f = tid - s * F;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    const float* lay = m.dense + m.off_layers + (long long)l * LAYER_FLOATS;
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const float inv_sqrt_dh = 1.f / sqrtf((float)DH);   // scores * (1/sqrt(dh)); differs from the reference's division by <= 1 ulp

    copy_to_smem(sWO, lay + OFF_WO, 1024, tid, NTH);
    copy_to_smem(sLN, lay + OFF_LN_G, 192, tid, NTH);
    for (int i = tid; i < 1024; i += NTH) sdWO[i] = 0.f;
    if (tid < 64) sdLN[tid] = 0.f;
    const float* qln = sLN + 64;
    const float* kln = pos ? sLN + 128 : sLN + 64;

    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int ntok = cnt * F;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        __syncthreads();                                   // smem init / previous tile done

        float dr[D];
        {   // ---- phase 1: y = drop(act(r @ WO^T)) + x ; LN bwd ; dO ; dR -------------------------------
            float r[D], dy[D], o[D];
            zero(o);
            uint32_t relu_mask = 0xffffffffu;
            float rstd = 0.f;
            if (valid) {
                load_row(act_l + ACT_R * a.act_stride + tokpos, r);
                vm_dot<D, D, D>(sWO, r, o);
                sts_row(sT0 + tid * KV_LD, r);             // staged for dW_out now, re-read for D_i below (frees 32 registers)
                const DropCtx dco = make_drop(a.seed, token_gid, l, SITE_OUT, a.p_drop, true);
                float xin[D];
                load_row(act_l + ACT_X * a.act_stride + tokpos, xin);
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float fk[2];
                    drop_pair(dco, i >> 1, fk[0], fk[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (relu_out && o[i + e] <= 0.f) relu_mask &= ~(1u << (i + e));
                        float v = relu_out ? fmaxf(o[i + e], 0.f) : o[i + e];
                        v *= fk[e];
                        o[i + e] = use_res ? v + xin[i + e] : v;      // o := y (pre-LN)
                    }
                }
                rstd = normalize32(o);                     // o := yhat
                load_row(a.dx + tokpos, dy);
            } else {
                zero(r); zero(dy);
            }
            layernorm_bwd32(dy, o, rstd, sLN, sdLN, sdLN + 32, lane, valid);     // dy := dLoss/dy
            if (valid) {
                if (use_res) store_row(a.dx + tokpos, dy);                  // residual branch of dX
                else { float zr[D]; zero(zr); store_row(a.dx + tokpos, zr); }
                const DropCtx dco = make_drop(a.seed, token_gid, l, SITE_OUT, a.p_drop, true);
#pragma unroll
                for (int i = 0; i < D; i += 2) {                                                  // dy := dO
                    float f0, f1;
                    drop_pair(dco, i >> 1, f0, f1);
                    dy[i] = ((relu_mask >> i) & 1u) ? dy[i] * f0 : 0.f;
                    dy[i + 1] = ((relu_mask >> (i + 1)) & 1u) ? dy[i + 1] * f1 : 0.f;
                }
                zero(dr);
                vm_outer<D, D, D>(sWO, dy, dr);            // dR[i] = sum_j dO[j] WO[j][i]
                lds_row(sT0 + tid * KV_LD, r);
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; ++c) acc = fmaf(dr[h * DH + c], r[h * DH + c], acc);
                    sSt[tid * ST_LD + 16 + h] = acc;               // D_i[h] = dR_i . R_i  (own row, read back below)
                }
                sts_row(sT1 + tid * KV_LD, dy);
            }
        }
        // ---- phase 2: recompute Q' K' V, stage K'/V.  The rows are requested before the dW_out phase so that their HBM/L2
        //      latency overlaps it. ------------------------------------------------------------------------
        float q[D], kst[D], vst[D];
#if SAT_BWDA_PREFETCH
        if (valid) {
            load_row(act_l + ACT_ZK * a.act_stride + tokpos, kst);
            load_row(act_l + ACT_V * a.act_stride + tokpos, vst);
            load_row(act_l + ACT_ZQ * a.act_stride + tokpos, q);
        }
#endif
        __syncthreads();
        wgrad32<NTH>(sT1, sT0, ntok, sdWO, D, tid);             // dWO[j][i] += dO[j] r[i]   (A = dO, G = r)
        __syncthreads();
#if !SAT_BWDA_PREFETCH
        if (valid) {
            load_row(act_l + ACT_ZK * a.act_stride + tokpos, kst);
            load_row(act_l + ACT_V * a.act_stride + tokpos, vst);
            load_row(act_l + ACT_ZQ * a.act_stride + tokpos, q);
        }
#endif
        if (valid) {
            if (meta_k) { float mean, rstd; layernorm32(kst, kln, kln + 32, mean, rstd); }
            sts_row(sT0 + tid * KV_LD, kst);
            sts_row(sT1 + tid * KV_LD, vst);
            if (meta_q) { float mean, rstd; layernorm32(q, qln, qln + 32, mean, rstd); }
        }
        __syncthreads();
        // ---- query-role pass: dQ'_i = sum_j dS_ij K'_j / sqrt(dh) ------------------------------------
        if (valid) {
            float dq[D];
            zero(dq);
            const DropCtx dca = make_drop(a.seed, token_gid, l, SITE_ATT, a.p_drop, true);
            const int sbase = s * F;
            sSt[tid * ST_LD + 24] = __uint_as_float(dca.key);          // read back by the key-role pass of the other tokens
            // rolled over heads: q / dr / dq are rotated by DH per head (compact code, see common.cuh)
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                // keys in chunks of JC per basic block; keys past F read row F-1 with score -inf (see attention_fwd)
                constexpr int JC = SAT_BWDA_JC;
                float sc[SAT_MAXF];
                float mx = -INFINITY;
                const float Dhh = sSt[tid * ST_LD + 16 + h];
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + JC; ++j) {
                            const float* kr = sT0 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            float acc = 0.f;
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                                acc = fmaf(q[c + 0], kk.x, acc); acc = fmaf(q[c + 1], kk.y, acc);
                                acc = fmaf(q[c + 2], kk.z, acc); acc = fmaf(q[c + 3], kk.w, acc);
                            }
                            sc[j] = j < F ? acc * inv_sqrt_dh : -INFINITY;
                            mx = fmaxf(mx, sc[j]);
                        }
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + JC; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
                    }
                }
                const float inv_sum = 1.f / sum;
                sSt[tid * ST_LD + h] = mx; sSt[tid * ST_LD + 8 + h] = inv_sum;
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
                    if (j0 < F) {
                        float fd[JC];
#pragma unroll
                        for (int j = j0; j < j0 + JC; j += 2) drop_pair(dca, (h * SAT_MAXF + j) >> 1, fd[j - j0], fd[j - j0 + 1]);
#pragma unroll
                        for (int j = j0; j < j0 + JC; ++j) {
                            const float pj = sc[j] * inv_sum;                  // exactly 0 for the padding keys
                            const float* vr = sT1 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            const float* kr = sT0 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            float dAd = 0.f;
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 vv = *reinterpret_cast<const float4*>(vr + c);
                                dAd = fmaf(dr[c + 0], vv.x, dAd); dAd = fmaf(dr[c + 1], vv.y, dAd);
                                dAd = fmaf(dr[c + 2], vv.z, dAd); dAd = fmaf(dr[c + 3], vv.w, dAd);
                            }
                            const float dsc = pj * (dAd * fd[j - j0] - Dhh) * inv_sqrt_dh;
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                                dq[c + 0] = fmaf(dsc, kk.x, dq[c + 0]); dq[c + 1] = fmaf(dsc, kk.y, dq[c + 1]);
                                dq[c + 2] = fmaf(dsc, kk.z, dq[c + 2]); dq[c + 3] = fmaf(dsc, kk.w, dq[c + 3]);
                            }
                        }
                    }
                }
                if (H > 1) { rotate_regs<D, DH>(q); rotate_regs<D, DH>(dr); rotate_regs<D, DH>(dq); }
            }
            store_row(a.dqkv + 0 * a.act_stride + tokpos, dq);
        }
        // ---- swap roles: own K', V to registers, tiles <- Q', dR ----------------------------------------
        float kreg[D], vreg[D];
        if (valid) { lds_row(sT0 + tid * KV_LD, kreg); lds_row(sT1 + tid * KV_LD, vreg); }
        __syncthreads();
        if (valid) { sts_row(sT0 + tid * KV_LD, q); sts_row(sT1 + tid * KV_LD, dr); }
        __syncthreads();
        // ---- key-role pass: dK'_j = sum_i dS_ij Q'_i / sqrt(dh) ; dV_j = sum_i Pd_ij dR_i ----------------
        // Head by head (rolled; kreg / vreg rotate by DH): the head's dK'/dV slices are complete after its F queries and go
        // straight to HBM, so only 2*DH accumulators are live (32-wide dk/dv next to kreg/vreg spilled to local memory).
        if (valid) {
            const int sbase = s * F;
            float* dk_out = a.dqkv + 1 * a.act_stride + tokpos;
            float* dv_out = a.dqkv + 2 * a.act_stride + tokpos;
            DropCtx dci = make_drop(a.seed, 0u, l, SITE_ATT, a.p_drop, true);     // .key is replaced per query below
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float dk[DH], dv[DH];
                zero(dk); zero(dv);
                const uint32_t e = h * SAT_MAXF + f;
#pragma unroll 2
                for (int i = 0; i < F; ++i) {
                    const float* qr = sT0 + (sbase + i) * KV_LD + h * DH;
                    const float* dR = sT1 + (sbase + i) * KV_LD + h * DH;
                    const float* st = sSt + (sbase + i) * ST_LD;
                    float qq[DH], dd[DH];
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 q4 = *reinterpret_cast<const float4*>(qr + c);
                        const float4 d4 = *reinterpret_cast<const float4*>(dR + c);
                        qq[c] = q4.x; qq[c + 1] = q4.y; qq[c + 2] = q4.z; qq[c + 3] = q4.w;
                        dd[c] = d4.x; dd[c + 1] = d4.y; dd[c + 2] = d4.z; dd[c + 3] = d4.w;
                    }
                    float acc = 0.f, dAd = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; ++c) { acc = fmaf(qq[c], kreg[c], acc); dAd = fmaf(dd[c], vreg[c], dAd); }
                    const float pij = expf(acc * inv_sqrt_dh - st[h]) * st[8 + h];
                    dci.key = __float_as_uint(st[24]);
                    const float df = drop_factor(dci, e);
                    const float dsc = pij * (dAd * df - st[16 + h]) * inv_sqrt_dh;
                    const float pd = pij * df;
#pragma unroll
                    for (int c = 0; c < DH; ++c) {
                        dk[c] = fmaf(dsc, qq[c], dk[c]);
                        dv[c] = fmaf(pd, dd[c], dv[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    *reinterpret_cast<float4*>(dk_out + h * DH + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
                    *reinterpret_cast<float4*>(dv_out + h * DH + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
                }
                if (H > 1) { rotate_regs<D, DH>(kreg); rotate_regs<D, DH>(vreg); }
            }
        }
    }
    __syncthreads();
    // flush: dWO and the output-LayerNorm parameter gradients
    float* dlay = a.dgrad + m.off_layers + (long long)l * LAYER_FLOATS;
    flush_add<NTH>(dlay + OFF_WO, sdWO, 1024, tid);
    flush_add<NTH>(dlay + OFF_LN_G, sdLN, 64, tid);
}

}  // namespace sat

using namespace sat;

#ifndef SAT_INST_H
#error "compile backward_attn.cu with -DSAT_INST_H={1,2,4,8}"
#endif
#define SAT_CAT_(a, b) a##b
#define SAT_CAT(a, b) SAT_CAT_(a, b)

template <int H, int NTH, int MINB>
static int launch_variant(const BwdArgs& a, int grid, cudaStream_t s) {
    const size_t smem = bwda_smem_floats(NTH) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_bwd_attn<H, NTH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = true;
    }
    k_bwd_attn<H, NTH, MINB><<<grid, NTH, smem, s>>>(a);
    return sat_host::check_launch("k_bwd_attn");
}

// variant 0: 192 threads x 2 CTAs/SM over a.tiles; 1: 128 x 2 and 2: 128 x 3 over the 128-token tile set (caller passes it in a.tiles)
int SAT_CAT(sat_launch_bwd_attn_h, SAT_INST_H)(const BwdArgs& a, int variant, int grid, cudaStream_t s) {
    constexpr int H = SAT_INST_H;
    if (variant == 1) return launch_variant<H, NT2, 2>(a, grid, s);
    if (variant == 2) return launch_variant<H, NT2, 3>(a, grid, s);
    return launch_variant<H, NT, 2>(a, grid, s);
}
