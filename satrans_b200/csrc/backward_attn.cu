// Fused SATrans backward (replaces torch autograd over the graph of reference satrans.py:197-256).
//
// Same mapping as the forward: one thread = one field token, a CTA = whole samples of one scenario
// (128 threads, the 128-token tile set of k_bucket).  Per layer, two kernels:
//   k_bwd_attn : LayerNorm-out bwd, out-projection bwd (+dW_out), attention bwd -> dQ', dK', dV      (this file)
//   k_bwd_meta : MetaNet-LayerNorm bwd, MetaNet bwd (+d theta), QKV projection bwd (+dW_q/k/v) -> dX
// Data gradients are thread-local GEMVs against weights in shared memory (as in the forward).  The weight
// gradient dW_out is a reduction over tokens: dO and r are staged as [token][36] shared tiles and every
// thread owns a 4x4 block of the 32x32 product for half of the tokens (bwd_common.cuh).
// Attention uses the FlashAttention identity  sum_j P_ij dP_ij = dR_i . R_i  so no (F x F) tensor is
// stored; probabilities are recomputed from Q'/K' for the query-role pass and again for the key-role pass.
//
// Four operand tiles: T0 = r -> Q', T1 = dO -> dR, T2 = K', T3 = V.  The r / z_k / V rows of a tile are
// fetched with cp.async at the top of the tile (no registers, all three latencies overlap the first phase).
#include "bwd_common.cuh"

#ifndef SAT_BWDA_JC
#define SAT_BWDA_JC 4          // keys per basic block in the query-role pass
#endif
#ifndef SAT_BWDA_IC
#define SAT_BWDA_IC 4          // queries per basic block in the key-role pass
#endif

namespace sat {

constexpr int ST_LD = 25;              // row stride of the per-token softmax statistics (odd: conflict-free)
constexpr int BA_NT = NT2;             // threads per CTA

static size_t bwda_smem_bytes(int F) {
    const int rows = ((BA_NT / F) * F + 3) & ~3;
    return (size_t)(1024 /*WO*/ + 1024 /*dWO*/ + 192 /*LN params*/ + (BA_NT / 32) * 64 /*dgamma,dbeta per warp*/ + 2048 /*dWO partial blocks*/ +
                    4 * rows * KV_LD /*tiles*/ + BA_NT * ST_LD /*mx,sum,Dh per head + dropout key*/) * sizeof(float);
}

// one 128-byte row global -> shared, asynchronously (8 x 16 B)
__device__ __forceinline__ void cp_async_row(float* __restrict__ dst, const float* __restrict__ src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
    for (int i = 0; i < D; i += 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int H>
__global__ void __launch_bounds__(BA_NT, 2) k_bwd_attn(const BwdArgs a) {
    constexpr int DH = D / H;
    extern __shared__ __align__(16) float smem[];
    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, lane = tid & 31, hsel = tid >> 6, F = m.F, l = a.layer;
    const int tile_rows = ((BA_NT / F) * F + 3) & ~3;
    float* sWO = smem;
    float* sdWO = sWO + 1024;
    float* sLN = sdWO + 1024;           // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;            // per warp: dgamma[32] dbeta[32] of the layer's output LayerNorm
    float* sPart = sdLN + (BA_NT / 32) * 64;   // [2][64][16] partial dW_out blocks of the two token halves
    float* sT0 = sPart + 2048;          // [rows][36]  r  -> Q'
    float* sT1 = sT0 + tile_rows * KV_LD;      //      dO -> dR
    float* sT2 = sT1 + tile_rows * KV_LD;      //      K'
    float* sT3 = sT2 + tile_rows * KV_LD;      //      V
    float* sSt = sT3 + tile_rows * KV_LD;      // [128][25]  per head: row max, 1/row sum, D = dR.R; [24] = attention-dropout key

    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool relu_out = m.flags & SAT_FLAG_RELU, use_res = m.flags & SAT_FLAG_RES;
    const int n_tiles = a.counters[CNT_NTILES2];
    const int s = tid / F, f = tid - s * F;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    const float* lay = m.dense + m.off_layers + (long long)l * LAYER_FLOATS;
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const float inv_sqrt_dh = 1.f / sqrtf((float)DH);   // scores * (1/sqrt(dh)); differs from the reference's division by <= 1 ulp
    const float sm_scale = inv_sqrt_dh;
    const bool stage = tid < tile_rows;                 // owns a row of the operand tiles (zeros past the last token)

    copy_to_smem(sWO, lay + OFF_WO, 1024, tid, BA_NT);
    copy_to_smem(sLN, lay + OFF_LN_G, 192, tid, BA_NT);
    for (int i = tid; i < 1024; i += BA_NT) sdWO[i] = 0.f;
    for (int i = tid; i < (BA_NT / 32) * 64; i += BA_NT) sdLN[i] = 0.f;
    const float* qln = sLN + 64;
    const float* kln = pos ? sLN + 128 : sLN + 64;

    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int ntok4 = (cnt * F + 3) & ~3;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        __syncthreads();                                   // smem init / previous tile done with the tiles

        if (valid) {
            cp_async_row(sT0 + tid * KV_LD, act_l + ACT_R * a.act_stride + tokpos);
            cp_async_row(sT2 + tid * KV_LD, act_l + ACT_ZK * a.act_stride + tokpos);
            cp_async_row(sT3 + tid * KV_LD, act_l + ACT_V * a.act_stride + tokpos);
        } else if (stage) {
            float z[D];
            zero(z);
            sts_row(sT0 + tid * KV_LD, z);
        }
        float dr[D];
        {   // ---- phase 1: y = drop(act(r @ WO^T)) + x ; LN bwd ; dO ; dR -------------------------------
            float r[D], dy[D], o[D];
            zero(o);
            uint32_t relu_mask = 0xffffffffu;
            float rstd = 0.f;
            if (valid && !relu_out) {
                // y (pre-LayerNorm output of the layer) was saved by the forward: nothing of the output projection is recomputed
                load_row(act_l + ACT_Y * a.act_stride + tokpos, o);
                load_row(a.dx + tokpos, dy);
                cp_async_wait_all();                       // own rows of T0 / T2 / T3 have landed
                rstd = normalize32(o);                     // o := yhat
            } else if (valid) {
                // 'relu' flag: the sign pattern of r @ WO^T is needed, so the projection is recomputed
                float xin[D];
                load_row(act_l + ACT_X * a.act_stride + tokpos, xin);
                load_row(a.dx + tokpos, dy);
                cp_async_wait_all();                       // own rows of T0 / T2 / T3 have landed
                lds_row(sT0 + tid * KV_LD, r);
                vm_dot<D, D, D>(sWO, r, o);
                const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, true);
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float fk[2];
                    drop_pair(dco, i >> 1, fk[0], fk[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (o[i + e] <= 0.f) relu_mask &= ~(1u << (i + e));
                        float v = fmaxf(o[i + e], 0.f);
                        v *= fk[e];
                        o[i + e] = use_res ? v + xin[i + e] : v;      // o := y (pre-LN)
                    }
                }
                rstd = normalize32(o);                     // o := yhat
            } else {
                zero(dy);
            }
            layernorm_bwd32(dy, o, rstd, sLN, sdLN + (tid >> 5) * 64, sdLN + (tid >> 5) * 64 + 32, lane, valid);   // this warp's slots; dy := dLoss/dy
            if (valid) {
                if (use_res) store_row(a.dx + tokpos, dy);                  // residual branch of dX
                else { float zr[D]; zero(zr); store_row(a.dx + tokpos, zr); }
                const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, true);
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
            }
            if (stage) sts_row(sT1 + tid * KV_LD, dy);             // zeros past the last token
        }
        // ---- phase 2: K' = LN(z_k) in place in T2 (own row); Q' to registers -------------------------------
        float q[D];
        if (valid) {
            if (meta_k) {
                float k[D];
                lds_row(sT2 + tid * KV_LD, k);
                float mean, rstd;
                layernorm32(k, kln, kln + 32, mean, rstd);
                sts_row(sT2 + tid * KV_LD, k);
            }
            load_row(act_l + ACT_ZQ * a.act_stride + tokpos, q);
            if (meta_q) { float mean, rstd; layernorm32(q, qln, qln + 32, mean, rstd); }
        }
        __syncthreads();                                   // r, dO, K', V tiles complete
        {   // dW_out[j][i] += dO[j] r[i]: each 64-thread half takes half of the token groups and parks its partial block
            float cw[4][4];
            const int mid = ((ntok4 >> 1) + 3) & ~3;
            wgrad_block(sT1, sT0, hsel ? mid : 0, hsel ? ntok4 : mid, tid, cw);
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4)
                *reinterpret_cast<float4*>(sPart + hsel * 1024 + (tid & 63) * 16 + r4 * 4) =
                    make_float4(cw[r4][0], cw[r4][1], cw[r4][2], cw[r4][3]);
        }
        // ---- query-role pass: dQ'_i = sum_j dS_ij K'_j / sqrt(dh) ------------------------------------
        if (valid) {
            float dq[D];
            zero(dq);
            const DropCtx dca = make_drop(seed32, token_gid, l, SITE_ATT, a.p_drop, true);
            const int sbase = s * F;
            sSt[tid * ST_LD + 24] = __uint_as_float(dca.key);          // read back by the key-role pass of the other tokens
            // rolled over heads: q / dr / dq are rotated by DH per head (compact code, see common.cuh)
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                // keys in chunks of JC per basic block; keys past F read row F-1 with score -inf (see attention_fwd)
                constexpr int JC = SAT_BWDA_JC;
                static_assert(JC % 2 == 0 && SAT_MAXF % JC == 0, "keys are processed in pairs");
                float sc[SAT_MAXF];
                float mx = -INFINITY;
                const float Dhh = sSt[tid * ST_LD + 16 + h];
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + JC; ++j) {
                            const float* kr = sT2 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            float acc = 0.f, acc1 = 0.f;     // even / odd channels, same order as attention_fwd
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                                fma2v(acc, acc1, q[c + 0], q[c + 1], kk.x, kk.y);
                                fma2v(acc, acc1, q[c + 2], q[c + 3], kk.z, kk.w);
                            }
                            acc += acc1;
                            sc[j] = j < F ? acc * sm_scale : -INFINITY;
                            mx = fmaxf(mx, sc[j]);
                        }
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + JC; j += 2) { softmax_exp_pair(sc[j], sc[j + 1], mx); sum += sc[j]; sum += sc[j + 1]; }
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
                            const float* vr = sT3 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            const float* kr = sT2 + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            float dAd = 0.f, dAd1 = 0.f;
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 vv = *reinterpret_cast<const float4*>(vr + c);
                                fma2v(dAd, dAd1, dr[c + 0], dr[c + 1], vv.x, vv.y);
                                fma2v(dAd, dAd1, dr[c + 2], dr[c + 3], vv.z, vv.w);
                            }
                            dAd += dAd1;
                            const float dsc = pj * (dAd * fd[j - j0] - Dhh) * inv_sqrt_dh;
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                                fma2(dq[c + 0], dq[c + 1], dsc, kk.x, kk.y);
                                fma2(dq[c + 2], dq[c + 3], dsc, kk.z, kk.w);
                            }
                        }
                    }
                }
                if (H > 1) { rotate_regs<D, DH>(q); rotate_regs<D, DH>(dr); rotate_regs<D, DH>(dq); }
            }
            store_row(a.dqkv + 0 * a.act_stride + tokpos, dq);
        }
        __syncthreads();                                   // dW_out phase is done with T0 / T1; partial blocks are visible
        if (!hsel) {
            float cw[4][4];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                const float4 lo = *reinterpret_cast<const float4*>(sPart + tid * 16 + r4 * 4);
                const float4 hi = *reinterpret_cast<const float4*>(sPart + 1024 + tid * 16 + r4 * 4);
                cw[r4][0] = lo.x + hi.x; cw[r4][1] = lo.y + hi.y; cw[r4][2] = lo.z + hi.z; cw[r4][3] = lo.w + hi.w;
            }
            wgrad_store(sdWO, D, tid, cw);
        }
        if (valid) { sts_row(sT0 + tid * KV_LD, q); sts_row(sT1 + tid * KV_LD, dr); }
        __syncthreads();
        // ---- key-role pass: dK'_j = sum_i dS_ij Q'_i / sqrt(dh) ; dV_j = sum_i Pd_ij dR_i ----------------
        // Head by head (rolled): the head's dK'/dV slices are complete after its F queries and go straight to HBM; the own
        // K'/V slices come from T2 / T3.  Queries in chunks of IC per basic block; queries past F read row F-1 with weight 0.
        if (valid) {
            constexpr int IC = SAT_BWDA_IC;
            const int sbase = s * F;
            float* dk_out = a.dqkv + 1 * a.act_stride + tokpos;
            float* dv_out = a.dqkv + 2 * a.act_stride + tokpos;
            DropCtx dci = make_drop(seed32, 0u, l, SITE_ATT, a.p_drop, true);     // .key is replaced per query below
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float kh[DH], vh[DH], dk[DH], dv[DH];
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(sT2 + tid * KV_LD + h * DH + c);
                    const float4 v4 = *reinterpret_cast<const float4*>(sT3 + tid * KV_LD + h * DH + c);
                    kh[c] = k4.x; kh[c + 1] = k4.y; kh[c + 2] = k4.z; kh[c + 3] = k4.w;
                    vh[c] = v4.x; vh[c + 1] = v4.y; vh[c + 2] = v4.z; vh[c + 3] = v4.w;
                }
                zero(dk); zero(dv);
                const uint32_t e = h * SAT_MAXF + f;
#pragma unroll 1
                for (int i0 = 0; i0 < F; i0 += IC) {
#pragma unroll
                    for (int ii = 0; ii < IC; ++ii) {
                        const int i = min(i0 + ii, F - 1);
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
                        float acc = 0.f, acc1 = 0.f, dAd = 0.f, dAd1 = 0.f;
#pragma unroll
                        for (int c = 0; c < DH; c += 2) {
                            fma2v(acc, acc1, qq[c], qq[c + 1], kh[c], kh[c + 1]);
                            fma2v(dAd, dAd1, dd[c], dd[c + 1], vh[c], vh[c + 1]);
                        }
                        acc += acc1; dAd += dAd1;
                        const float pij = (i0 + ii < F) ? softmax_exp(acc * sm_scale - st[h]) * st[8 + h] : 0.f;
                        dci.key = __float_as_uint(st[24]);
                        const float df = drop_factor(dci, e);
                        const float dsc = pij * (dAd * df - st[16 + h]) * inv_sqrt_dh;
                        const float pd = pij * df;
#pragma unroll
                        for (int c = 0; c < DH; c += 2) {
                            fma2(dk[c], dk[c + 1], dsc, qq[c], qq[c + 1]);
                            fma2(dv[c], dv[c + 1], pd, dd[c], dd[c + 1]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    *reinterpret_cast<float4*>(dk_out + h * DH + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
                    *reinterpret_cast<float4*>(dv_out + h * DH + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
                }
            }
        }
    }
    __syncthreads();
    // this CTA's row of the partial buffer: dWO and the output-LayerNorm parameter gradients (k_dense_reduce adds the rows)
    const GPart gp = gpart_layout(m);
    float* part = a.gpart + gp.attn + ((long long)l * SAT_GRID_MAX + blockIdx.x) * GP_ATTN;
    store_partial<BA_NT>(part, sdWO, 1024, tid);
    store_partial_warps<BA_NT>(part + 1024, sdLN, 64, tid);
}

}  // namespace sat

using namespace sat;

#ifndef SAT_INST_H
#error "compile backward_attn.cu with -DSAT_INST_H={1,2,4,8}"
#endif
#define SAT_CAT_(a, b) a##b
#define SAT_CAT(a, b) SAT_CAT_(a, b)

// a.tiles must be the 128-token tile set (SatWork.tiles2)
int SAT_CAT(sat_launch_bwd_attn_h, SAT_INST_H)(const BwdArgs& a, int grid, cudaStream_t s) {
    constexpr int H = SAT_INST_H;
    const size_t smem = bwda_smem_bytes(a.m.F);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_bwd_attn<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    k_bwd_attn<H><<<grid, BA_NT, smem, s>>>(a);
    return sat_host::check_launch("k_bwd_attn");
}
