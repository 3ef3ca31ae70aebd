// k_bwd_attn with every token split over TWO threads (SAT_BWD_ATTN_SPLIT, H >= 2): same math and the same tile set as
// backward_attn.cu, but a CTA is 256 threads = 128 tokens x 2 halves.  Thread (token, hf) owns channels [16 hf, 16 hf + 16) of
// every 32-channel vector of its token, i.e. heads [hf H/2, (hf + 1) H/2): attention stays thread-local per head, every
// channel-wise operation splits trivially, the two GEMVs read the full input row from the shared tiles, and the only cross-
// thread traffic of a token is the LayerNorm statistics (pair exchange through shared memory + a 64-thread named barrier of the
// two sibling warps w and w + 4).
//
// Why it was built: backward_attn.cu keeps a token's whole state in one thread -> 255 registers, 8 warps per SM, issue slots
// 36 % busy; the hypothesis was that the kernel is bound by the latency of each warp's serial chain, so that halving the chain
// per thread and doubling the warps (<= 128 registers, 16 warps per SM, same shared memory) would nearly double its rate.
// MEASURED (B200, config 2, profiles/r2_details_k_bwd_attn2.txt): parity-green on every gradient test, and 11 % SLOWER
// (245 vs 219 us per launch).  Twice the warps raised the IPC only from 1.39 to 1.49 while the issued instructions grew by
// 17 % (per-thread overheads -- addressing, dropout contexts, the GEMVs' full input rows -- do not halve, plus 416 B of spill
// loads at the 128-register cap), and the stall per issued instruction moved to the shared-memory pipe (short scoreboard
// 1.2 -> 3.4, MIO throttle 0.14 -> 0.35): what bounds the kernel is the latency of its LDS-fed inner loops (one 16-byte shared
// load per two packed FMAs), not the number of warps.  NOT the default (SAT_BWD_ATTN_SPLIT=1 selects it); kept, with its test,
// as the measured answer to "would more warps per SM help".
// The summation orders differ from backward_attn.cu in the last bits (LayerNorm statistics combined from two halves,
// weight-gradient blocks over four token groups instead of two); both are deterministic.
#include "bwd_common.cuh"

#ifndef SAT_BWDA_JC
#define SAT_BWDA_JC 4          // keys per basic block in the query-role pass
#endif
#ifndef SAT_BWDA_IC
#define SAT_BWDA_IC 4          // queries per basic block in the key-role pass
#endif

namespace sat {

constexpr int B2_NT = 256;             // threads per CTA
constexpr int B2_TOK = 128;            // tokens (rows) per CTA
constexpr int HD = D / 2;              // channels per thread
constexpr int ST2_LD = 25;             // row stride of the per-token softmax statistics (odd: conflict-free)

static size_t bwda2_smem_bytes(int F) {
    const int rows = ((B2_TOK / F) * F + 3) & ~3;
    return (size_t)(1024 /*WO*/ + 1024 /*dWO*/ + 192 /*LN params*/ + (B2_NT / 32) * 64 /*dgamma,dbeta per warp*/ + 4096 /*dWO partial blocks*/ +
                    4 * rows * KV_LD /*tiles*/ + B2_TOK * ST2_LD /*mx,sum,Dh per head + dropout key*/ + 2 * 2 * B2_TOK * 2 /*pair slots*/) * sizeof(float);
}

// half a row (64 B) global -> shared, asynchronously
__device__ __forceinline__ void cp_async_half(float* __restrict__ dst, const float* __restrict__ src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
#pragma unroll
    for (int i = 0; i < HD; i += 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all2() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void pair_sync(int warp) {           // the two warps that share 32 tokens
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
}
// exchange two floats with the thread that owns the other half of the token (double-buffered slots, one barrier per exchange)
__device__ __forceinline__ float2 pair_exchange(float a, float b, float2* __restrict__ slots, int tk, int hf, int warp, int& par) {
    slots[(par * 2 + hf) * B2_TOK + tk] = make_float2(a, b);
    pair_sync(warp);
    const float2 o = slots[(par * 2 + (hf ^ 1)) * B2_TOK + tk];
    par ^= 1;
    return o;
}
__device__ __forceinline__ void load_half(const float* __restrict__ p, float (&v)[HD]) {
    ldg256_nc(p, &v[0]); ldg256_nc(p + 8, &v[8]);
}
__device__ __forceinline__ void store_half(float* __restrict__ p, const float (&v)[HD]) {
    stg256(p, &v[0]); stg256(p + 8, &v[8]);
}
__device__ __forceinline__ void sts_half(float* __restrict__ dst, const float (&v)[HD]) {
#pragma unroll
    for (int i = 0; i < HD; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
__device__ __forceinline__ void lds_half(const float* __restrict__ src, float (&v)[HD]) {
#pragma unroll
    for (int i = 0; i < HD; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}
// LayerNorm statistics of a 32-channel row held as two 16-channel halves: local mean / sum of squared deviations, combined with
// the sibling's (Chan et al.): mean = (m0 + m1) / 2, M2 = M2_0 + M2_1 + 8 (m0 - m1)^2.  Symmetric in the halves.
__device__ __forceinline__ void ln_stats_pair(const float (&v)[HD], float2* slots, int tk, int hf, int warp, int& par, float& mean, float& rstd) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < HD; i += 2) { s0 += v[i]; s1 += v[i + 1]; }
    const float m_loc = (s0 + s1) * (1.f / HD);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < HD; i += 2) {
        const float t0 = v[i] - m_loc, t1 = v[i + 1] - m_loc;
        q0 = fmaf(t0, t0, q0); q1 = fmaf(t1, t1, q1);
    }
    const float q_loc = q0 + q1;
    const float2 o = pair_exchange(m_loc, q_loc, slots, tk, hf, warp, par);
    const float dm = m_loc - o.x;
    mean = 0.5f * (m_loc + o.x);
    const float m2 = (q_loc + o.y) + 8.f * dm * dm;
    rstd = 1.f / sqrtf(m2 * (1.f / D) + LN_EPS);
}
// on return lane L holds the sum over the 32 lanes of v[L & 15] (in v[0])
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[HD], int lane) {
#pragma unroll
    for (int i = 0; i < HD; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

template <int H>
__global__ void __launch_bounds__(B2_NT, 2) k_bwd_attn2(const BwdArgs a) {
    static_assert(H >= 2, "a head must not span the two halves of a token");
    constexpr int DH = D / H, HH = H / 2;                 // channels per head, heads per thread
    extern __shared__ __align__(16) float smem[];
    const SatModel& m = a.m;
    const uint32_t seed32 = step_seed32(a.seed, a.seed_ptr);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, hf = tid >> 7, tk = tid & (B2_TOK - 1);
    const int ch0 = hf * HD;                               // first channel of this thread
    const int F = m.F, l = a.layer;
    const int tile_rows = ((B2_TOK / F) * F + 3) & ~3;
    float* sWO = smem;
    float* sdWO = sWO + 1024;
    float* sLN = sdWO + 1024;           // LN_G LN_B QLN_G QLN_B KLN_G KLN_B
    float* sdLN = sLN + 192;            // per warp: dgamma[32] dbeta[32] of the layer's output LayerNorm
    float* sPart = sdLN + (B2_NT / 32) * 64;   // [4][64][16] partial dW_out blocks of the four token groups
    float* sT0 = sPart + 4096;          // [rows][36]  r  -> Q'
    float* sT1 = sT0 + tile_rows * KV_LD;      //      dO -> dR
    float* sT2 = sT1 + tile_rows * KV_LD;      //      K'
    float* sT3 = sT2 + tile_rows * KV_LD;      //      V
    float* sSt = sT3 + tile_rows * KV_LD;      // [128][25]  per head: row max, 1/row sum, D = dR.R; [24] = attention-dropout key
    float2* sPair = reinterpret_cast<float2*>(sSt + B2_TOK * ST2_LD);     // [2 parities][2 halves][128]

    const bool pos = m.flags & SAT_FLAG_POS;
    const bool meta_q = m.flags & SAT_FLAG_META_Q, meta_k = m.flags & SAT_FLAG_META_K;
    const bool relu_out = m.flags & SAT_FLAG_RELU, use_res = m.flags & SAT_FLAG_RES;
    const int n_tiles = a.counters[CNT_NTILES2];
    const int s = tk / F, f = tk - s * F;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    const float* lay = m.dense + m.off_layers + (long long)l * LAYER_FLOATS;
    const float* act_l = a.acts + (long long)l * ACT_PER_LAYER * a.act_stride;
    const float inv_sqrt_dh = 1.f / sqrtf((float)DH);
    const float sm_scale = inv_sqrt_dh;
    const bool stage = tk < tile_rows;                  // owns (half of) a row of the operand tiles (zeros past the last token)
    int par = 0;

    copy_to_smem(sWO, lay + OFF_WO, 1024, tid, B2_NT);
    copy_to_smem(sLN, lay + OFF_LN_G, 192, tid, B2_NT);
    for (int i = tid; i < 1024; i += B2_NT) sdWO[i] = 0.f;
    for (int i = tid; i < (B2_NT / 32) * 64; i += B2_NT) sdLN[i] = 0.f;
    const float* qln = sLN + 64;
    const float* kln = pos ? sLN + 128 : sLN + 64;
    float* const dgam = sdLN + warp * 64;
    float* const dbet = dgam + 32;

    for (int t = t0; t < t1; ++t) {
        const int4 td = *reinterpret_cast<const int4*>(a.tiles + 4 * t);
        const int start = td.y, cnt = td.z;
        const bool valid = s < cnt;
        const int ntok4 = (cnt * F + 3) & ~3;
        const int p = start + s;
        const int b = valid ? a.perm[p] : 0;
        const long long tokpos = ((long long)p * F + f) * D + ch0;          // this thread's half of the token's row
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        __syncthreads();                                   // smem init / previous tile done with the tiles

        if (valid) {
            cp_async_half(sT0 + tk * KV_LD + ch0, act_l + ACT_R * a.act_stride + tokpos);
            cp_async_half(sT2 + tk * KV_LD + ch0, act_l + ACT_ZK * a.act_stride + tokpos);
            cp_async_half(sT3 + tk * KV_LD + ch0, act_l + ACT_V * a.act_stride + tokpos);
        } else if (stage) {
            float z[HD];
            zero(z);
            sts_half(sT0 + tk * KV_LD + ch0, z);
        }
        float dr[HD];
        zero(dr);
        {   // ---- phase 1: y = drop(act(r @ WO^T)) + x ; LN bwd ; dO ; dR -------------------------------
            float dy[HD], o[HD];
            zero(o); zero(dy);
            uint32_t relu_mask = 0xffffu;
            float xin[HD];
            zero(xin);
            if (valid) {
                load_half(act_l + ACT_X * a.act_stride + tokpos, xin);
                load_half(a.dx + tokpos, dy);
            }
            cp_async_wait_all2();
            pair_sync(warp);                               // both halves of the own rows of T0 / T2 / T3 have landed
            if (valid) {
                float r[D];
                lds_row(sT0 + tk * KV_LD, r);
                vm_dot<D, HD, D>(sWO + ch0 * D, r, o);     // o[j] = sum_i r[i] WO[ch0 + j][i]
                const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, true);
#pragma unroll
                for (int i = 0; i < HD; i += 2) {
                    float fk[2];
                    drop_pair(dco, (ch0 + i) >> 1, fk[0], fk[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        if (relu_out && o[i + e] <= 0.f) relu_mask &= ~(1u << (i + e));
                        float v = relu_out ? fmaxf(o[i + e], 0.f) : o[i + e];
                        v *= fk[e];
                        o[i + e] = use_res ? v + xin[i + e] : v;      // o := y (pre-LN)
                    }
                }
            }
            float mean, rstd;
            ln_stats_pair(o, sPair, tk, hf, warp, par, mean, rstd);
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] = valid ? (o[i] - mean) * rstd : 0.f;      // o := yhat
            if (!valid) rstd = 0.f;
            {   // LayerNorm backward (this thread's 16 channels; the two row sums meet through the pair exchange)
                float t1v[HD], t2v[HD];
#pragma unroll
                for (int i = 0; i < HD; ++i) { t1v[i] = dy[i] * o[i]; t2v[i] = dy[i]; }
                const float sg = warp_reduce_scatter16(t1v, lane);
                const float sb = warp_reduce_scatter16(t2v, lane);
                if (lane < HD) { dgam[ch0 + lane] += sg; dbet[ch0 + lane] += sb; }
                float m1a = 0.f, m1b = 0.f, m2a = 0.f, m2b = 0.f;
#pragma unroll
                for (int i = 0; i < HD; i += 2) {
                    dy[i] *= sLN[ch0 + i]; dy[i + 1] *= sLN[ch0 + i + 1];
                    m1a += dy[i]; m1b += dy[i + 1];
                    m2a = fmaf(dy[i], o[i], m2a); m2b = fmaf(dy[i + 1], o[i + 1], m2b);
                }
                const float p1 = m1a + m1b, p2 = m2a + m2b;
                const float2 ot = pair_exchange(p1, p2, sPair, tk, hf, warp, par);
                const float m1 = (hf ? ot.x + p1 : p1 + ot.x) * (1.f / D), m2 = (hf ? ot.y + p2 : p2 + ot.y) * (1.f / D);
#pragma unroll
                for (int i = 0; i < HD; ++i) dy[i] = rstd * (dy[i] - m1 - o[i] * m2);     // dy := dLoss/dy
            }
            if (valid) {
                if (use_res) store_half(a.dx + tokpos, dy);                 // residual branch of dX
                else { float zr[HD]; zero(zr); store_half(a.dx + tokpos, zr); }
                const DropCtx dco = make_drop(seed32, token_gid, l, SITE_OUT, a.p_drop, true);
#pragma unroll
                for (int i = 0; i < HD; i += 2) {                                                  // dy := dO
                    float f0, f1;
                    drop_pair(dco, (ch0 + i) >> 1, f0, f1);
                    dy[i] = ((relu_mask >> i) & 1u) ? dy[i] * f0 : 0.f;
                    dy[i + 1] = ((relu_mask >> (i + 1)) & 1u) ? dy[i + 1] * f1 : 0.f;
                }
            }
            if (stage) sts_half(sT1 + tk * KV_LD + ch0, dy);           // zeros past the last token
            pair_sync(warp);                                           // the token's full dO row is in T1
            if (valid) {
                float dO[D];
                lds_row(sT1 + tk * KV_LD, dO);
                vm_outer<D, HD, D>(sWO + ch0, dO, dr);                 // dR[ch0 + i] = sum_j dO[j] WO[j][ch0 + i]
                float r[HD];
                lds_half(sT0 + tk * KV_LD + ch0, r);
#pragma unroll
                for (int h = 0; h < HH; ++h) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; ++c) acc = fmaf(dr[h * DH + c], r[h * DH + c], acc);
                    sSt[tk * ST2_LD + 16 + hf * HH + h] = acc;         // D_i[h] = dR_i . R_i  (own heads, read back below)
                }
            }
        }
        // ---- phase 2: K' = LN(z_k) in place in T2 (own half row); Q' to registers ---------------------------
        float q[HD];
        zero(q);
        {
            float k[HD];
            zero(k);
            if (valid) {
                lds_half(sT2 + tk * KV_LD + ch0, k);
                load_half(act_l + ACT_ZQ * a.act_stride + tokpos, q);
            }
            if (meta_k) {
                float mean, rstd;
                ln_stats_pair(k, sPair, tk, hf, warp, par, mean, rstd);
                if (valid) {
#pragma unroll
                    for (int i = 0; i < HD; ++i) k[i] = fmaf((k[i] - mean) * rstd, kln[ch0 + i], kln[32 + ch0 + i]);
                    sts_half(sT2 + tk * KV_LD + ch0, k);
                }
            }
            if (meta_q) {
                float mean, rstd;
                ln_stats_pair(q, sPair, tk, hf, warp, par, mean, rstd);
#pragma unroll
                for (int i = 0; i < HD; ++i) q[i] = fmaf((q[i] - mean) * rstd, qln[ch0 + i], qln[32 + ch0 + i]);
            }
        }
        __syncthreads();                                   // r, dO, K', V tiles complete
        {   // dW_out[j][i] += dO[j] r[i]: each 64-thread group takes a quarter of the token groups and parks its partial block
            float cw[4][4];
            const int g4 = tid >> 6;
            const int q4 = ((ntok4 >> 2) + 3) & ~3;
            const int tb = min(g4 * q4, ntok4), te = min(tb + q4, ntok4);
            wgrad_block(sT1, sT0, tb, g4 == 3 ? ntok4 : te, tid, cw);
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4)
                *reinterpret_cast<float4*>(sPart + g4 * 1024 + (tid & 63) * 16 + r4 * 4) =
                    make_float4(cw[r4][0], cw[r4][1], cw[r4][2], cw[r4][3]);
        }
        // ---- query-role pass: dQ'_i = sum_j dS_ij K'_j / sqrt(dh) (own heads) -------------------------
        if (valid) {
            float dq[HD];
            zero(dq);
            const DropCtx dca = make_drop(seed32, token_gid, l, SITE_ATT, a.p_drop, true);
            const int sbase = s * F;
            if (hf == 0) sSt[tk * ST2_LD + 24] = __uint_as_float(dca.key);     // read back by the key-role pass of the other tokens
#pragma unroll 1
            for (int hl = 0; hl < HH; ++hl) {
                const int h = hf * HH + hl;                // global head
                constexpr int JC = SAT_BWDA_JC;
                static_assert(JC % 2 == 0 && SAT_MAXF % JC == 0, "keys are processed in pairs");
                float sc[SAT_MAXF];
                float mx = -INFINITY;
                const float Dhh = sSt[tk * ST2_LD + 16 + h];
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
                sSt[tk * ST2_LD + h] = mx; sSt[tk * ST2_LD + 8 + h] = inv_sum;
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
                if (HH > 1) { rotate_regs<HD, DH>(q); rotate_regs<HD, DH>(dr); rotate_regs<HD, DH>(dq); }
            }
            store_half(a.dqkv + 0 * a.act_stride + tokpos, dq);
        }
        __syncthreads();                                   // dW_out phase is done with T0 / T1; partial blocks are visible
        if (tid < 64) {
            float cw[4][4];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) {
                float4 acc = *reinterpret_cast<const float4*>(sPart + tid * 16 + r4 * 4);
#pragma unroll
                for (int g4 = 1; g4 < 4; ++g4) {
                    const float4 o = *reinterpret_cast<const float4*>(sPart + g4 * 1024 + tid * 16 + r4 * 4);
                    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                }
                cw[r4][0] = acc.x; cw[r4][1] = acc.y; cw[r4][2] = acc.z; cw[r4][3] = acc.w;
            }
            wgrad_store(sdWO, D, tid, cw);
        }
        if (valid) { sts_half(sT0 + tk * KV_LD + ch0, q); sts_half(sT1 + tk * KV_LD + ch0, dr); }
        __syncthreads();
        // ---- key-role pass: dK'_j = sum_i dS_ij Q'_i / sqrt(dh) ; dV_j = sum_i Pd_ij dR_i (own heads) ------
        if (valid) {
            constexpr int IC = SAT_BWDA_IC;
            const int sbase = s * F;
            float* dk_out = a.dqkv + 1 * a.act_stride + tokpos - ch0;      // (row base: the head offset below is global)
            float* dv_out = a.dqkv + 2 * a.act_stride + tokpos - ch0;
            DropCtx dci = make_drop(seed32, 0u, l, SITE_ATT, a.p_drop, true);     // .key is replaced per query below
#pragma unroll 1
            for (int hl = 0; hl < HH; ++hl) {
                const int h = hf * HH + hl;
                float kh[DH], vh[DH], dk[DH], dv[DH];
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(sT2 + tk * KV_LD + h * DH + c);
                    const float4 v4 = *reinterpret_cast<const float4*>(sT3 + tk * KV_LD + h * DH + c);
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
                        const float* st = sSt + (sbase + i) * ST2_LD;
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
    store_partial<B2_NT>(part, sdWO, 1024, tid);
    store_partial_warps<B2_NT>(part + 1024, sdLN, 64, tid);
}

}  // namespace sat

using namespace sat;

#ifndef SAT_INST_H
#error "compile backward_attn2.cu with -DSAT_INST_H={2,4,8}"
#endif
#define SAT_CAT_(a, b) a##b
#define SAT_CAT(a, b) SAT_CAT_(a, b)

// a.tiles must be the 128-token tile set (SatWork.tiles2)
int SAT_CAT(sat_launch_bwd_attn2_h, SAT_INST_H)(const BwdArgs& a, int grid, cudaStream_t s) {
    constexpr int H = SAT_INST_H;
    const size_t smem = bwda2_smem_bytes(a.m.F);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_bwd_attn2<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_backward: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    k_bwd_attn2<H><<<grid, B2_NT, smem, s>>>(a);
    return sat_host::check_launch("k_bwd_attn2");
}
