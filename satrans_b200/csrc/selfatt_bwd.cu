// Backward of SelfAttention_Layer (reference submodules.py:208-238 under torch autograd) and of deepctr-torch's InteractingLayer
// (the same graph without scaling / dropout / LayerNorm) -- "next" row 1 of SURVEY.md section 8f.
//
//   forward:  Q,K,V = X W_{Q,K,V};  P = softmax(Q K^T / sqrt(dh));  R = drop(P) V;  Z = drop(R) + X W_Res;  Y = LN(relu(Z))
//   backward: recompute the forward per tile (nothing but X is saved), then
//       dA = LN'(dY) (+ d gamma, d beta);  dZ = dA * [Z > 0];  dW_Res += X^T dZ;  dX = dZ W_Res^T;  dR = drop(dZ)
//       attention backward with the FlashAttention identity sum_j P_ij dP_ij = dR_i . R_i (no F x F tensor is stored):
//       query-role pass -> dQ, key-role pass -> dK, dV;  dW_{Q,K,V} += X^T d{Q,K,V};  dX += dQ W_Q^T + dK W_K^T + dV W_V^T
// Mapping as everywhere: one thread = one field token, a CTA = floor(128/F) whole samples, 128 threads.  Weight gradients:
// operand tiles [token][36] in shared memory, 4x4 register blocks, token groups added in a fixed order; every CTA stores its
// sums into its row of the caller's scratch buffer and k_selfatt_bwd_reduce adds the rows in CTA order (no float atomics).
#include "bwd_common.cuh"
#include "fwd_common.cuh"

namespace sat {

constexpr int SB_NT = 128;
constexpr int SB_PART = 4096 + 64;             // dW_Q | dW_K | dW_V | dW_Res | d gamma | d beta
constexpr int SB_ST_LD = 25;                   // per token: row max [8], 1 / row sum [8], D = dR.R [8], attention-dropout key

struct SelfAttBwdArgs {
    const float* x;        // [B][F][32]
    const float* dy;       // [B][F][32]
    float* dx;             // [B][F][32]
    const float* wq; const float* wk; const float* wv; const float* wres;
    const float* ln_g;     // NULL: no LayerNorm (InteractingLayer)
    const int32_t* idx;    // NULL: samples 0..B-1; else the B sample indices this launch works on (one scenario of an
                           // MDR_InteractingLayer): rows of x / dy / dx and the dropout counters are those of sample idx[i]
    float* part;           // [gridDim.x][SB_PART]
    int B, F, use_res, training, scaling;
    float p_drop;
    unsigned long long seed;
};

static size_t sb_smem_bytes(int F) {
    const int rows = ((SB_NT / F) * F + 3) & ~3;
    return (size_t)(4096 + 4096 + 64 + (SB_NT / 32) * 64 + 5 * rows * KV_LD + SB_NT * SB_ST_LD) * sizeof(float);
}

template <int H>
__global__ void __launch_bounds__(SB_NT, 1) k_selfatt_bwd(const SelfAttBwdArgs a) {
    constexpr int DH = D / H;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, F = a.F;
    const int rows = ((SB_NT / F) * F + 3) & ~3;
    float* sW = smem;                          // W_Query | W_Key | W_Value | W_Res
    float* sdW = sW + 4096;                    // their gradients
    float* sLN = sdW + 4096;                   // gamma | beta
    float* sdLN = sLN + 64;                    // per warp: d gamma | d beta
    float* TX = sdLN + (SB_NT / 32) * 64;      // x                      (A operand of every weight gradient)
    float* TA = TX + rows * KV_LD;             // dZ -> q -> dq, dk, dv  (G operands / key-role pass)
    float* TB = TA + rows * KV_LD;             // dR
    float* TK = TB + rows * KV_LD;             // k
    float* TV = TK + rows * KV_LD;             // v
    float* sSt = TV + rows * KV_LD;            // [128][25]

    copy_to_smem(sW, a.wq, 1024, tid, SB_NT);
    copy_to_smem(sW + 1024, a.wk, 1024, tid, SB_NT);
    copy_to_smem(sW + 2048, a.wv, 1024, tid, SB_NT);
    if (a.use_res) copy_to_smem(sW + 3072, a.wres, 1024, tid, SB_NT);
    for (int i = tid; i < 4096; i += SB_NT) sdW[i] = 0.f;
    for (int i = tid; i < (SB_NT / 32) * 64; i += SB_NT) sdLN[i] = 0.f;
    if (tid < 32) { sLN[tid] = a.ln_g ? a.ln_g[tid] : 1.f; sLN[32 + tid] = 0.f; }      // (beta does not enter the backward)
    const int spt = SB_NT / F;
    const int s = tid / F, f = tid - s * F;
    const int n_tiles = (a.B + spt - 1) / spt;
    const int t0 = (int)((long long)n_tiles * blockIdx.x / gridDim.x), t1 = (int)((long long)n_tiles * (blockIdx.x + 1) / gridDim.x);
    const float inv_sqrt_dh = 1.f / sqrtf((float)DH);
    const float qscale = a.scaling ? 1.f : sqrtf((float)DH);        // scores = (q * qscale) . k / sqrt(dh)
    const uint32_t seed32 = fold_seed(a.seed);
    const bool stage = tid < rows;
    float* sdLNw = sdLN + (tid >> 5) * 64;

    for (int t = t0; t < t1; ++t) {
        const int cnt = min(spt, a.B - t * spt);
        const bool valid = s < cnt;
        const int b = !valid ? 0 : a.idx ? a.idx[t * spt + s] : t * spt + s;
        const int ntok = cnt * F;
        const long long tok = ((long long)b * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        const int sbase = s * F;
        __syncthreads();                                   // weights staged / previous tile done with the tiles
        float x[D], q[D];
        zero(x); zero(q);
        {
            float k[D], v[D];
            zero(k); zero(v);
            if (valid) {
                load_row(a.x + tok, x);
                vm_outer<D, D, D>(sW, x, q);
                vm_outer<D, D, D>(sW + 1024, x, k);
                vm_outer<D, D, D>(sW + 2048, x, v);
                if (!a.scaling) {
#pragma unroll
                    for (int i = 0; i < D; ++i) q[i] *= qscale;
                }
            }
            if (stage) { sts_row(TX + tid * KV_LD, x); sts_row(TK + tid * KV_LD, k); sts_row(TV + tid * KV_LD, v); }
        }
        __syncthreads();
        // ---- forward recompute, LayerNorm / ReLU / residual backward ------------------------------------------------------------
        float dr[D], dxa[D];
        zero(dr); zero(dxa);
        {
            float dz[D], z[D], r[D];
            zero(dz); zero(z); zero(r);
            uint32_t zmask = 0u;
            float rstd = 0.f;
            const DropCtx dco = make_drop(seed32, token_gid, 0, SITE_OUT, a.p_drop, a.training);
            if (valid) {
                attention_fwd<H>(q, TK, TV, sbase, F, make_drop(seed32, token_gid, 0, SITE_ATT, a.p_drop, a.training), r);
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float f0, f1;
                    drop_pair(dco, i >> 1, f0, f1);
                    z[i] = r[i] * f0; z[i + 1] = r[i + 1] * f1;
                }
                if (a.use_res) vm_outer<D, D, D>(sW + 3072, x, z);
#pragma unroll
                for (int i = 0; i < D; ++i) { if (z[i] > 0.f) zmask |= 1u << i; z[i] = fmaxf(z[i], 0.f); }
                load_row(a.dy + tok, dz);
                if (a.ln_g) rstd = normalize32(z);                 // z := normalised relu(Z)
            }
            // (warp shuffles inside: every lane of the warp takes part, lanes without a token contribute zeros)
            if (a.ln_g) layernorm_bwd32(dz, z, rstd, sLN, sdLNw, sdLNw + 32, lane, valid);          // dz := dLoss/d relu(Z)
            if (valid) {
#pragma unroll
                for (int i = 0; i < D; ++i) dz[i] = ((zmask >> i) & 1u) ? dz[i] : 0.f;              // dz := dLoss/dZ
                if (a.use_res) vm_dot<D, D, D>(sW + 3072, dz, dxa);                                 // dX = dZ W_Res^T
#pragma unroll
                for (int i = 0; i < D; i += 2) {
                    float f0, f1;
                    drop_pair(dco, i >> 1, f0, f1);
                    dr[i] = dz[i] * f0; dr[i + 1] = dz[i + 1] * f1;
                }
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float acc = 0.f;
#pragma unroll
                    for (int c = 0; c < DH; ++c) acc = fmaf(dr[h * DH + c], r[h * DH + c], acc);
                    sSt[tid * SB_ST_LD + 16 + h] = acc;            // D_i[h] = dR_i . R_i
                }
            } else {
                zero(dz);
            }
            if (stage) sts_row(TA + tid * KV_LD, dz);
        }
        __syncthreads();
        if (a.use_res) wgrad32<SB_NT>(TX, TA, ntok, sdW + 3072, D, tid);       // dW_Res += X^T dZ   (ends with a barrier)
        // ---- query-role pass: dQ_i = sum_j dS_ij K_j / sqrt(dh) ----------------------------------------------------------------
        float dq[D];
        zero(dq);
        if (valid) {
            const DropCtx dca = make_drop(seed32, token_gid, 0, SITE_ATT, a.p_drop, a.training);
            sSt[tid * SB_ST_LD + 24] = __uint_as_float(dca.key);
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float sc[SAT_MAXF];
                float mx = -INFINITY;
                const float Dhh = sSt[tid * SB_ST_LD + 16 + h];
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += 4) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + 4; ++j) {
                            const float* kr = TK + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            float acc = 0.f, acc1 = 0.f;     // even / odd channels, same order as attention_fwd
#pragma unroll
                            for (int c = 0; c < DH; c += 4) {
                                const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                                fma2v(acc, acc1, q[c + 0], q[c + 1], kk.x, kk.y);
                                fma2v(acc, acc1, q[c + 2], q[c + 3], kk.z, kk.w);
                            }
                            acc += acc1;
                            sc[j] = j < F ? acc * inv_sqrt_dh : -INFINITY;
                            mx = fmaxf(mx, sc[j]);
                        }
                    }
                }
                float sum = 0.f;
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += 4) {
                    if (j0 < F) {
#pragma unroll
                        for (int j = j0; j < j0 + 4; j += 2) { softmax_exp_pair(sc[j], sc[j + 1], mx); sum += sc[j]; sum += sc[j + 1]; }
                    }
                }
                const float inv_sum = 1.f / sum;
                sSt[tid * SB_ST_LD + h] = mx; sSt[tid * SB_ST_LD + 8 + h] = inv_sum;
#pragma unroll
                for (int j0 = 0; j0 < SAT_MAXF; j0 += 4) {
                    if (j0 < F) {
                        float fd[4];
#pragma unroll
                        for (int j = j0; j < j0 + 4; j += 2) drop_pair(dca, (h * SAT_MAXF + j) >> 1, fd[j - j0], fd[j - j0 + 1]);
#pragma unroll
                        for (int j = j0; j < j0 + 4; ++j) {
                            const float pj = sc[j] * inv_sum;                  // exactly 0 for the padding keys
                            const float* vr = TV + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                            const float* kr = TK + (sbase + min(j, F - 1)) * KV_LD + h * DH;
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
        }
        if (stage) { sts_row(TA + tid * KV_LD, q); sts_row(TB + tid * KV_LD, dr); }       // (dW_Res is done with TA)
        __syncthreads();
        // ---- key-role pass: dK_j = sum_i dS_ij Q_i / sqrt(dh);  dV_j = sum_i drop(P)_ij dR_i ----------------------------------------
        float dk[D], dv[D];
        zero(dk); zero(dv);
        if (valid) {
            DropCtx dci = make_drop(seed32, 0u, 0, SITE_ATT, a.p_drop, a.training);      // .key is replaced per query below
#pragma unroll 1
            for (int h = 0; h < H; ++h) {
                float kh[DH], vh[DH], dkh[DH], dvh[DH];
#pragma unroll
                for (int c = 0; c < DH; c += 4) {
                    const float4 k4 = *reinterpret_cast<const float4*>(TK + tid * KV_LD + h * DH + c);
                    const float4 v4 = *reinterpret_cast<const float4*>(TV + tid * KV_LD + h * DH + c);
                    kh[c] = k4.x; kh[c + 1] = k4.y; kh[c + 2] = k4.z; kh[c + 3] = k4.w;
                    vh[c] = v4.x; vh[c + 1] = v4.y; vh[c + 2] = v4.z; vh[c + 3] = v4.w;
                }
                zero(dkh); zero(dvh);
                const uint32_t e = h * SAT_MAXF + f;
#pragma unroll 1
                for (int i0 = 0; i0 < F; i0 += 4) {
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        const int i = min(i0 + ii, F - 1);
                        const float* qr = TA + (sbase + i) * KV_LD + h * DH;
                        const float* dR = TB + (sbase + i) * KV_LD + h * DH;
                        const float* st = sSt + (sbase + i) * SB_ST_LD;
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
                        const float pij = (i0 + ii < F) ? softmax_exp(acc * inv_sqrt_dh - st[h]) * st[8 + h] : 0.f;
                        dci.key = __float_as_uint(st[24]);
                        const float df = drop_factor(dci, e);
                        const float dsc = pij * (dAd * df - st[16 + h]) * inv_sqrt_dh;
                        const float pd = pij * df;
#pragma unroll
                        for (int c = 0; c < DH; c += 2) {
                            fma2(dkh[c], dkh[c + 1], dsc, qq[c], qq[c + 1]);
                            fma2(dvh[c], dvh[c + 1], pd, dd[c], dd[c + 1]);
                        }
                    }
                }
                // head slice -> slot 0, rotate: after H heads dk / dv are in channel order
#pragma unroll
                for (int c = 0; c < DH; ++c) { dk[c] = dkh[c]; dv[c] = dvh[c]; }
                if (H > 1) { rotate_regs<D, DH>(dk); rotate_regs<D, DH>(dv); }
            }
        }
        __syncthreads();                                   // the key-role pass is done with TA / TB
        // ---- projection weight gradients and dX ---------------------------------------------------------------------------------
        if (!a.scaling) {
#pragma unroll
            for (int i = 0; i < D; ++i) dq[i] *= qscale;     // q entered the scores scaled by qscale
        }
        if (stage) sts_row(TA + tid * KV_LD, dq);
        __syncthreads();
        wgrad32<SB_NT>(TX, TA, ntok, sdW, D, tid);
        if (stage) sts_row(TA + tid * KV_LD, dk);
        __syncthreads();
        wgrad32<SB_NT>(TX, TA, ntok, sdW + 1024, D, tid);
        if (stage) sts_row(TA + tid * KV_LD, dv);
        __syncthreads();
        wgrad32<SB_NT>(TX, TA, ntok, sdW + 2048, D, tid);
        if (valid) {
            vm_dot<D, D, D>(sW, dq, dxa);                    // dX[i] += sum_j dQ[j] W_Q[i][j]
            vm_dot<D, D, D>(sW + 1024, dk, dxa);
            vm_dot<D, D, D>(sW + 2048, dv, dxa);
            store_row(a.dx + tok, dxa);
        }
    }
    __syncthreads();
    float* part = a.part + (long long)blockIdx.x * SB_PART;
    store_partial<SB_NT>(part, sdW, 4096, tid);
    store_partial_warps<SB_NT>(part + 4096, sdLN, 64, tid);
}

// rows of the scratch buffer added in CTA order: out[0:4096] -> dW_Q | dW_K | dW_V | dW_Res, out[4096:4160] -> d gamma | d beta
__global__ void __launch_bounds__(256) k_selfatt_bwd_reduce(const float* __restrict__ part, int n_rows, float* __restrict__ dwq,
                                                            float* __restrict__ dwk, float* __restrict__ dwv, float* __restrict__ dwres,
                                                            float* __restrict__ dg, float* __restrict__ db) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= SB_PART) return;
    float acc = 0.f;
    for (int r = 0; r < n_rows; ++r) acc += part[(long long)r * SB_PART + e];
    float* out = e < 1024 ? dwq + e : e < 2048 ? dwk + (e - 1024) : e < 3072 ? dwv + (e - 2048) : e < 4096 ? (dwres ? dwres + (e - 3072) : nullptr)
                 : e < 4128 ? (dg ? dg + (e - 4096) : nullptr) : (db ? db + (e - 4128) : nullptr);
    if (out) *out = acc;
}

}  // namespace sat

using namespace sat;

static int sab_grid(int B, int F) {
    const int spt = SB_NT / F;
    const int n_tiles = (B + spt - 1) / spt;
    return n_tiles < 148 ? n_tiles : 148;
}

extern "C" int64_t sat_field_attention_bwd_scratch(int32_t B, int32_t F) {
    if (B <= 0 || F <= 0 || F > SAT_MAXF) return 0;
    return (int64_t)sab_grid(B, F) * SB_PART * sizeof(float);
}

extern "C" int sat_field_attention_bwd(const float* x, const float* dy, float* dx, const float* wq, const float* wk, const float* wv,
                                       const float* wres, const float* ln_g, const int32_t* sample_idx, int32_t B, int32_t F,
                                       int32_t H, int32_t use_res,
                                       int32_t scaling, int32_t training, float dropout_p, uint64_t seed, float* dwq, float* dwk,
                                       float* dwv, float* dwres, float* dln_g, float* dln_b, float* scratch, void* stream) {
    if (!x || !dy || !dx || !wq || !wk || !wv || !dwq || !dwk || !dwv || !scratch || (use_res && (!wres || !dwres)) ||
        (ln_g && (!dln_g || !dln_b)))
        return sat_host::fail(SAT_EINVAL, "sat_field_attention_bwd: null argument");
    if (F <= 0 || F > SAT_MAXF) return sat_host::fail(SAT_EINVAL, "sat_field_attention_bwd: F=%d not in 1..%d", F, SAT_MAXF);
    if (dropout_p < 0.f || dropout_p >= 1.f) return sat_host::fail(SAT_EINVAL, "sat_field_attention_bwd: dropout_p=%f", dropout_p);
    if (B <= 0) return SAT_OK;
    SelfAttBwdArgs a;
    a.x = x; a.dy = dy; a.dx = dx; a.wq = wq; a.wk = wk; a.wv = wv; a.wres = wres; a.ln_g = ln_g; a.idx = sample_idx; a.part = scratch;
    a.B = B; a.F = F; a.use_res = use_res; a.training = training; a.scaling = scaling; a.p_drop = dropout_p; a.seed = seed;
    const int grid = sab_grid(B, F);
    const size_t smem = sb_smem_bytes(F);
    cudaStream_t s = (cudaStream_t)stream;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_selfatt_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_bwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_bwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_field_attention_bwd: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    switch (H) {
        case 1: k_selfatt_bwd<1><<<grid, SB_NT, smem, s>>>(a); break;
        case 2: k_selfatt_bwd<2><<<grid, SB_NT, smem, s>>>(a); break;
        case 4: k_selfatt_bwd<4><<<grid, SB_NT, smem, s>>>(a); break;
        case 8: k_selfatt_bwd<8><<<grid, SB_NT, smem, s>>>(a); break;
        default: return sat_host::fail(SAT_EINVAL, "sat_field_attention_bwd: heads=%d not in {1,2,4,8}", H);
    }
    int rc = sat_host::check_launch("k_selfatt_bwd");
    if (rc) return rc;
    k_selfatt_bwd_reduce<<<(SB_PART + 255) / 256, 256, 0, s>>>(scratch, grid, dwq, dwk, dwv, use_res ? dwres : nullptr,
                                                               ln_g ? dln_g : nullptr, ln_g ? dln_b : nullptr);
    return sat_host::check_launch("k_selfatt_bwd_reduce");
}
