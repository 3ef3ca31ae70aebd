// Device code shared by the CUDA-core forward (forward.cu) and the tcgen05 forward (forward_tc.cu).
#pragma once
#include "args.cuh"

#ifndef SAT_FWD_JC
#define SAT_FWD_JC 8           // keys per basic block in attention_fwd (8: forward 0.42 -> 0.41 ms at F = 19 vs 4; 2: 0.48)
#endif

namespace sat {

// Multi-head attention of one query token against the F tokens of its sample (reference satrans.py:75-90).
// K'/V rows of the sample start at row `sbase` of the [token][36] tiles.
// Rolled over heads: q is rotated by DH per head so that the head's channels are always q[0..DH) (compact code).
template <int H>
__device__ __forceinline__ void attention_fwd(float (&q)[D], const float* __restrict__ sK,
                                              const float* __restrict__ sV, int sbase, int F, const DropCtx& dc,
                                              float (&r)[D]) {
    constexpr int DH = D / H;
    const float sm_scale = 1.f / sqrtf((float)DH);       // scores * (1/sqrt(dh)); differs from the reference's division by <= 1 ulp
    // Keys are processed in chunks of JC inside ONE basic block (a per-key `if (j < F)` made every key its own block and
    // serialised the dot-product / hash / exp chains: "wait" was 41 % of the forward's stall samples).  Keys past F in the
    // last chunk read row F-1 and get score -inf: probability exactly 0, so the sums are unchanged bit for bit.
    constexpr int JC = SAT_FWD_JC;
    static_assert(JC % 2 == 0 && SAT_MAXF % JC == 0, "keys are processed in pairs");
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
        float sc[SAT_MAXF];
        float mx = -INFINITY;
#pragma unroll
        for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
            if (j0 < F) {
#pragma unroll
                for (int j = j0; j < j0 + JC; ++j) {
                    const float* kr = sK + (sbase + min(j, F - 1)) * KV_LD + h * DH;
                    float acc = 0.f, acc1 = 0.f;             // even / odd channels (one FFMA2 per channel pair)
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 kk = *reinterpret_cast<const float4*>(kr + c);
                        fma2v(acc, acc1, q[c + 0], q[c + 1], kk.x, kk.y);
                        fma2v(acc, acc1, q[c + 2], q[c + 3], kk.z, kk.w);
                    }
                    acc += acc1;
                    sc[j] = j < F ? acc * sm_scale : -INFINITY;     // inner_product /= dh ** 0.5   (satrans.py:85-86)
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
        float o[DH];
        zero(o);
        const float inv_sum = 1.f / sum;
#pragma unroll
        for (int j0 = 0; j0 < SAT_MAXF; j0 += JC) {
            if (j0 < F) {
                float fd[JC];
#pragma unroll
                for (int j = j0; j < j0 + JC; j += 2) drop_pair(dc, (h * SAT_MAXF + j) >> 1, fd[j - j0], fd[j - j0 + 1]);
#pragma unroll
                for (int j = j0; j < j0 + JC; ++j) {
                    const float pj = (sc[j] * inv_sum) * fd[j - j0];                        // softmax, attn_dropout
                    const float* vr = sV + (sbase + min(j, F - 1)) * KV_LD + h * DH;
#pragma unroll
                    for (int c = 0; c < DH; c += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(vr + c);
                        fma2(o[c + 0], o[c + 1], pj, vv.x, vv.y);
                        fma2(o[c + 2], o[c + 3], pj, vv.z, vv.w);
                    }
                }
            }
        }
        // r[h*DH + c] = o[c]: write the head's slot, then rotate so that after H heads r is in channel order
#pragma unroll
        for (int c = 0; c < DH; ++c) r[c] = o[c];
        if (H > 1) { rotate_regs<D, DH>(r); rotate_regs<D, DH>(q); }
    }
}

}  // namespace sat
