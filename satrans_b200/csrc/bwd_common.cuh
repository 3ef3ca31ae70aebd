// Device helpers shared by the backward kernels.
#pragma once
#include "args.cuh"

namespace sat {

// ---- warp reduce-scatter: on return lane L holds sum over the 32 lanes of v[L] (in v[0]) -------------
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[D], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
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

// LayerNorm backward for one token.  xhat = normalised input, g = gamma (smem).  dy is overwritten by dx.
// dgam/dbet (smem, 32 floats each, PRIVATE TO THE CALLING WARP) receive the token-summed parameter gradients: lane L owns
// element L, so the update is a plain read-modify-write and the summation order (tiles in order, then warps in order at the
// flush) does not depend on timing.
__device__ __forceinline__ void layernorm_bwd32(float (&dy)[D], const float (&xhat)[D], float rstd,
                                                const float* __restrict__ g, float* __restrict__ dgam,
                                                float* __restrict__ dbet, int lane, bool valid) {
    float t1[D], t2[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { t1[i] = valid ? dy[i] * xhat[i] : 0.f; t2[i] = valid ? dy[i] : 0.f; }
    const float sg = warp_reduce_scatter32(t1, lane);
    const float sb = warp_reduce_scatter32(t2, lane);
    dgam[lane] += sg;
    dbet[lane] += sb;
    float m1a = 0.f, m1b = 0.f, m2a = 0.f, m2b = 0.f;
#pragma unroll
    for (int i = 0; i < D; i += 2) {
        dy[i] *= g[i]; dy[i + 1] *= g[i + 1];
        m1a += dy[i]; m1b += dy[i + 1];
        m2a = fmaf(dy[i], xhat[i], m2a); m2b = fmaf(dy[i + 1], xhat[i + 1], m2b);
    }
    const float m1 = (m1a + m1b) * (1.f / D), m2 = (m2a + m2b) * (1.f / D);
#pragma unroll
    for (int i = 0; i < D; ++i) dy[i] = rstd * (dy[i] - m1 - xhat[i] * m2);
}

// normalise in place, return rstd (no affine)
__device__ __forceinline__ float normalize32(float (&v)[D]) {
    const float mean = sum32(v) * (1.f / D);
    const float q = sumsq_dev32(v, mean);              // same reduction order as layernorm32 (the forward)
#pragma unroll
    for (int i = 0; i < D; i += 2) add2(v[i], v[i + 1], v[i], v[i + 1], -mean, -mean);
    const float rstd = 1.f / sqrtf(q * (1.f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < D; i += 2) mul2(v[i], v[i + 1], v[i], v[i + 1], rstd, rstd);
    return rstd;
}

// acc[(ia+r)*ldacc + jg+c] += sum_{t<ntok} A[t][ia+r] * G[t][jg+c]  -- 32x32 product, A/G tiles [tok][36].
template <int NTHREADS = NT>
__device__ __forceinline__ void wgrad32(const float* __restrict__ A, const float* __restrict__ G, int ntok,
                                        float* __restrict__ acc, int ldacc, int tid) {
    const int blk = tid & 63, grp = tid >> 6;
    const int ia = (blk >> 3) * 4, jg = (blk & 7) * 4;
    float c[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) c[r][cc] = 0.f;
#pragma unroll 4
    for (int t = grp; t < ntok; t += NTHREADS / 64) {
        const float4 a = *reinterpret_cast<const float4*>(A + t * KV_LD + ia);
        const float4 g = *reinterpret_cast<const float4*>(G + t * KV_LD + jg);
        c[0][0] = fmaf(a.x, g.x, c[0][0]); c[0][1] = fmaf(a.x, g.y, c[0][1]); c[0][2] = fmaf(a.x, g.z, c[0][2]); c[0][3] = fmaf(a.x, g.w, c[0][3]);
        c[1][0] = fmaf(a.y, g.x, c[1][0]); c[1][1] = fmaf(a.y, g.y, c[1][1]); c[1][2] = fmaf(a.y, g.z, c[1][2]); c[1][3] = fmaf(a.y, g.w, c[1][3]);
        c[2][0] = fmaf(a.z, g.x, c[2][0]); c[2][1] = fmaf(a.z, g.y, c[2][1]); c[2][2] = fmaf(a.z, g.z, c[2][2]); c[2][3] = fmaf(a.z, g.w, c[2][3]);
        c[3][0] = fmaf(a.w, g.x, c[3][0]); c[3][1] = fmaf(a.w, g.y, c[3][1]); c[3][2] = fmaf(a.w, g.z, c[3][2]); c[3][3] = fmaf(a.w, g.w, c[3][3]);
    }
    // the NTHREADS / 64 token groups add their blocks one after the other (fixed order, no atomics); every thread of the
    // CTA calls wgrad32, so the barriers are uniform
    for (int gsel = 0; gsel < NTHREADS / 64; ++gsel) {
        if (grp == gsel) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float4* p = reinterpret_cast<float4*>(acc + (ia + r) * ldacc + jg);
                float4 v = *p;
                v.x += c[r][0]; v.y += c[r][1]; v.z += c[r][2]; v.w += c[r][3];
                *p = v;
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void fma_block(float (&c)[4][4], const float4& a, const float4& g) {
    fma2(c[0][0], c[0][1], a.x, g.x, g.y); fma2(c[0][2], c[0][3], a.x, g.z, g.w);
    fma2(c[1][0], c[1][1], a.y, g.x, g.y); fma2(c[1][2], c[1][3], a.y, g.z, g.w);
    fma2(c[2][0], c[2][1], a.z, g.x, g.y); fma2(c[2][2], c[2][3], a.z, g.z, g.w);
    fma2(c[3][0], c[3][1], a.w, g.x, g.y); fma2(c[3][2], c[3][3], a.w, g.z, g.w);
}

// c[r][cc] = sum_{t in [tb, te)} A[t][ia + r] * G[t][jg + cc] for this thread's 4x4 block (blk = tid & 63) of a 32x32 product.
// te - tb is a multiple of 4 (the tiles' rows past the last token are zero); tokens go four at a time with the next four
// already requested from shared memory (software pipeline: the FMAs of one group hide the LDS latency of the next).
__device__ __forceinline__ void wgrad_block(const float* __restrict__ A, const float* __restrict__ G, int tb, int te, int tid,
                                            float (&c)[4][4]) {
    const int blk = tid & 63;
    const float* pa = A + (blk >> 3) * 4 + tb * KV_LD;
    const float* pg = G + (blk & 7) * 4 + tb * KV_LD;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) c[r][cc] = 0.f;
    if (te <= tb) return;
    float4 a[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = *reinterpret_cast<const float4*>(pa + k * KV_LD);
        g[k] = *reinterpret_cast<const float4*>(pg + k * KV_LD);
    }
#pragma unroll 1
    for (int t = tb + 4; t < te; t += 4) {
        pa += 4 * KV_LD; pg += 4 * KV_LD;
        float4 na[4], ng[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            na[k] = *reinterpret_cast<const float4*>(pa + k * KV_LD);
            ng[k] = *reinterpret_cast<const float4*>(pg + k * KV_LD);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) fma_block(c, a[k], g[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = na[k]; g[k] = ng[k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) fma_block(c, a[k], g[k]);
}

// acc[(ia + r) * ldacc + jg + cc] += c[r][cc] with plain 16 B read-modify-writes: the caller guarantees one owner per element
__device__ __forceinline__ void wgrad_store(float* __restrict__ acc, int ldacc, int tid, const float (&c)[4][4]) {
    const int blk = tid & 63;
    const int ia = (blk >> 3) * 4, jg = (blk & 7) * 4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float4* p = reinterpret_cast<float4*>(acc + (ia + r) * ldacc + jg);
        float4 v = *p;
        v.x += c[r][0]; v.y += c[r][1]; v.z += c[r][2]; v.w += c[r][3];
        *p = v;
    }
}

// Two 32x32 products at once in a 128-thread CTA: the caller hands each 64-thread half (tid >> 6, warp-uniform) its own
// A / G / acc.  Every thread owns one 4x4 block over ALL tokens, so the shared-memory accumulator is updated without
// atomics (fp32 atomicAdd on shared memory is a CAS loop: 17 % of k_bwd_meta_tc's stall samples before).
__device__ __forceinline__ void wgrad_half(const float* __restrict__ A, const float* __restrict__ G, int ntok4,
                                           float* __restrict__ acc, int ldacc, int tid) {
    float c[4][4];
    wgrad_block(A, G, 0, ntok4, tid, c);
    wgrad_store(acc, ldacc, tid, c);
}

// This CTA's partial sums -> its row of a per-CTA partial buffer (plain stores; k_dense_reduce adds the rows in CTA order).
template <int NTHREADS = NT>
__device__ __forceinline__ void store_partial(float* __restrict__ dst, const float* __restrict__ sacc, int n, int tid) {
    for (int i = tid * 4; i < n; i += NTHREADS * 4)
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(sacc + i);
}
// the same for the per-warp LayerNorm slots [NW][n]: warps added in order
template <int NTHREADS = NT>
__device__ __forceinline__ void store_partial_warps(float* __restrict__ dst, const float* __restrict__ sacc, int n, int tid) {
    for (int i = tid; i < n; i += NTHREADS) {
        float v = sacc[i];
#pragma unroll
        for (int w = 1; w < NTHREADS / 32; ++w) v += sacc[w * n + i];
        dst[i] = v;
    }
}

// ---- per-CTA partial gradient buffers (SatWork.gpart) ---------------------------------------------------------------------
constexpr int GP_ATTN = 1088;      // dW_out [1024] | d layer_norm gamma, beta [64]
constexpr int GP_META = 3200;      // dW_Q, dW_K, dW_V [3072] | d MetaNet-LN (Q gamma, beta, K gamma, beta) [128]
struct GPart {
    long long attn, meta, theta, head, total;   // float offsets
    int nside, head_ld;
};
__host__ __device__ __forceinline__ GPart gpart_layout(const SatModel& m) {
    GPart g;
    g.nside = (m.flags & SAT_FLAG_POS) ? 2 : 1;
    g.head_ld = (m.F * D + m.n_dense + 1 + 3) & ~3;
    g.attn = 0;
    g.meta = g.attn + (long long)m.L * SAT_GRID_MAX * GP_ATTN;
    g.theta = g.meta + (long long)m.L * SAT_GRID_MAX * GP_META;
    g.head = g.theta + (long long)m.L * SAT_GRID_MAX * g.nside * PSZ;
    g.total = g.head + (long long)SAT_GRID_MAX * g.head_ld;
    return g;
}

// ---- CTA -> tile range with every CTA inside ONE scenario ------------------------------------------------------------------
// The MetaNet backward accumulates d theta of one scenario per CTA, and the reduction over CTAs must not depend on timing, so
// the CTAs are dealt to the scenarios in proportion to their tile counts (every non-empty scenario gets at least one) and a
// scenario's tiles are split evenly over its CTAs.  hist = samples per scenario (counters[CNT_HIST..]), spt = samples per
// tile.  c_sel >= 0 asks for the range of that CTA; g_sel >= 0 asks for the CTA range [c0, c1) of that scenario.
__device__ __forceinline__ void scenario_partition(const int32_t* __restrict__ hist, int n_dom, int spt, int G, int c_sel,
                                                   int g_sel, int& t0, int& t1, int& g_of, int& c0, int& c1) {
    long long T = 0;
    int K = 0;
    for (int g = 0; g < n_dom; ++g) {
        const int nt = (hist[g] + spt - 1) / spt;
        T += nt; K += nt > 0;
    }
    const int extra = G - K;
    int cbase = 0, tbase = 0;
    t0 = t1 = 0; g_of = -1; c0 = c1 = 0;
    for (int g = 0; g < n_dom; ++g) {
        const int nt = (hist[g] + spt - 1) / spt;
        if (nt == 0) continue;
        int c = 1 + (int)((long long)extra * nt / T);
        if (c > nt) c = nt;
        if (c_sel >= cbase && c_sel < cbase + c) {
            const int j = c_sel - cbase;
            t0 = tbase + (int)((long long)nt * j / c);
            t1 = tbase + (int)((long long)nt * (j + 1) / c);
            g_of = g;
        }
        if (g == g_sel) { c0 = cbase; c1 = cbase + c; }
        cbase += c; tbase += nt;
    }
}


}  // namespace sat
