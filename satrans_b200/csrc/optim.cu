// Optimiser + loss finalisation.
//
// Reference semantics being replaced: every step the reference
//   (a) adds l2 * sum(square(table)) over ALL rows of all tables to the loss  (meta_basemodel.py:577-593),
//   (b) materialises a dense (vocab, 32) gradient per table (embedding_dense_backward + 2*l2*w), and
//   (c) runs dense torch.optim.Adam over every parameter                       (main.py:343, meta_basemodel.py:328).
// Here no dense gradient exists.  Rows read by the batch ("touched") get  g = dE_accum + 2*l2*w  and Adam in
// k_adam_rows; with dense_mode=1 every other row gets  g = 2*l2*w  and Adam in k_adam_sweep, which streams
// p, m, v exactly once (6 x 128 B per row: the HBM roofline of the reference-semantics step) and also
// produces the sum of squares for (a).  The sweep touches no row the batch reads, so it can run on a second
// stream concurrently with forward/backward.
//
// Adam math follows torch.optim.Adam (amsgrad=False, weight_decay=0):
//   m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sat {

struct AdamConst {
    float one_minus_b1, b2, one_minus_b2, step_size, bc2_sqrt, eps, lr, two_l2;
    int kind;
};
// with a device-side step state (CUDA-graph replay) the first seven constants come from SatWork.step_state[SS_CONST..]
__device__ __forceinline__ AdamConst resolve_const(const AdamConst& by_value, const float* __restrict__ dev) {
    AdamConst c = by_value;
    if (dev) {
        c.one_minus_b1 = dev[0]; c.b2 = dev[1]; c.one_minus_b2 = dev[2]; c.step_size = dev[3]; c.bc2_sqrt = dev[4];
        c.eps = dev[5]; c.lr = dev[6];
    }
    return c;
}

// one parameter element; follows the op order of torch.optim's single-tensor implementations
__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, const AdamConst& c) {
    if (c.kind == SAT_OPT_ADAM) {
        m = fmaf(g - m, c.one_minus_b1, m);
        v = fmaf(c.one_minus_b2 * g, g, v * c.b2);
        const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
        p = p - c.step_size * (m / denom);
    } else if (c.kind == SAT_OPT_SGD) {
        p = p - c.lr * g;
    } else if (c.kind == SAT_OPT_ADAGRAD) {
        v = fmaf(g, g, v);
        p = p - c.lr * (g / (sqrtf(v) + c.eps));
    } else {    // RMSprop, alpha in b2
        v = fmaf(c.one_minus_b2 * g, g, v * c.b2);
        p = p - c.lr * (g / (sqrtf(v) + c.eps));
    }
}

// The same update for a row that no batch read (g = 2*l2*w, a smooth function of the row itself), used by BOTH table sweeps
// (every-step k_adam_sweep and the deferred k_defer_*), so that deferring is bit-neutral.  The multi-step sweep is bound by the
// SFU (MUFU issues at a quarter of the FMA rate), so the arithmetic is trimmed to 6 MUFU + ~40 FMA-pipe instructions per float4
// instead of 12 + ~130 for IEEE sqrt and two IEEE divisions per element: approximate square roots (sqrt.approx.ftz, <= 2 ulp),
// the bias correction as a multiplication by the reciprocal (computed once per step), ONE reciprocal per PAIR of denominators,
// and two channels per FMA-pipe instruction (FFMA2 / FMUL2 / FADD2).  The deviation from torch's op order is a few ulp of an update that is itself <= lr per step.
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Packed fp32 arithmetic (add2 / mul2 / fma2v of common.cuh: two IEEE operations per instruction, written as PTX so that the
// compiler can neither contract nor re-associate them -- it decides contraction per call site, and the every-step and the
// deferred sweep must round identically).  The two reciprocals serve the pair products of the four denominators
// (1/a = b / (a b); every denominator is >= eps, so the product is far from underflow for eps >= 1e-18).
__device__ __forceinline__ void adam4_untouched(float4& p, float4& m, float4& v, const AdamConst& c, float inv_bc2_sqrt) {
    float4 g;
    mul2(g.x, g.y, p.x, p.y, c.two_l2, c.two_l2); mul2(g.z, g.w, p.z, p.w, c.two_l2, c.two_l2);
    if (c.kind == SAT_OPT_SGD) {
        fma2v(p.x, p.y, g.x, g.y, -c.lr, -c.lr); fma2v(p.z, p.w, g.z, g.w, -c.lr, -c.lr);
        return;
    }
    float4 num = g;
    float scale = c.lr, dscale = 1.f;
    if (c.kind == SAT_OPT_ADAM) {            // m += (g - m) * (1 - b1)
        float4 t;
        add2(t.x, t.y, g.x, g.y, -m.x, -m.y); add2(t.z, t.w, g.z, g.w, -m.z, -m.w);
        fma2v(m.x, m.y, t.x, t.y, c.one_minus_b1, c.one_minus_b1); fma2v(m.z, m.w, t.z, t.w, c.one_minus_b1, c.one_minus_b1);
        num = m; scale = c.step_size; dscale = inv_bc2_sqrt;
    }
    if (c.kind == SAT_OPT_ADAGRAD) {         // v += g * g
        fma2v(v.x, v.y, g.x, g.y, g.x, g.y); fma2v(v.z, v.w, g.z, g.w, g.z, g.w);
    } else {                                 // Adam, RMSprop (alpha in b2): v = v * b2 + ((1 - b2) * g) * g
        float4 a;
        mul2(a.x, a.y, g.x, g.y, c.one_minus_b2, c.one_minus_b2); mul2(a.z, a.w, g.z, g.w, c.one_minus_b2, c.one_minus_b2);
        mul2(v.x, v.y, v.x, v.y, c.b2, c.b2); mul2(v.z, v.w, v.z, v.w, c.b2, c.b2);
        fma2v(v.x, v.y, a.x, a.y, g.x, g.y); fma2v(v.z, v.w, a.z, a.w, g.z, g.w);
    }
    // denominators sqrt(v) * dscale + eps, their reciprocals from the two pair products
    float4 d = make_float4(c.eps, c.eps, c.eps, c.eps);
    fma2v(d.x, d.y, sqrt_approx(v.x), sqrt_approx(v.y), dscale, dscale);
    fma2v(d.z, d.w, sqrt_approx(v.z), sqrt_approx(v.w), dscale, dscale);
    float ab, cd;
    mul2(ab, cd, d.x, d.z, d.y, d.w);
    const float rab = rcp_approx(ab), rcd = rcp_approx(cd);
    float4 r;
    mul2(r.x, r.y, rab, rab, d.y, d.x); mul2(r.z, r.w, rcd, rcd, d.w, d.z);
    mul2(r.x, r.y, r.x, r.y, num.x, num.y); mul2(r.z, r.w, r.z, r.w, num.z, num.w);
    fma2v(p.x, p.y, r.x, r.y, -scale, -scale); fma2v(p.z, p.w, r.z, r.w, -scale, -scale);      // p -= scale * (num / denom)
}

__device__ __forceinline__ void adam4(float4& p, float4& m, float4& v, const float4& g, const AdamConst& c) {
    adam1(p.x, m.x, v.x, g.x, c); adam1(p.y, m.y, v.y, g.y, c);
    adam1(p.z, m.z, v.z, g.z, c); adam1(p.w, m.w, v.w, g.w, c);
}

__device__ __forceinline__ void block_sum_to(double val, double* __restrict__ dst) {
    __shared__ double s_red[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if (lane == 0) s_red[wid] = val;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (blockDim.x >> 5) ? s_red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) *dst = t;
    }
}

constexpr int SWEEP_THREADS = 256;
constexpr int SWEEP_UNROLL = 4;
constexpr int SWEEP_BLOCKS = 148 * 256;   // slots of reg_part reserved for the sweep: SAT_REG_PARTS = SWEEP_BLOCKS + ROWS_BLOCKS
// CTAs the sweep is launched with.  Many short-lived CTAs instead of one resident wave of grid-stride loops: the short launch
// chain that updates the touched rows / dense parameters runs beside the sweep on a high-priority stream and gets SM slots as
// sweep CTAs retire (with 148 x 8 persistent CTAs it had to wait for the whole sweep).  SAT_SWEEP_BPS overrides (tuning aid).
static int sweep_blocks() {
    static int n = 0;
    if (!n) {
        const char* e = getenv("SAT_SWEEP_BPS");
        int bps = e ? atoi(e) : 128;
        if (bps < 1) bps = 1;
        if (bps > 256) bps = 256;
        n = 148 * bps;
    }
    return n;
}

// Dense sweep over all table rows NOT touched by the batch.  One float4 per thread per access; p/m/v are
// read once and written once.  reg_part[blockIdx] receives sum(p_old^2) over the rows this block swept.
__global__ void __launch_bounds__(SWEEP_THREADS)
k_adam_sweep(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, const uint32_t* __restrict__ bitmap,
             long long n_vec /* total_rows * 8 */, const AdamConst c_in, const float* __restrict__ dev_const,
             double* __restrict__ reg_part) {
    const AdamConst c = resolve_const(c_in, dev_const);
    const float inv_bc2_sqrt = 1.f / c.bc2_sqrt;
    float sq = 0.f;
    const long long stride = (long long)gridDim.x * SWEEP_THREADS;
    for (long long i0 = (long long)blockIdx.x * SWEEP_THREADS + threadIdx.x; i0 < n_vec; i0 += stride * SWEEP_UNROLL) {
        float4 p[SWEEP_UNROLL], m[SWEEP_UNROLL], v[SWEEP_UNROLL];
        bool act[SWEEP_UNROLL];
#pragma unroll
        for (int u = 0; u < SWEEP_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            act[u] = false;
            if (i < n_vec) {
                const long long row = i >> 3;
                act[u] = !((__ldg(bitmap + (row >> 5)) >> (row & 31)) & 1u);
            }
            if (act[u]) {
                p[u] = __ldcs(P + i);
                if (c.kind == SAT_OPT_ADAM) m[u] = __ldcs(M + i);
                if (c.kind != SAT_OPT_SGD) v[u] = __ldcs(V + i);
            }
        }
#pragma unroll
        for (int u = 0; u < SWEEP_UNROLL; ++u) {
            if (act[u]) {
                const long long i = i0 + u * stride;
                sq = fmaf(p[u].x, p[u].x, sq); sq = fmaf(p[u].y, p[u].y, sq);
                sq = fmaf(p[u].z, p[u].z, sq); sq = fmaf(p[u].w, p[u].w, sq);
                adam4_untouched(p[u], m[u], v[u], c, inv_bc2_sqrt);
                __stcs(P + i, p[u]);
                if (c.kind == SAT_OPT_ADAM) __stcs(M + i, m[u]);
                if (c.kind != SAT_OPT_SGD) __stcs(V + i, v[u]);
            }
        }
    }
    block_sum_to((double)sq, reg_part + blockIdx.x);
}

constexpr int ROWS_THREADS = 256;
constexpr int ROWS_BLOCKS = 148 * 4;

// Touched rows: g = accumulated dE + 2*l2*p, Adam, clear the accumulator for the next step.
// Deferred mode (upto != nullptr): the rows are marked as up to date through this step (index n - 1 of the sweep epoch, n =
// defer_state[0]) and the sum-of-squares partials go to this step's slots of the deferred buffer.
__global__ void __launch_bounds__(ROWS_THREADS)
k_adam_rows(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, const int32_t* __restrict__ uniq_rows,
            float4* __restrict__ row_grad, const int32_t* __restrict__ counters, const AdamConst c_in,
            const float* __restrict__ dev_const, double* __restrict__ reg_part, uint8_t* __restrict__ upto,
            const float* __restrict__ defer_state) {
    const AdamConst c = resolve_const(c_in, dev_const);
    int n_begun = 0;
    if (upto) {
        n_begun = reinterpret_cast<const int*>(defer_state)[0];
        reg_part += (long long)(n_begun - 1) * ROWS_BLOCKS;
    }
    const long long n_vec = (long long)counters[CNT_NUNIQ] * 8;
    float sq = 0.f;
    for (long long i = (long long)blockIdx.x * ROWS_THREADS + threadIdx.x; i < n_vec;
         i += (long long)gridDim.x * ROWS_THREADS) {
        const long long slot = i >> 3;
        const long long e = (long long)__ldg(uniq_rows + slot) * 8 + (i & 7);
        float4 p = P[e], m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
        if (c.kind == SAT_OPT_ADAM) m = M[e];
        if (c.kind != SAT_OPT_SGD) v = V[e];
        float4 g = row_grad[i];                  // (fully rewritten by sat_rows_reduce / sat_rows_combine every step)
        sq = fmaf(p.x, p.x, sq); sq = fmaf(p.y, p.y, sq); sq = fmaf(p.z, p.z, sq); sq = fmaf(p.w, p.w, sq);
        g.x = fmaf(c.two_l2, p.x, g.x); g.y = fmaf(c.two_l2, p.y, g.y);
        g.z = fmaf(c.two_l2, p.z, g.z); g.w = fmaf(c.two_l2, p.w, g.w);
        adam4(p, m, v, g, c);
        P[e] = p;
        if (c.kind == SAT_OPT_ADAM) M[e] = m;
        if (c.kind != SAT_OPT_SGD) V[e] = v;
        if (upto && (i & 7) == 0) upto[e >> 3] = (uint8_t)n_begun;
    }
    block_sum_to((double)sq, reg_part + blockIdx.x);
}

// ---- deferred table sweep (include/satrans_b200.h: sat_defer_*) -----------------------------------------------------------
constexpr int DEFER_MAX = SAT_DEFER_MAX;
constexpr int DEFER_RING = 16;            // words per ring entry of SatAdam.defer_state
static_assert(SAT_DEFER_STATE_WORDS == 8 + DEFER_RING * DEFER_MAX, "layout of SatAdam.defer_state");
constexpr int DF_RES = 0, DF_TOUCH = 32, DF_CATCH = DF_TOUCH + DEFER_MAX * SAT_DEFER_ROWS_BLOCKS,
              DF_SWEEP = DF_CATCH + DEFER_MAX * SAT_DEFER_ROWS_BLOCKS * DEFER_MAX;
static_assert(SAT_DEFER_ROWS_BLOCKS == ROWS_BLOCKS, "the touched-row kernel writes one partial per CTA");
static_assert(DF_SWEEP + SAT_DEFER_SWEEP_BLOCKS * DEFER_MAX == SAT_DEFER_REG_DOUBLES, "layout of SatAdam.defer_reg");

// constants of the epoch's step j (st: the block's shared-memory copy of SatAdam.defer_state); inv = 1 / bc2_sqrt
__device__ __forceinline__ AdamConst ring_const(const float* __restrict__ st, int j, int kind, float& inv) {
    const float4* r = reinterpret_cast<const float4*>(st + 8 + DEFER_RING * j);
    const float4 a = r[0], b = r[1];
    AdamConst c;
    c.one_minus_b1 = a.x; c.b2 = a.y; c.one_minus_b2 = a.z; c.step_size = a.w; c.bc2_sqrt = b.x; c.eps = b.y; c.lr = b.z;
    c.two_l2 = b.w; c.kind = kind;
    inv = st[8 + DEFER_RING * j + 8];
    return c;
}
// the block's copy of the epoch state
__device__ __forceinline__ void load_defer_state(float* __restrict__ s_st, const float* __restrict__ st) {
    for (int i = threadIdx.x; i < SAT_DEFER_STATE_WORDS; i += blockDim.x) s_st[i] = st[i];
    __syncthreads();
}

__global__ void k_defer_begin(float* __restrict__ st, const AdamConst c_in, const float* __restrict__ dev_const) {
    const AdamConst c = resolve_const(c_in, dev_const);
    int* sti = reinterpret_cast<int*>(st);
    const int j = sti[0];
    if (j >= DEFER_MAX) { sti[1] = 1; return; }            // the caller did not sweep in time: reported by sat_defer_sweep
    float* r = st + 8 + DEFER_RING * j;
    r[0] = c.one_minus_b1; r[1] = c.b2; r[2] = c.one_minus_b2; r[3] = c.step_size; r[4] = c.bc2_sqrt; r[5] = c.eps; r[6] = c.lr;
    r[7] = c.two_l2; r[8] = 1.f / c.bc2_sqrt;           // (the same IEEE quotient k_adam_sweep forms once per launch)
    sti[0] = j + 1;
}

// One float4 of a row through the steps [j0, j1) of the epoch as a row that no batch read (g = 2*l2*w): exactly what
// k_adam_sweep does to it once per step.  The sum of squares each step saw goes to this thread's column of s_sq[step].
__device__ __forceinline__ void defer_steps(float4& p, float4& m, float4& v, int j0, int j1, const float* __restrict__ st,
                                            int kind, float (*s_sq)[ROWS_THREADS]) {
    for (int j = j0; j < j1; ++j) {
        float inv;
        const AdamConst c = ring_const(st, j, kind, inv);
        float sq = p.x * p.x;
        sq = fmaf(p.y, p.y, sq); sq = fmaf(p.z, p.z, sq); sq = fmaf(p.w, p.w, sq);
        s_sq[j][threadIdx.x] += sq;
        adam4_untouched(p, m, v, c, inv);
    }
}

// per-step block sums of s_sq -> dst[j], j < DEFER_MAX (every slot is written: no memset between epochs)
__device__ __forceinline__ void defer_block_sums(float (*s_sq)[ROWS_THREADS], double* __restrict__ dst) {
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = wid; j < DEFER_MAX; j += ROWS_THREADS / 32) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ROWS_THREADS / 32; ++k) t += (double)s_sq[j][lane + 32 * k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) dst[j] = t;
    }
}

// The rows this batch reads (sorted unique list) receive the steps of the epoch they have not seen: after this kernel they
// hold exactly the values a sweep after every step would have left, and the forward may gather them.
__global__ void __launch_bounds__(ROWS_THREADS)
k_defer_catchup(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, uint8_t* __restrict__ upto,
                const int32_t* __restrict__ uniq_rows, const int32_t* __restrict__ counters, const float* __restrict__ st,
                int kind, double* __restrict__ reg_catch) {
    __shared__ float s_sq[DEFER_MAX][ROWS_THREADS];
    __shared__ __align__(16) float s_st[SAT_DEFER_STATE_WORDS];
#pragma unroll
    for (int j = 0; j < DEFER_MAX; ++j) s_sq[j][threadIdx.x] = 0.f;
    load_defer_state(s_st, st);
    st = s_st;
    const int cur = reinterpret_cast<const int*>(st)[0] - 1;           // index of the step now running
    const long long n_vec = (long long)counters[CNT_NUNIQ] * 8;
    if (cur > 0) {
        // (warp-uniform trip count: n_vec is a multiple of 8, the warps' first elements of 32)
        const int lane = threadIdx.x & 31;
        for (long long b0 = (long long)blockIdx.x * ROWS_THREADS + (threadIdx.x & ~31); b0 < n_vec; b0 += (long long)gridDim.x * ROWS_THREADS) {
            const long long i = b0 + lane;
            long long row = 0;
            int j0 = cur;
            if (i < n_vec) { row = __ldg(uniq_rows + (i >> 3)); j0 = upto[row]; }
            __syncwarp();                        // the 8 lanes of a row have read upto[row] before one of them rewrites it
            if (j0 < cur) {
                const long long e = row * 8 + (i & 7);
                float4 p = P[e], m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
                if (kind == SAT_OPT_ADAM) m = M[e];
                if (kind != SAT_OPT_SGD) v = V[e];
                defer_steps(p, m, v, j0, cur, st, kind, s_sq);
                P[e] = p;
                if (kind == SAT_OPT_ADAM) M[e] = m;
                if (kind != SAT_OPT_SGD) V[e] = v;
                if ((i & 7) == 0) upto[row] = (uint8_t)cur;
            }
        }
    }
    defer_block_sums(s_sq, reg_catch + ((long long)max(cur, 0) * ROWS_BLOCKS + blockIdx.x) * DEFER_MAX);
}

// End of the epoch: every row receives the steps it has not seen, upto = 0.
#ifndef SAT_DSWEEP_U
#define SAT_DSWEEP_U 1
#endif
#ifndef SAT_DSWEEP_MINB
#define SAT_DSWEEP_MINB 6
#endif
__global__ void __launch_bounds__(ROWS_THREADS, SAT_DSWEEP_MINB)
k_defer_sweep(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, uint8_t* __restrict__ upto,
              long long n_vec, const float* __restrict__ st, int kind, double* __restrict__ reg_sweep) {
    __shared__ float s_sq[DEFER_MAX][ROWS_THREADS];
    __shared__ __align__(16) float s_st[SAT_DEFER_STATE_WORDS];
#pragma unroll
    for (int j = 0; j < DEFER_MAX; ++j) s_sq[j][threadIdx.x] = 0.f;
    load_defer_state(s_st, st);
    st = s_st;
    const int n = min(reinterpret_cast<const int*>(st)[0], DEFER_MAX);
    constexpr int U = SAT_DSWEEP_U;
    const long long stride = (long long)gridDim.x * ROWS_THREADS;
    const int lane = threadIdx.x & 31;
    // (warp-uniform trip count: n_vec is a multiple of 8, the warps' first elements of 32)
    for (long long b0 = (long long)blockIdx.x * ROWS_THREADS + (threadIdx.x & ~31); b0 < n_vec; b0 += stride * U) {
        float4 p[U], m[U], v[U];
        int j0[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = b0 + lane + u * stride;
            j0[u] = -1;                                                  // past the end
            if (i < n_vec) j0[u] = upto[i >> 3];
            if (j0[u] >= 0 && j0[u] < n) {
                p[u] = __ldcs(P + i);
                if (kind == SAT_OPT_ADAM) m[u] = __ldcs(M + i);
                if (kind != SAT_OPT_SGD) v[u] = __ldcs(V + i);
            }
        }
        __syncwarp();                            // the 8 lanes of a row have read upto[row] before one of them clears it
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = b0 + lane + u * stride;
            if (j0[u] >= 0 && j0[u] < n) {
                defer_steps(p[u], m[u], v[u], j0[u], n, st, kind, s_sq);
                __stcs(P + i, p[u]);
                if (kind == SAT_OPT_ADAM) __stcs(M + i, m[u]);
                if (kind != SAT_OPT_SGD) __stcs(V + i, v[u]);
            }
            if (j0[u] > 0 && (i & 7) == 0) upto[i >> 3] = 0;
        }
    }
    defer_block_sums(s_sq, reg_sweep + (long long)blockIdx.x * DEFER_MAX);
}

// Per-step regulariser terms of the closed epoch, in a fixed order: touched rows of the step, then the catch-up launches, then
// the sweep.  res[j] = l2 * sum, res[DEFER_MAX] += their total, res[DEFER_MAX + 1] = n; the epoch counter returns to 0.
__global__ void __launch_bounds__(256)
k_defer_finalize(float* __restrict__ st, double* __restrict__ reg, float l2, int sweep_blocks) {
    __shared__ double s_out;
    int* sti = reinterpret_cast<int*>(st);
    const int n = min(sti[0], DEFER_MAX);
    double total = 0.0;
    for (int j = 0; j < n; ++j) {
        double a = 0.0;
        for (int b = threadIdx.x; b < ROWS_BLOCKS; b += 256) a += reg[DF_TOUCH + (long long)j * ROWS_BLOCKS + b];
        for (int l = j + 1; l < n; ++l)
            for (int b = threadIdx.x; b < ROWS_BLOCKS; b += 256) a += reg[DF_CATCH + ((long long)l * ROWS_BLOCKS + b) * DEFER_MAX + j];
        for (int b = threadIdx.x; b < sweep_blocks; b += 256) a += reg[DF_SWEEP + (long long)b * DEFER_MAX + j];
        block_sum_to(a, &s_out);
        __syncthreads();
        const double r = (double)l2 * s_out;
        total += r;
        if (threadIdx.x == 0) reg[DF_RES + j] = r;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        reg[DF_RES + DEFER_MAX] += total;
        reg[DF_RES + DEFER_MAX + 1] = (double)n;
        reg[DF_RES + DEFER_MAX + 2] = (double)sti[1];       // != 0: more than SAT_DEFER_MAX steps were begun without a sweep
        sti[0] = 0;
    }
}

// Flat dense-parameter buffer (148 k floats in the EN configuration).
__global__ void __launch_bounds__(256)
k_adam_dense(float* __restrict__ P, float* __restrict__ M, float* __restrict__ V, const float* __restrict__ G, long long n,
             const AdamConst c_in, const float* __restrict__ dev_const) {
    const AdamConst c = resolve_const(c_in, dev_const);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float p = P[i], m = 0.f, v = 0.f;
        if (c.kind == SAT_OPT_ADAM) m = M[i];
        if (c.kind != SAT_OPT_SGD) v = V[i];
        adam1(p, m, v, G[i], c);
        P[i] = p;
        if (c.kind == SAT_OPT_ADAM) M[i] = m;
        if (c.kind != SAT_OPT_SGD) V[i] = v;
    }
}

// scalars[0] = sum of per-tile BCE partials; scalars[1] = l2 * (sum of squares partials); scalars[3] = error flags
__global__ void __launch_bounds__(256)
k_finalize(const double* __restrict__ loss_part, const int32_t* __restrict__ counters, const double* __restrict__ reg_part,
           int n_reg_part, float l2, float* __restrict__ scalars, int tile_set) {
    __shared__ double s_out[2];
    double a = 0.0, r = 0.0;
    const int n_tiles = counters[tile_set ? CNT_NTILES2 : CNT_NTILES];
    for (int i = threadIdx.x; i < n_tiles; i += 256) a += loss_part[i];
    for (int i = threadIdx.x; i < n_reg_part; i += 256) r += reg_part[i];
    block_sum_to(a, &s_out[0]);
    __syncthreads();
    block_sum_to(r, &s_out[1]);
    __syncthreads();
    if (threadIdx.x == 0) {
        scalars[0] = (float)s_out[0];
        scalars[1] = (float)((double)l2 * s_out[1]);
        scalars[3] = (float)counters[CNT_ERR];
    }
}

// sum of squares of the whole table (used by the lazy mode / evaluation of the regulariser only)
__global__ void __launch_bounds__(SWEEP_THREADS)
k_sumsq(const float4* __restrict__ P, long long n_vec, double* __restrict__ reg_part) {
    float sq = 0.f;
    for (long long i = (long long)blockIdx.x * SWEEP_THREADS + threadIdx.x; i < n_vec; i += (long long)gridDim.x * SWEEP_THREADS) {
        const float4 p = __ldcs(P + i);
        sq = fmaf(p.x, p.x, sq); sq = fmaf(p.y, p.y, sq); sq = fmaf(p.z, p.z, sq); sq = fmaf(p.w, p.w, sq);
    }
    block_sum_to((double)sq, reg_part + blockIdx.x);
}

}  // namespace sat

using namespace sat;

static AdamConst make_const(const SatAdam* a, bool with_l2) {
    AdamConst c;
    const double bc1 = 1.0 - pow((double)a->beta1, (double)a->step);
    const double bc2 = 1.0 - pow((double)a->beta2, (double)a->step);
    c.one_minus_b1 = (float)(1.0 - (double)a->beta1);
    c.b2 = a->beta2;
    c.one_minus_b2 = (float)(1.0 - (double)a->beta2);
    c.step_size = (float)((double)a->lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = a->eps;
    c.two_l2 = with_l2 ? 2.f * a->l2_emb : 0.f;
    c.lr = a->lr;
    c.kind = a->kind;
    return c;
}

// reg_part layout: [0, SWEEP_BLOCKS) sweep partials, [SWEEP_BLOCKS, SWEEP_BLOCKS + ROWS_BLOCKS) touched-row partials
extern "C" int sat_tables_adam(const SatModel* m, SatWork* w, const SatAdam* a, int32_t which, void* stream) {
    if (!m || !w || !a || !w->reg_part || (a->kind == SAT_OPT_ADAM && !a->tab_m) || (a->kind != SAT_OPT_SGD && !a->tab_v) ||
        a->kind < 0 || a->kind > SAT_OPT_RMSPROP)
        return sat_host::fail(SAT_EINVAL, "sat_tables_adam: null argument / bad optimiser kind");
    if (a->step < 1) return sat_host::fail(SAT_EINVAL, "sat_tables_adam: step must be >= 1");
    const AdamConst c = make_const(a, true);
    const float* dev_const = w->device_step ? w->step_state + SS_CONST : nullptr;
    const long long n_vec = m->local_rows * 8;
    cudaStream_t s = (cudaStream_t)stream;
    if (which & SAT_ADAM_TOUCHED) {
        const bool deferred = which & SAT_ADAM_DEFERRED;
        if (deferred && (!a->row_upto || !a->defer_state || !a->defer_reg || !a->dense_mode))
            return sat_host::fail(SAT_EINVAL, "sat_tables_adam: SAT_ADAM_DEFERRED needs dense_mode and the deferred-sweep buffers");
        k_adam_rows<<<ROWS_BLOCKS, ROWS_THREADS, 0, s>>>(
            reinterpret_cast<float4*>(m->tables), reinterpret_cast<float4*>(a->tab_m), reinterpret_cast<float4*>(a->tab_v),
            w->uniq_rows, reinterpret_cast<float4*>(w->row_grad), w->counters, c, dev_const,
            deferred ? a->defer_reg + DF_TOUCH : w->reg_part + SWEEP_BLOCKS, deferred ? a->row_upto : nullptr, a->defer_state);
        int rc = sat_host::check_launch("k_adam_rows");
        if (rc) return rc;
    }
    if ((which & SAT_ADAM_UNTOUCHED) && !a->dense_mode) {
        // lazy mode: no sweep runs, so its slice of the sum-of-squares partials must not keep an earlier step's values; the
        // regulariser term reported by sat_finalize_loss then covers the touched rows only
        if (cudaMemsetAsync(w->reg_part, 0, SWEEP_BLOCKS * sizeof(double), s) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_tables_adam: memset failed");
    }
    if ((which & SAT_ADAM_UNTOUCHED) && a->dense_mode) {
        if (cudaMemsetAsync(w->reg_part, 0, SWEEP_BLOCKS * sizeof(double), s) != cudaSuccess)       // (slots of CTAs not launched)
            return sat_host::fail(SAT_ECUDA, "sat_tables_adam: memset failed");
        k_adam_sweep<<<sweep_blocks(), SWEEP_THREADS, 0, s>>>(
            reinterpret_cast<float4*>(m->tables), reinterpret_cast<float4*>(a->tab_m), reinterpret_cast<float4*>(a->tab_v),
            w->bitmap, n_vec, c, dev_const, w->reg_part);
        int rc = sat_host::check_launch("k_adam_sweep");
        if (rc) return rc;
    }
    return SAT_OK;
}

static int defer_args_ok(const SatModel* m, const SatAdam* a, const char* who) {
    if (!m || !a || !a->row_upto || !a->defer_state || !a->defer_reg || (a->kind == SAT_OPT_ADAM && !a->tab_m) ||
        (a->kind != SAT_OPT_SGD && !a->tab_v) || a->kind < 0 || a->kind > SAT_OPT_RMSPROP)
        return sat_host::fail(SAT_EINVAL, "%s: null argument / bad optimiser kind", who);
    return SAT_OK;
}

extern "C" int sat_defer_begin(const SatModel* m, SatWork* w, const SatAdam* a, void* stream) {
    int rc = defer_args_ok(m, a, "sat_defer_begin");
    if (rc) return rc;
    if (!w) return sat_host::fail(SAT_EINVAL, "sat_defer_begin: null workspace");
    if (a->step < 1) return sat_host::fail(SAT_EINVAL, "sat_defer_begin: step must be >= 1");
    const AdamConst c = make_const(a, true);
    k_defer_begin<<<1, 1, 0, (cudaStream_t)stream>>>(a->defer_state, c, w->device_step ? w->step_state + SS_CONST : nullptr);
    return sat_host::check_launch("k_defer_begin");
}

extern "C" int sat_defer_catchup(const SatModel* m, SatWork* w, const SatAdam* a, void* stream) {
    int rc = defer_args_ok(m, a, "sat_defer_catchup");
    if (rc) return rc;
    if (!w || !w->uniq_rows || !w->counters) return sat_host::fail(SAT_EINVAL, "sat_defer_catchup: null workspace");
    k_defer_catchup<<<ROWS_BLOCKS, ROWS_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(m->tables), reinterpret_cast<float4*>(a->tab_m), reinterpret_cast<float4*>(a->tab_v), a->row_upto,
        w->uniq_rows, w->counters, a->defer_state, a->kind, a->defer_reg + DF_CATCH);
    return sat_host::check_launch("k_defer_catchup");
}

extern "C" int sat_defer_sweep(const SatModel* m, const SatAdam* a, void* stream) {
    int rc = defer_args_ok(m, a, "sat_defer_sweep");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    k_defer_sweep<<<SAT_DEFER_SWEEP_BLOCKS, ROWS_THREADS, 0, s>>>(
        reinterpret_cast<float4*>(m->tables), reinterpret_cast<float4*>(a->tab_m), reinterpret_cast<float4*>(a->tab_v), a->row_upto,
        m->local_rows * 8, a->defer_state, a->kind, a->defer_reg + DF_SWEEP);
    rc = sat_host::check_launch("k_defer_sweep");
    if (rc) return rc;
    k_defer_finalize<<<1, 256, 0, s>>>(a->defer_state, a->defer_reg, a->l2_emb, SAT_DEFER_SWEEP_BLOCKS);
    return sat_host::check_launch("k_defer_finalize");
}

extern "C" int sat_tables_sumsq(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->reg_part) return sat_host::fail(SAT_EINVAL, "sat_tables_sumsq: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(w->reg_part + SWEEP_BLOCKS, 0, ROWS_BLOCKS * sizeof(double), s) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_tables_sumsq: memset failed");
    if (cudaMemsetAsync(w->reg_part, 0, SWEEP_BLOCKS * sizeof(double), s) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_tables_sumsq: memset failed");
    k_sumsq<<<148 * 8, SWEEP_THREADS, 0, s>>>(reinterpret_cast<const float4*>(m->tables), m->local_rows * 8, w->reg_part);
    return sat_host::check_launch("k_sumsq");
}

extern "C" int sat_dense_adam(const SatModel* m, SatWork* w, const SatAdam* a, void* stream) {
    if (!m || !w || !a || !w->dgrad || (a->kind == SAT_OPT_ADAM && !a->dense_m) || (a->kind != SAT_OPT_SGD && !a->dense_v))
        return sat_host::fail(SAT_EINVAL, "sat_dense_adam: null argument");
    if (a->step < 1) return sat_host::fail(SAT_EINVAL, "sat_dense_adam: step must be >= 1");
    const AdamConst c = make_const(a, false);
    int blocks = (int)((m->n_dense_params + 255) / 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    const float* dev_const = w->device_step ? w->step_state + SS_CONST : nullptr;
    k_adam_dense<<<blocks, 256, 0, (cudaStream_t)stream>>>(m->dense, a->dense_m, a->dense_v, w->dgrad, m->n_dense_params, c, dev_const);
    return sat_host::check_launch("k_adam_dense");
}

// ---- device-side step state -------------------------------------------------------------------------------------------
__global__ void k_step_begin(float* __restrict__ st, unsigned long long seed_base) {
    int* sti = reinterpret_cast<int*>(st);
    const int step = sti[SS_STEP] + 1;
    sti[SS_STEP] = step;
    *reinterpret_cast<unsigned long long*>(st + SS_SEED) = seed_base + (unsigned long long)step;
    const float lr = st[SS_HYPER + 0], b1 = st[SS_HYPER + 1], b2 = st[SS_HYPER + 2], eps = st[SS_HYPER + 3];
    const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
    float* c = st + SS_CONST;
    c[0] = (float)(1.0 - (double)b1); c[1] = b2; c[2] = (float)(1.0 - (double)b2);
    c[3] = (float)((double)lr / bc1); c[4] = (float)sqrt(bc2); c[5] = eps; c[6] = lr;
}
__global__ void k_step_set(float* __restrict__ st, int step, float lr, float b1, float b2, float eps) {
    reinterpret_cast<int*>(st)[SS_STEP] = step;
    st[SS_HYPER + 0] = lr; st[SS_HYPER + 1] = b1; st[SS_HYPER + 2] = b2; st[SS_HYPER + 3] = eps;
}
extern "C" int sat_step_begin(SatWork* w, uint64_t seed_base, void* stream) {
    if (!w || !w->step_state) return sat_host::fail(SAT_EINVAL, "sat_step_begin: null argument");
    k_step_begin<<<1, 1, 0, (cudaStream_t)stream>>>(w->step_state, seed_base);
    return sat_host::check_launch("k_step_begin");
}
extern "C" int sat_step_set(SatWork* w, int32_t step, float lr, float beta1, float beta2, float eps, void* stream) {
    if (!w || !w->step_state || step < 0) return sat_host::fail(SAT_EINVAL, "sat_step_set: bad argument");
    k_step_set<<<1, 1, 0, (cudaStream_t)stream>>>(w->step_state, step, lr, beta1, beta2, eps);
    return sat_host::check_launch("k_step_set");
}

extern "C" int sat_finalize_loss(const SatModel* m, SatWork* w, float l2_emb, int32_t tile_set, void* stream) {
    if (!m || !w || !w->scalars) return sat_host::fail(SAT_EINVAL, "sat_finalize_loss: null argument");
    k_finalize<<<1, 256, 0, (cudaStream_t)stream>>>(w->loss_part, w->counters, w->reg_part, SWEEP_BLOCKS + ROWS_BLOCKS, l2_emb,
                                                    w->scalars, tile_set);
    return sat_host::check_launch("k_finalize");
}
