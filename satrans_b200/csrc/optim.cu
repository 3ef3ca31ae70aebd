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
                const float4 g = make_float4(c.two_l2 * p[u].x, c.two_l2 * p[u].y, c.two_l2 * p[u].z, c.two_l2 * p[u].w);
                adam4(p[u], m[u], v[u], g, c);
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
__global__ void __launch_bounds__(ROWS_THREADS)
k_adam_rows(float4* __restrict__ P, float4* __restrict__ M, float4* __restrict__ V, const int32_t* __restrict__ uniq_rows,
            float4* __restrict__ row_grad, const int32_t* __restrict__ counters, const AdamConst c_in,
            const float* __restrict__ dev_const, double* __restrict__ reg_part) {
    const AdamConst c = resolve_const(c_in, dev_const);
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
    }
    block_sum_to((double)sq, reg_part + blockIdx.x);
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
        k_adam_rows<<<ROWS_BLOCKS, ROWS_THREADS, 0, s>>>(
            reinterpret_cast<float4*>(m->tables), reinterpret_cast<float4*>(a->tab_m), reinterpret_cast<float4*>(a->tab_v),
            w->uniq_rows, reinterpret_cast<float4*>(w->row_grad), w->counters, c, dev_const, w->reg_part + SWEEP_BLOCKS);
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
