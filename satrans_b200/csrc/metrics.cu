// Per-batch training metrics of fit(verbose > 0): the reference evaluates sklearn's log_loss / roc_auc_score /
// mean_squared_error / accuracy_score on every training batch on the host (meta_basemodel.py:330-337, 352-356, metric
// table :655-671) -- 94 % of the wall time of fit(verbose=1) once the step itself runs on the GPU.  Here one CTA per batch
// computes the same four numbers from the predictions that are already in HBM:
//   logloss : mean of -(y log(clip(p)) + (1-y) log(clip(1-p))) in float64, clip to [eps, 1-eps], eps = 2^-52 (sklearn's
//             float64 'auto' eps); NaN for a single-class batch (sklearn raises, the host path records NaN)
//   auc     : Mann-Whitney U with average ranks for tied scores == roc_auc_score's trapezoid; scores are sorted in shared
//             memory (bitonic sort of 64-bit keys: order-preserving float bits << 32 | label)
//   mse, accuracy (p > 0.5)
#include <math.h>

#include "common.cuh"

namespace sat {

constexpr int MET_THREADS = 1024;
constexpr int MET_MAX = 16384;            // samples per batch sortable in shared memory (128 KB of keys)

__device__ __forceinline__ double block_sum_double(double v, double* s_red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < MET_THREADS / 32; ++i) t += s_red[i];      // same order in every thread
    return t;
}

__global__ void __launch_bounds__(MET_THREADS) k_batch_metrics(const float* __restrict__ y_pred, const float* __restrict__ y,
                                                              const int32_t* __restrict__ off, double* __restrict__ out) {
    extern __shared__ unsigned long long keys[];
    __shared__ double s_red[MET_THREADS / 32];
    const int tid = threadIdx.x, b = blockIdx.x;
    const int s0 = off[b], n = off[b + 1] - s0;
    int n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    const double eps = 2.220446049250313e-16;
    double ll = 0.0, se = 0.0, correct = 0.0, npos = 0.0;
    for (int i = tid; i < n_pad; i += MET_THREADS) {
        unsigned long long key = ~0ull;                               // padding sorts to the end
        if (i < n) {
            const float p = y_pred[s0 + i], t = y[s0 + i];
            uint32_t u = __float_as_uint(p);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);           // order-preserving bits
            const int lab = t == 1.f;
            key = ((unsigned long long)u << 32) | (unsigned long long)lab;
            const double pd = (double)p, td = (double)lab;
            const double a = fmin(fmax(pd, eps), 1.0 - eps), c = fmin(fmax(1.0 - pd, eps), 1.0 - eps);
            ll -= td * log(a) + (1.0 - td) * log(c);
            se += ((double)t - pd) * ((double)t - pd);
            correct += (t == ((p > 0.5f) ? 1.f : 0.f)) ? 1.0 : 0.0;
            npos += td;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n_pad; i += MET_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > c) == asc) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // rank sum of the positives; tied scores share the average of their 1-based ranks
    double rs = 0.0;
    for (int i = tid; i < n; i += MET_THREADS) {
        const unsigned long long key = keys[i];
        if (key & 1ull) {
            const uint32_t u = (uint32_t)(key >> 32);
            int lo = 0, hi = n;                                       // first index with score >= u
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((uint32_t)(keys[mid] >> 32) < u) lo = mid + 1; else hi = mid; }
            const int first = lo;
            hi = n;                                                   // first index with score > u
            while (lo < hi) { const int mid = (lo + hi) >> 1; if ((uint32_t)(keys[mid] >> 32) <= u) lo = mid + 1; else hi = mid; }
            rs += 0.5 * (double)(first + (lo - 1)) + 1.0;
        }
    }
    ll = block_sum_double(ll, s_red);
    se = block_sum_double(se, s_red);
    correct = block_sum_double(correct, s_red);
    npos = block_sum_double(npos, s_red);
    rs = block_sum_double(rs, s_red);
    if (tid == 0) {
        const double nn = (double)n, nneg = nn - npos;
        const bool one_class = npos == 0.0 || nneg == 0.0;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        out[4 * b + 0] = one_class ? nan : ll / nn;
        out[4 * b + 1] = one_class ? nan : (rs - npos * (npos + 1.0) * 0.5) / (npos * nneg);
        out[4 * b + 2] = se / nn;
        out[4 * b + 3] = correct / nn;
    }
}

}  // namespace sat

using namespace sat;

extern "C" int sat_batch_metrics(const float* y_pred, const float* y, const int32_t* batch_off, int32_t n_batches,
                                 int32_t max_batch, double* out, void* stream) {
    if (!y_pred || !y || !batch_off || !out || n_batches < 0)
        return sat_host::fail(SAT_EINVAL, "sat_batch_metrics: null argument");
    if (n_batches == 0) return SAT_OK;
    if (max_batch < 1 || max_batch > MET_MAX)
        return sat_host::fail(SAT_EINVAL, "sat_batch_metrics: batches of 1..%d samples are supported, got %d", MET_MAX, max_batch);
    int n_pad = 1;
    while (n_pad < max_batch) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(unsigned long long);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_batch_metrics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_batch_metrics: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    k_batch_metrics<<<n_batches, MET_THREADS, smem, (cudaStream_t)stream>>>(y_pred, y, batch_off, out);
    return sat_host::check_launch("k_batch_metrics");
}
