// Multi-field embedding gather (one launch for all F fields) and touched-row marking.
//
// Replaces BaseModel.input_from_feature_columns (reference meta_basemodel.py:519-545: per field one
// column slice, one float->long cast and one nn.Embedding call) plus the torch.cat at satrans.py:211.
//
// Layout: all tables back to back in one [total_rows][32] fp32 buffer; row_off[f] is the first row of
// field f.  A row is 128 B = one L2 line = 8 lanes x float4, so a warp moves 4 rows per instruction and
// every request is a full-line transaction; each thread keeps GATHER_UNROLL independent rows in flight.
#include "common.cuh"

namespace sat {

constexpr int GATHER_THREADS = 256;
constexpr int GATHER_UNROLL = 4;

template <int KIND>
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather(const SatModel m, const void* __restrict__ x, int ld, int B, float* __restrict__ out, int32_t* __restrict__ err) {
    const int64_t* __restrict__ row_off = m.row_off;
    const int32_t* __restrict__ cols = m.sparse_cols;
    const int F = m.F;
    const int lane8 = threadIdx.x & 7;                       // which float4 of the row
    const long long n_items = (long long)B * F;
    const long long group = ((long long)blockIdx.x * GATHER_THREADS + threadIdx.x) >> 3;   // row slot of this thread
    const long long n_groups = ((long long)gridDim.x * GATHER_THREADS) >> 3;
    bool bad = false;
    for (long long base = group; base < n_items; base += n_groups * GATHER_UNROLL) {
        const float* src[GATHER_UNROLL];
        long long item[GATHER_UNROLL];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
            item[u] = base + (long long)u * n_groups;
            src[u] = nullptr;
            if (item[u] < n_items) {
                const int b = (int)(item[u] / F), f = (int)(item[u] - (long long)b * F);
                long long id = load_id(x, KIND, (long long)b * ld + __ldg(cols + f));
                const long long lo = __ldg(row_off + f), hi = __ldg(row_off + f + 1);
                if (id < 0 || id >= hi - lo) { bad = true; id = 0; }
                src[u] = table_row(m, lo + id) + lane8 * 4;
            }
        }
        float4 v[GATHER_UNROLL];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u)
            if (src[u]) v[u] = __ldg(reinterpret_cast<const float4*>(src[u]));
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u)
            if (src[u]) __stcs(reinterpret_cast<float4*>(out + item[u] * D + lane8 * 4), v[u]);
    }
    if (bad && err) atomicOr(err, 1);
}

// One bit per table row read by this batch (used to rank touched rows and to let the dense optimiser
// sweep skip them).
template <int KIND>
__global__ void __launch_bounds__(256)
k_rows_mark(const SatModel m, const void* __restrict__ x, int ld, int B) {
    const int64_t* __restrict__ row_off = m.row_off;
    const int32_t* __restrict__ cols = m.sparse_cols;
    const int F = m.F;
    const long long n_items = (long long)B * F;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < n_items;
         it += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(it / F), f = (int)(it - (long long)b * F);
        long long id = load_id(x, KIND, (long long)b * ld + __ldg(cols + f));
        const long long lo = __ldg(row_off + f), hi = __ldg(row_off + f + 1);
        if (id < 0 || id >= hi - lo) id = 0;
        const long long row = lo + id;
        const long long lrow = row / m.n_shards;
        atomicOr(m.shard_bitmap[(int)(row % m.n_shards)] + (lrow >> 5), 1u << (lrow & 31));
    }
}

// exclusive prefix popcount over the bitmap words: two launches (block sums, then per-word ranks).
__global__ void __launch_bounds__(1024) k_rank_blocksum(const uint32_t* __restrict__ bitmap, long long n_words,
                                                        int32_t* __restrict__ block_sums) {
    __shared__ int s_w[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    int c = i < n_words ? __popc(bitmap[i]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_rank_words(const uint32_t* __restrict__ bitmap, long long n_words, const int32_t* __restrict__ block_sums,
             int32_t* __restrict__ word_rank, int32_t* __restrict__ uniq_rows, int32_t* __restrict__ counters,
             float* __restrict__ scalars) {
    __shared__ int s_w[32];
    __shared__ int s_base;
    // base of this block = sum of the block sums before it (<= a few hundred values)
    int part = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += 1024) part += block_sums[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = s_w[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) s_base = t;
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    const uint32_t word = i < n_words ? bitmap[i] : 0u;
    const int c = __popc(word);
    // inclusive warp scan
    int inc = c;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = s_w[threadIdx.x], v = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_w[threadIdx.x] = v - t;    // exclusive prefix of warp totals
    }
    __syncthreads();
    const int excl = s_base + s_w[wid] + inc - c;
    if (i < n_words) {
        word_rank[i] = excl;
        uint32_t wbits = word;
        int k = excl;
        while (wbits) {
            const int bit = __ffs(wbits) - 1;
            uniq_rows[k++] = (int32_t)(i * 32 + bit);
            wbits &= wbits - 1;
        }
        if (i == n_words - 1) {
            counters[CNT_NUNIQ] = excl + c;
            scalars[2] = (float)(excl + c);
        }
    }
}

}  // namespace sat

using namespace sat;

extern "C" int sat_gather_fwd(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, int32_t B, float* out,
                              int32_t* err_flag, void* stream) {
    if (!m || !x || !out || B < 0 || m->F <= 0) return sat_host::fail(SAT_EINVAL, "sat_gather_fwd: bad argument");
    if (B == 0) return SAT_OK;
    const long long items = (long long)B * m->F;
    const long long per_block = (GATHER_THREADS / 8) * GATHER_UNROLL;
    long long blocks = (items + per_block - 1) / per_block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t s = (cudaStream_t)stream;
    switch (id_kind) {
        case SAT_IDS_F32:
            k_gather<SAT_IDS_F32><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag);
            break;
        case SAT_IDS_I32:
            k_gather<SAT_IDS_I32><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag);
            break;
        case SAT_IDS_I64:
            k_gather<SAT_IDS_I64><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag);
            break;
        default:
            return sat_host::fail(SAT_EINVAL, "sat_gather_fwd: id_kind %d", id_kind);
    }
    return sat_host::check_launch("k_gather");
}

extern "C" int sat_rows_clear(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->bitmap) return sat_host::fail(SAT_EINVAL, "sat_rows_clear: bad argument");
    const long long n_words = (m->local_rows + 31) / 32;
    if (cudaMemsetAsync(w->bitmap, 0, (size_t)n_words * 4, (cudaStream_t)stream) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_rows_clear: memset failed");
    return SAT_OK;
}

extern "C" int sat_rows_mark(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w,
                             void* stream) {
    if (!m || !x || !w || !w->bitmap) return sat_host::fail(SAT_EINVAL, "sat_rows_mark: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long long items = (long long)w->B * m->F;
    if (items > 0) {
        int blocks = (int)((items + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        switch (id_kind) {
            case SAT_IDS_F32:
                k_rows_mark<SAT_IDS_F32><<<blocks, 256, 0, s>>>(*m, x, ld, w->B);
                break;
            case SAT_IDS_I32:
                k_rows_mark<SAT_IDS_I32><<<blocks, 256, 0, s>>>(*m, x, ld, w->B);
                break;
            case SAT_IDS_I64:
                k_rows_mark<SAT_IDS_I64><<<blocks, 256, 0, s>>>(*m, x, ld, w->B);
                break;
            default:
                return sat_host::fail(SAT_EINVAL, "sat_rows_mark: id_kind %d", id_kind);
        }
    }
    return sat_host::check_launch("k_rows_mark");
}

extern "C" int sat_rows_rank(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->bitmap || !w->word_rank || !w->block_sums || !w->uniq_rows)
        return sat_host::fail(SAT_EINVAL, "sat_rows_rank: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long long n_words = (m->local_rows + 31) / 32;
    const int nb = (int)((n_words + 1023) / 1024);
    k_rank_blocksum<<<nb, 1024, 0, s>>>(w->bitmap, n_words, w->block_sums);
    k_rank_words<<<nb, 1024, 0, s>>>(w->bitmap, n_words, w->block_sums, w->word_rank, w->uniq_rows, w->counters,
                                     w->scalars);
    return sat_host::check_launch("k_rank_words");
}

extern "C" int sat_rows_mark_and_rank(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w,
                                      void* stream) {
    int rc = sat_rows_clear(m, w, stream);
    if (rc) return rc;
    rc = sat_rows_mark(m, x, id_kind, ld, w, stream);
    if (rc) return rc;
    return sat_rows_rank(m, w, stream);
}

extern "C" int sat_enable_peer_access(int32_t peer_device) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return sat_host::fail(SAT_ECUDA, "sat_enable_peer_access: cudaGetDevice failed");
    if (cur == peer_device) return SAT_OK;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, cur, peer_device);
    if (!can) return sat_host::fail(SAT_ECUDA, "device %d cannot access peer %d", cur, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return sat_host::fail(SAT_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    cudaGetLastError();
    return SAT_OK;
}
