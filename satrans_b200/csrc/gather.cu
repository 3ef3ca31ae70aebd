// Multi-field embedding gather (one launch for all F fields) and touched-row marking.
//
// Replaces BaseModel.input_from_feature_columns (reference meta_basemodel.py:519-545: per field one
// column slice, one float->long cast and one nn.Embedding call) plus the torch.cat at satrans.py:211.
//
// Layout: all tables back to back in one [total_rows][32] fp32 buffer; row_off[f] is the first row of
// field f.  A row is 128 B = one L2 line = 8 lanes x float4, so a warp moves 4 rows per instruction and
// every request is a full-line transaction; each thread keeps GATHER_UNROLL independent rows in flight.
#include <stdlib.h>

#include "tc_common.cuh"

namespace sat {

constexpr int GATHER_THREADS = 256;
#ifndef SAT_GATHER_DEFAULT
#define SAT_GATHER_DEFAULT 5
#endif

template <int KIND, int GATHER_UNROLL, int MIN_BLOCKS>
__global__ void __launch_bounds__(GATHER_THREADS, MIN_BLOCKS)
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

// ------------------------------------------------------------------------------------------------------------------------
// TMA variant: every row travels table -> shared memory as ONE 128-byte bulk asynchronous copy (cp.async.bulk, no registers
// held while it is in flight), and a chunk of GB_ROWS consecutive output rows leaves as ONE bulk store shared -> global.
// GB_STAGES chunks per CTA circulate: GB_STAGES - 2 chunks of row fetches are in flight while the oldest is being stored.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int GB_ROWS = 128, GB_STAGES = 6, GB_LAG = GB_STAGES - 2;
constexpr int GB_SMEM_BYTES = GB_STAGES * GB_ROWS * D * 4 + GB_STAGES * 8;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

template <int KIND>
__global__ void __launch_bounds__(GB_ROWS)
k_gather_tma(const SatModel m, const void* __restrict__ x, int ld, int B, float* __restrict__ out, int32_t* __restrict__ err) {
    extern __shared__ __align__(128) unsigned char gsm[];
    float* stage = reinterpret_cast<float*>(gsm);
    uint64_t* full = reinterpret_cast<uint64_t*>(gsm + GB_STAGES * GB_ROWS * D * 4);
    const int tid = threadIdx.x, F = m.F;
    const long long n_items = (long long)B * F;
    const long long n_chunks = (n_items + GB_ROWS - 1) / GB_ROWS;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < GB_STAGES; ++s) mbar_init(full + s, 1);
    }
    __syncthreads();
    bool bad = false;
    // this CTA's chunks: c = blockIdx.x + i * gridDim.x
    const long long my_n = n_chunks > (long long)blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (long long i = 0; i < my_n + GB_LAG; ++i) {
        if (i < my_n) {
            const int s = (int)(i % GB_STAGES);
            // the store that last read this stage (chunk i - GB_STAGES, committed at iteration i - 2) must have read it
            if (tid == 0) bulk_wait_read<1>();
            __syncthreads();
            const long long c = blockIdx.x + i * gridDim.x;
            const long long item0 = c * GB_ROWS;
            const int rows = (int)min((long long)GB_ROWS, n_items - item0);
            if (tid == 0) mbar_expect_tx(full + s, (uint32_t)rows * D * 4);
            if (tid < rows) {
                const long long item = item0 + tid;
                const int b = (int)(item / F), f = (int)(item - (long long)b * F);
                long long id = load_id(x, KIND, (long long)b * ld + __ldg(m.sparse_cols + f));
                const long long lo = __ldg(m.row_off + f), hi = __ldg(m.row_off + f + 1);
                if (id < 0 || id >= hi - lo) { bad = true; id = 0; }
                tma_load_1d(stage + ((long long)s * GB_ROWS + tid) * D, table_row(m, lo + id), D * 4, full + s);
            }
        }
        const long long j = i - GB_LAG;                        // the chunk whose rows should have landed by now
        if (j >= 0 && tid == 0) {
            const int s = (int)(j % GB_STAGES);
            mbar_wait(full + s, (uint32_t)((j / GB_STAGES) & 1));
            const long long c = blockIdx.x + j * gridDim.x;
            const long long item0 = c * GB_ROWS;
            const int rows = (int)min((long long)GB_ROWS, n_items - item0);
            bulk_store(out + item0 * D, stage + (long long)s * GB_ROWS * D, (uint32_t)rows * D * 4);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
        // Row-sharded: the bit goes into THIS process' staging region for the owner shard (a local L2 atomic); the owner ORs the
        // peers' regions together with coalesced NVLink loads after a barrier (k_rows_merge).  One remote atomic per id -- a
        // 4-byte NVLink transaction each -- cost 2 ms of a 16 ms step at B = 65,536 on two GPUs.
        uint32_t* dst = m.n_shards == 1 ? m.shard_bitmap[0]
                                        : m.shard_bitmap[m.shard_rank] + (1 + (row % m.n_shards)) * sat_stage_words(&m);
        atomicOr(dst + (lrow >> 5), 1u << (lrow & 31));
    }
}

// final bitmap of this shard = OR over the processes p of p's staging region for this shard (volatile: peer memory)
__global__ void __launch_bounds__(256) k_rows_merge(const SatModel m, long long n_words) {
    const long long sw = sat_stage_words(&m);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_words; i += (long long)gridDim.x * 256) {
        uint32_t v = 0u;
        for (int p = 0; p < m.n_shards; ++p)
            v |= *reinterpret_cast<const volatile uint32_t*>(m.shard_bitmap[p] + (1 + m.shard_rank) * sw + i);
        m.shard_bitmap[m.shard_rank][i] = v;
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
    cudaStream_t s = (cudaStream_t)stream;
    // SAT_GATHER_VARIANT (tuning aid; measured in profiles/): 0 = 4 rows in flight per thread, 1 = 8 rows, 2 = 4 rows at full
    // occupancy (<= 32 registers), 3 = 8 rows at 6 CTAs per SM, 4 = TMA bulk row fetch + bulk store, 5 = 2 rows at full occupancy
    // (the default: measured best), 6 = 1 row at full occupancy
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("SAT_GATHER_VARIANT"); variant = e ? atoi(e) : SAT_GATHER_DEFAULT; }
#define SAT_GATHER_LAUNCH(KIND)                                                                                              \
    do {                                                                                                                     \
        if (variant == 4 && m->n_shards == 1) {                                                                              \
            static bool cfg = false;                                                                                         \
            if (!cfg) {                                                                                                      \
                if (cudaFuncSetAttribute(k_gather_tma<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, GB_SMEM_BYTES) != cudaSuccess) \
                    return sat_host::fail(SAT_ECUDA, "sat_gather_fwd: cannot raise dynamic shared memory");                  \
                cfg = true;                                                                                                  \
            }                                                                                                                \
            long long chunks = (items + GB_ROWS - 1) / GB_ROWS;                                                              \
            const int grid = (int)(chunks < 148 * 2 ? chunks : 148 * 2);                                                     \
            k_gather_tma<KIND><<<grid, GB_ROWS, GB_SMEM_BYTES, s>>>(*m, x, ld, B, out, err_flag);                            \
        } else {                                                                                                             \
            const int unroll = (variant == 1 || variant == 3) ? 8 : (variant == 5 ? 2 : (variant == 6 ? 1 : 4));                                                       \
            const long long per_block = (GATHER_THREADS / 8) * unroll;                                                       \
            long long blocks = (items + per_block - 1) / per_block;                                                          \
            if (blocks > 148 * 16) blocks = 148 * 16;                                                                        \
            if (variant == 1) k_gather<KIND, 8, 1><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag);      \
            else if (variant == 3) k_gather<KIND, 8, 6><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag); \
            else if (variant == 5) k_gather<KIND, 2, 8><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag); \
            else if (variant == 6) k_gather<KIND, 1, 8><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag); \
            else if (variant == 0) k_gather<KIND, 4, 1><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag); \
            else if (variant == 2) k_gather<KIND, 4, 8><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag); \
            else k_gather<KIND, 2, 8><<<(int)blocks, GATHER_THREADS, 0, s>>>(*m, x, ld, B, out, err_flag);                   \
        }                                                                                                                    \
    } while (0)
    switch (id_kind) {
        case SAT_IDS_F32: SAT_GATHER_LAUNCH(SAT_IDS_F32); break;
        case SAT_IDS_I32: SAT_GATHER_LAUNCH(SAT_IDS_I32); break;
        case SAT_IDS_I64: SAT_GATHER_LAUNCH(SAT_IDS_I64); break;
        default:
            return sat_host::fail(SAT_EINVAL, "sat_gather_fwd: id_kind %d", id_kind);
    }
#undef SAT_GATHER_LAUNCH
    return sat_host::check_launch("k_gather");
}

extern "C" int sat_rows_clear(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->bitmap) return sat_host::fail(SAT_EINVAL, "sat_rows_clear: bad argument");
    const long long n_words = (m->local_rows + 31) / 32;
    // row-sharded: the staging regions (the final bitmap is rewritten by sat_rows_rank's merge)
    void* p = m->n_shards == 1 ? (void*)w->bitmap : (void*)(w->bitmap + sat_stage_words(m));
    const size_t bytes = m->n_shards == 1 ? (size_t)n_words * 4 : (size_t)m->n_shards * sat_stage_words(m) * 4;
    if (cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream) != cudaSuccess)
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
    if (m->n_shards > 1) {              // (after the barrier that follows sat_rows_mark on every process)
        int mb = (int)((n_words + 255) / 256);
        if (mb > 148 * 8) mb = 148 * 8;
        k_rows_merge<<<mb, 256, 0, s>>>(*m, n_words);
    }
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
