// Sorted-index scatter-add of the embedding-row gradients (the backward of the gather, reference meta_basemodel.py:533-535
// under torch autograd: embedding_dense_backward) without a single floating-point atomic.
//
//   sat_rows_sort    key(token) = (owner shard << slot_bits) | slot, slot = rank of the token's table row among the rows
//                    this batch touches (prefix popcount of the touched-row bitmap);  LSD radix sort of (key, token) pairs,
//                    SAT_SORT_RADIX_BITS per pass, stable, so equal keys stay in ascending token order.
//   sat_rows_reduce  the sorted array is cut into fixed chunks of SAT_ROWS_CHUNK entries; an 8-lane group (one float4 of the
//                    128-byte row per lane) adds the dE0 rows of a chunk run by run in sorted order.  A run inside one chunk
//                    is final; the pieces of a run that crosses chunk boundaries are added in chunk order by k_rows_fixup.
//                    n_shards == 1: the sum goes to row_grad[slot] (consumed by k_adam_rows).  n_shards > 1: ONE row per
//                    unique (source shard, table row) is pushed into the owner's recv buffer over NVLink (plain stores).
//   sat_rows_combine (n_shards > 1, on the owner) adds the pushed rows of a slot in source-shard order.
// The grouping of every sum is a function of the sorted order alone, so two runs give bit-identical gradients.
#include "common.cuh"

namespace sat {

constexpr int RB = SAT_SORT_RADIX_BITS, RADIX = 1 << RB;
constexpr int SORT_THREADS = 256, SORT_WARPS = SORT_THREADS / 32, SORT_TILE = SAT_SORT_TILE;
constexpr int RCH = SAT_ROWS_CHUNK;
constexpr int CI_HEAD = 1, CI_TAIL = 2, CI_SINGLE = 4;     // chunk_info flags
constexpr int PF_EPOCH = 8, PF_TIMEOUT = 9, PF_CURSOR = 16;   // words of SatWork.peer_flags

// ---------------------------------------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256)
k_sort_keys(const SatModel m, const void* __restrict__ x, int ld, int B, const int32_t* __restrict__ perm, int slot_bits,
            uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int F = m.F;
    const long long N = (long long)B * F;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < N; t += (long long)gridDim.x * 256) {
        const int p = (int)(t / F), f = (int)(t - (long long)p * F);
        const int b = perm[p];
        long long id = load_id(x, KIND, (long long)b * ld + __ldg(m.sparse_cols + f));
        const long long lo = __ldg(m.row_off + f), hi = __ldg(m.row_off + f + 1);
        if (id < 0 || id >= hi - lo) id = 0;
        const long long row = lo + id;
        int owner = 0;
        long long lrow = row;
        if (m.n_shards > 1) { owner = (int)(row % m.n_shards); lrow = row / m.n_shards; }
        // row-sharded: the owner's final bitmap / ranks were mirrored into this process' regions 1 + owner by k_rows_pull
        const long long off = m.n_shards > 1 ? (1 + owner) * sat_stage_words(&m) : 0;
        const uint32_t wbits = m.shard_bitmap[m.shard_rank][off + (lrow >> 5)];
        const int slot = m.shard_word_rank[m.shard_rank][off + (lrow >> 5)] + __popc(wbits & ((1u << (lrow & 31)) - 1u));
        keys[t] = ((uint32_t)owner << slot_bits) | (uint32_t)slot;
        vals[t] = (int32_t)t;
    }
}

// Row-sharded: every shard's final touched-row bitmap and word ranks -> this process' mirror regions, with coalesced 16-byte
// NVLink loads (volatile: peer memory).  One remote 4-byte load pair per TOKEN in k_sort_keys cost 1.5 ms of a 15 ms step at
// B = 65,536 on two GPUs; the bulk pull moves total_rows / 4 bytes per step.
__global__ void __launch_bounds__(256) k_rows_pull(const SatModel m) {
    const long long sw = sat_stage_words(&m), n4 = sw / 4;
    for (int d = 0; d < m.n_shards; ++d) {
        const uint4* sb = reinterpret_cast<const uint4*>(m.shard_bitmap[d]);
        const uint4* sr = reinterpret_cast<const uint4*>(m.shard_word_rank[d]);
        uint4* db = reinterpret_cast<uint4*>(m.shard_bitmap[m.shard_rank] + (1 + d) * sw);
        uint4* dr = reinterpret_cast<uint4*>(m.shard_word_rank[m.shard_rank] + (1 + d) * sw);
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
            uint4 a, b;
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(sb + i));
            asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(sr + i));
            db[i] = a;
            dr[i] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// LSD radix sort pass: histogram per tile -> exclusive scan (digit-major, tile-minor) -> stable scatter
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, int N, int shift, int32_t* __restrict__ hist, int nblk) {
    __shared__ int s_h[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) s_h[i] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
    for (int j = threadIdx.x; j < SORT_TILE; j += SORT_THREADS) {
        const int i = base + j;
        if (i < N) atomicAdd(&s_h[(keys[i] >> shift) & (RADIX - 1)], 1);      // integer counts: order-independent
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) hist[d * nblk + blockIdx.x] = s_h[d];
}

__global__ void __launch_bounds__(1024) k_radix_scan(int32_t* __restrict__ hist, int n) {
    __shared__ int s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int a0 = min(n, tid * per), a1 = min(n, a0 + per);
    int sum = 0;
    for (int i = a0; i < a1; ++i) sum += hist[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (tid < 32) {
        const int t = s_w[tid];
        int v = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_w[tid] = v - t;
    }
    __syncthreads();
    int run = s_w[wid] + inc - sum;
    for (int i = a0; i < a1; ++i) {
        const int c = hist[i];
        hist[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                int32_t* __restrict__ vals_out, int N, int shift, const int32_t* __restrict__ hist, int nblk) {
    __shared__ int s_base[RADIX];
    __shared__ int s_cnt[SORT_WARPS][RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int d = tid; d < RADIX; d += SORT_THREADS) {
        s_base[d] = hist[d * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) s_cnt[w][d] = 0;
    }
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
    for (int j = 0; j < SORT_TILE; j += SORT_THREADS) {          // elements of the tile in order, SORT_THREADS at a time
        const int i = base + j + tid;
        const bool valid = i < N;
        uint32_t key = 0u;
        int32_t val = 0;
        uint32_t d = 0xFFFFFFFFu;                                 // lanes past the end form their own group
        if (valid) { key = keys_in[i]; val = vals_in[i]; d = (key >> shift) & (RADIX - 1); }
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(same & ((1u << lane) - 1u));      // earlier lanes with the same digit
        if (valid && rank == 0) s_cnt[warp][d] = __popc(same);
        __syncthreads();
        if (valid) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_cnt[w][d];
            const int pos = s_base[d] + before + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        for (int dd = tid; dd < RADIX; dd += SORT_THREADS) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) { tot += s_cnt[w][dd]; s_cnt[w][dd] = 0; }
            s_base[dd] += tot;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// segmented reduce over the sorted order
// ---------------------------------------------------------------------------------------------------------------------
struct RowsOut {
    float* row_grad;            // n_shards == 1: [slot][32]
    // n_shards > 1: push to the owner
    float* recv[SAT_MAX_SHARDS];
    int32_t* recv_idx[SAT_MAX_SHARDS];
    int32_t* cursor;            // [n_shards] rows pushed to each owner so far (local)
    long long recv_stride;      // floats between the source blocks of recv: cap_tok * 32
    long long idx_stride;       // entries between the source blocks of recv_idx
    int n_shards, rank, slot_bits;
};

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// the finished sum of one (owner, slot) run; executed by the 8 lanes of a group together
__device__ __forceinline__ void rows_out(const RowsOut& o, uint32_t key, const float4& acc, int lane8, unsigned gmask) {
    if (o.n_shards == 1) {
        *reinterpret_cast<float4*>(o.row_grad + (long long)key * D + lane8 * 4) = acc;
        return;
    }
    const int owner = (int)(key >> o.slot_bits);
    const uint32_t slot = key & ((1u << o.slot_bits) - 1u);
    int j = 0;
    if (lane8 == 0) j = atomicAdd(o.cursor + owner, 1);       // only a storage location: the arithmetic order is fixed by the owner
    j = __shfl_sync(gmask, j, (threadIdx.x & 31) & 24);
    *reinterpret_cast<float4*>(o.recv[owner] + (long long)o.rank * o.recv_stride + (long long)j * D + lane8 * 4) = acc;
    if (lane8 == 0) o.recv_idx[owner][(long long)o.rank * o.idx_stride + slot] = j;
}

__global__ void __launch_bounds__(256)
k_rows_reduce(const uint32_t* __restrict__ keys, const int32_t* __restrict__ vals, int N, const float* __restrict__ dx,
              int32_t* __restrict__ chunk_info, float* __restrict__ pieces, const RowsOut o) {
    const int group = (blockIdx.x * 256 + threadIdx.x) >> 3, lane8 = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & 24);
    const int n_chunks = (N + RCH - 1) / RCH;
    if (group >= n_chunks) return;
    const int i0 = group * RCH, i1 = min(N, i0 + RCH);
    uint32_t cur = keys[i0];
    const bool head_cont = i0 > 0 && keys[i0 - 1] == cur;
    bool first_run = true;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = i0; j0 < i1; j0 += 8) {
        uint32_t kk[8];
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {                             // 8 independent row fetches in flight
            const int i = j0 + u;
            kk[u] = 0u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < i1) {
                kk[u] = keys[i];
                v[u] = __ldcg(reinterpret_cast<const float4*>(dx + (long long)vals[i] * D + lane8 * 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (j0 + u < i1) {
                if (kk[u] != cur) {
                    if (first_run && head_cont)
                        *reinterpret_cast<float4*>(pieces + ((long long)group * 2 + 0) * D + lane8 * 4) = acc;
                    else
                        rows_out(o, cur, acc, lane8, gmask);
                    acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    cur = kk[u];
                    first_run = false;
                }
                add4(acc, v[u]);
            }
        }
    }
    const bool tail_cont = i1 < N && keys[i1] == cur;
    if (first_run && head_cont)                                   // the run came in from the previous chunk (and may go on)
        *reinterpret_cast<float4*>(pieces + ((long long)group * 2 + 0) * D + lane8 * 4) = acc;
    else if (tail_cont)                                           // the run starts here and goes on: this chunk owns it
        *reinterpret_cast<float4*>(pieces + ((long long)group * 2 + 1) * D + lane8 * 4) = acc;
    else
        rows_out(o, cur, acc, lane8, gmask);
    if (lane8 == 0) {
        int4 ci;
        ci.x = (head_cont ? CI_HEAD : 0) | (tail_cont ? CI_TAIL : 0) | (first_run ? CI_SINGLE : 0);
        ci.y = (int)cur; ci.z = 0; ci.w = 0;
        *reinterpret_cast<int4*>(chunk_info + 4 * group) = ci;
    }
}

// runs that cross chunk boundaries: the chunk where a run starts adds the pieces of the following chunks in chunk order
__global__ void __launch_bounds__(256)
k_rows_fixup(int N, const int32_t* __restrict__ chunk_info, const float* __restrict__ pieces, const RowsOut o) {
    const int group = (blockIdx.x * 256 + threadIdx.x) >> 3, lane8 = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & 24);
    const int n_chunks = (N + RCH - 1) / RCH;
    if (group >= n_chunks) return;
    const int4 ci = *reinterpret_cast<const int4*>(chunk_info + 4 * group);
    const bool owns = (ci.x & CI_TAIL) && !((ci.x & CI_SINGLE) && (ci.x & CI_HEAD));
    if (!owns) return;
    float4 acc = *reinterpret_cast<const float4*>(pieces + ((long long)group * 2 + 1) * D + lane8 * 4);
    bool done = false;
    for (int j = group + 1; !done; j += 8) {
        float4 pv[8];
        int fl[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {                             // speculative: 8 following chunks in flight
            const int jj = min(j + u, n_chunks - 1);
            pv[u] = *reinterpret_cast<const float4*>(pieces + ((long long)jj * 2 + 0) * D + lane8 * 4);
            fl[u] = chunk_info[4 * jj];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (!done) {
                add4(acc, pv[u]);                                 // chunk j+u continues the run (head piece)
                const bool goes_on = (fl[u] & CI_SINGLE) && (fl[u] & CI_TAIL) && (j + u < n_chunks - 1);
                if (!goes_on) done = true;
            }
        }
    }
    rows_out(o, (uint32_t)ci.y, acc, lane8, gmask);
}

// row-sharded mode, on the owner: rows pushed by the source shards -> row_grad[slot], sources added in shard order
__global__ void __launch_bounds__(256)
k_rows_combine(const int32_t* __restrict__ counters, int n_shards, const float* __restrict__ recv, long long recv_stride,
               int32_t* __restrict__ recv_idx, long long idx_stride, float* __restrict__ row_grad) {
    const int n_unique = counters[CNT_NUNIQ];
    const int lane8 = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & 24);
    for (int s = (blockIdx.x * 256 + threadIdx.x) >> 3; s < n_unique; s += (gridDim.x * 256) >> 3) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j[SAT_MAX_SHARDS];
#pragma unroll
        for (int r = 0; r < SAT_MAX_SHARDS; ++r) j[r] = r < n_shards ? recv_idx[(long long)r * idx_stride + s] : -1;
#pragma unroll
        for (int r = 0; r < SAT_MAX_SHARDS; ++r)
            if (j[r] >= 0) add4(acc, __ldcg(reinterpret_cast<const float4*>(recv + (long long)r * recv_stride + (long long)j[r] * D + lane8 * 4)));
        *reinterpret_cast<float4*>(row_grad + (long long)s * D + lane8 * 4) = acc;
        __syncwarp(gmask);                                        // every lane of the group has read the indices
        if (lane8 == 0) {
#pragma unroll
            for (int r = 0; r < SAT_MAX_SHARDS; ++r)
                if (j[r] >= 0) recv_idx[(long long)r * idx_stride + s] = -1;      // empty for the next step
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// cross-process barrier through peer-mapped flags
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_peer_barrier(const SatModel m, uint32_t* __restrict__ my_flags) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        const uint32_t e = my_flags[PF_EPOCH] + 1u;
        my_flags[PF_EPOCH] = e;
        s_epoch = e;
        __threadfence_system();                 // everything earlier kernels of this stream wrote (also into peer memory) first
    }
    __syncthreads();
    const uint32_t e = s_epoch;
    const int r = threadIdx.x;
    if (r < m.n_shards) {
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(m.shard_flags[r] + m.shard_rank), "r"(e) : "memory");
        uint32_t seen;
        unsigned long long t_start, t_now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_flags + r) : "memory");
            if ((int32_t)(seen - e) < 0) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                if (t_now - t_start > 20000000000ull) {       // 20 s: a peer never arrived -- give up instead of hanging the GPU
                    my_flags[PF_TIMEOUT] = 1u + (uint32_t)r;    // (the host raises on it: Engine.check_errors)
                    break;
                }
            }
        } while ((int32_t)(seen - e) < 0);
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace sat

using namespace sat;

static int bits_for(long long n) {              // smallest b with 2^b >= n
    int b = 0;
    while ((1ll << b) < n) ++b;
    return b;
}

// sizes shared by sat_workspace_sizes and the entry points below
long long sat_touched_cap(const SatModel* m, long long cap_tok) {
    long long cap = cap_tok * (m->n_shards > 0 ? m->n_shards : 1);      // rows of this shard any process may touch
    if (cap > m->local_rows) cap = m->local_rows;
    return cap;
}
long long sat_idx_stride(const SatModel* m, long long cap_tok) {        // identical on every shard
    long long cap = cap_tok * m->n_shards;
    const long long max_local = (m->total_rows + m->n_shards - 1) / m->n_shards;
    return cap < max_local ? cap : max_local;
}
static int slot_bits_of(const SatModel* m, long long cap_tok) { return bits_for(sat_idx_stride(m, cap_tok) + 1); }
static int sort_passes(const SatModel* m, long long cap_tok) {
    const int key_bits = slot_bits_of(m, cap_tok) + bits_for(m->n_shards);
    return (key_bits + RB - 1) / RB;
}

extern "C" int sat_rows_sort(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, void* stream) {
    if (!m || !x || !w || !w->sort_keys || !w->sort_vals || !w->sort_hist || !w->perm)
        return sat_host::fail(SAT_EINVAL, "sat_rows_sort: null argument / workspace not sized for training");
    const long long N = (long long)w->B * m->F, cap = (long long)w->cap_B * m->F;
    if (N == 0) return SAT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int slot_bits = slot_bits_of(m, cap);
    if (slot_bits + bits_for(m->n_shards) > 32) return sat_host::fail(SAT_EINVAL, "sat_rows_sort: key does not fit 32 bits");
    int blocks = (int)((N + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (m->n_shards > 1) {          // (after the barrier that follows sat_rows_rank on every process)
        k_rows_pull<<<148 * 4, 256, 0, s>>>(*m);
        int rc = sat_host::check_launch("k_rows_pull");
        if (rc) return rc;
    }
    switch (id_kind) {
        case SAT_IDS_F32: k_sort_keys<SAT_IDS_F32><<<blocks, 256, 0, s>>>(*m, x, ld, w->B, w->perm, slot_bits, w->sort_keys, w->sort_vals); break;
        case SAT_IDS_I32: k_sort_keys<SAT_IDS_I32><<<blocks, 256, 0, s>>>(*m, x, ld, w->B, w->perm, slot_bits, w->sort_keys, w->sort_vals); break;
        case SAT_IDS_I64: k_sort_keys<SAT_IDS_I64><<<blocks, 256, 0, s>>>(*m, x, ld, w->B, w->perm, slot_bits, w->sort_keys, w->sort_vals); break;
        default: return sat_host::fail(SAT_EINVAL, "sat_rows_sort: id_kind %d", id_kind);
    }
    const int passes = sort_passes(m, cap);
    const int nblk = (int)((N + SORT_TILE - 1) / SORT_TILE);
    for (int p = 0; p < passes; ++p) {
        const uint32_t* kin = w->sort_keys + (p & 1) * cap;
        const int32_t* vin = w->sort_vals + (p & 1) * cap;
        uint32_t* kout = w->sort_keys + ((p + 1) & 1) * cap;
        int32_t* vout = w->sort_vals + ((p + 1) & 1) * cap;
        k_radix_hist<<<nblk, SORT_THREADS, 0, s>>>(kin, (int)N, p * RB, w->sort_hist, nblk);
        k_radix_scan<<<1, 1024, 0, s>>>(w->sort_hist, RADIX * nblk);
        k_radix_scatter<<<nblk, SORT_THREADS, 0, s>>>(kin, vin, kout, vout, (int)N, p * RB, w->sort_hist, nblk);
    }
    return sat_host::check_launch("sat_rows_sort");
}

static void fill_out(const SatModel* m, SatWork* w, RowsOut& o) {
    const long long cap = (long long)w->cap_B * m->F;
    o.row_grad = w->row_grad;
    for (int r = 0; r < SAT_MAX_SHARDS; ++r) { o.recv[r] = m->shard_recv[r]; o.recv_idx[r] = m->shard_recv_idx[r]; }
    o.cursor = w->peer_flags ? reinterpret_cast<int32_t*>(w->peer_flags) + PF_CURSOR : nullptr;
    o.recv_stride = cap * D;
    o.idx_stride = sat_idx_stride(m, cap);
    o.n_shards = m->n_shards; o.rank = m->shard_rank; o.slot_bits = slot_bits_of(m, cap);
}

extern "C" int sat_rows_reduce(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->sort_keys || !w->sort_vals || !w->dx || !w->chunk_info || !w->pieces || !w->row_grad)
        return sat_host::fail(SAT_EINVAL, "sat_rows_reduce: null argument / workspace not sized for training");
    if (m->n_shards > 1 && (!w->peer_flags || !m->shard_recv[m->shard_rank] || !m->shard_recv_idx[m->shard_rank]))
        return sat_host::fail(SAT_EINVAL, "sat_rows_reduce: row-sharded mode needs the peer-mapped recv buffers");
    const long long N = (long long)w->B * m->F, cap = (long long)w->cap_B * m->F;
    if (N == 0) return SAT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int fin = sort_passes(m, cap) & 1;                     // ping-pong half that holds the sorted order
    const uint32_t* keys = w->sort_keys + fin * cap;
    const int32_t* vals = w->sort_vals + fin * cap;
    RowsOut o;
    fill_out(m, w, o);
    if (m->n_shards > 1 &&
        cudaMemsetAsync(w->peer_flags + PF_CURSOR, 0, SAT_MAX_SHARDS * sizeof(int32_t), s) != cudaSuccess)
        return sat_host::fail(SAT_ECUDA, "sat_rows_reduce: memset failed");
    const int n_chunks = (int)((N + RCH - 1) / RCH);
    const int blocks = (n_chunks * 8 + 255) / 256;
    k_rows_reduce<<<blocks, 256, 0, s>>>(keys, vals, (int)N, w->dx, w->chunk_info, w->pieces, o);
    k_rows_fixup<<<blocks, 256, 0, s>>>((int)N, w->chunk_info, w->pieces, o);
    return sat_host::check_launch("sat_rows_reduce");
}

extern "C" int sat_rows_combine(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->row_grad || !w->counters) return sat_host::fail(SAT_EINVAL, "sat_rows_combine: null argument");
    if (m->n_shards == 1) return SAT_OK;
    if (!w->recv || !w->recv_idx) return sat_host::fail(SAT_EINVAL, "sat_rows_combine: workspace has no recv buffers");
    const long long cap = (long long)w->cap_B * m->F;
    k_rows_combine<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(w->counters, m->n_shards, w->recv, cap * D, w->recv_idx,
                                                            sat_idx_stride(m, cap), w->row_grad);
    return sat_host::check_launch("k_rows_combine");
}

extern "C" int sat_peer_barrier(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w) return sat_host::fail(SAT_EINVAL, "sat_peer_barrier: null argument");
    if (m->n_shards == 1) return SAT_OK;
    if (!w->peer_flags) return sat_host::fail(SAT_EINVAL, "sat_peer_barrier: workspace has no peer flags");
    for (int r = 0; r < m->n_shards; ++r)
        if (!m->shard_flags[r]) return sat_host::fail(SAT_EINVAL, "sat_peer_barrier: peer flags of shard %d are not mapped", r);
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(*m, w->peer_flags);
    return sat_host::check_launch("k_peer_barrier");
}
