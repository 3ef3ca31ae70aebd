// Scenario bucketing and the scenario encoder.
//
// The reference evaluates the scenario encoder per SAMPLE (reference satrans.py:203-234: embedding lookup,
// ReLU, DNN_v2 = one Linear(32|64 -> 4096), materialising a (B, 4096) tensor of generated weights) although
// its input takes only n_dom (EN) or n_dom*L*2 (ENP) distinct values.  Here samples are counting-sorted by
// scenario id so that every tile of the fused kernels shares one generated MetaNet, and the encoder is
// evaluated once per distinct key into theta[n_keys][4096].
#include "common.cuh"

namespace sat {

// ---------------------------------------------------------------------------------------------------
// Stable counting sort of the B samples by scenario id + tile list.  Single CTA (B is a batch, the keys
// are a handful); stable so that the tiling -- hence the fp32 summation order of weight gradients per
// tile -- is a deterministic function of the input.
// ---------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(1024)
k_bucket(const void* __restrict__ x, int ld, int dom_col, int B, int n_dom, int spt, int32_t* __restrict__ perm,
         int32_t* __restrict__ dom, int32_t* __restrict__ tiles, int32_t* __restrict__ counters, int spt2,
         int32_t* __restrict__ tiles2) {
    __shared__ int s_hist[MAX_KEYS_HIST];
    __shared__ int s_base[MAX_KEYS_HIST];
    __shared__ int s_tile_base[MAX_KEYS_HIST + 1];
    __shared__ int s_tile_base2[MAX_KEYS_HIST + 1];
    __shared__ int s_cnt[32][MAX_KEYS_HIST];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = tid; k < MAX_KEYS_HIST; k += 1024) s_hist[k] = 0;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int b = tid; b < B; b += 1024) {
        long long g = load_id(x, KIND, (long long)b * ld + dom_col);
        if (g < 0 || g >= n_dom) { s_bad = 1; g = 0; }
        atomicAdd(&s_hist[(int)g], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, tacc = 0, tacc2 = 0;
        for (int k = 0; k < n_dom; ++k) {
            s_base[k] = acc;
            s_tile_base[k] = tacc;
            s_tile_base2[k] = tacc2;
            acc += s_hist[k];
            tacc += (s_hist[k] + spt - 1) / spt;
            tacc2 += (s_hist[k] + spt2 - 1) / spt2;
            counters[CNT_HIST + k] = s_hist[k];
        }
        s_tile_base[n_dom] = tacc;
        s_tile_base2[n_dom] = tacc2;
        counters[CNT_NTILES] = tacc;
        counters[CNT_NTILES2] = tiles2 ? tacc2 : 0;
        if (s_bad) atomicOr(&counters[CNT_ERR], 2);
    }
    __syncthreads();
    // tiles: (key, start, count, 0)
    for (int k = 0; k < n_dom; ++k) {
        const int nt = s_tile_base[k + 1] - s_tile_base[k];
        for (int t = tid; t < nt; t += 1024) {
            int32_t* e = tiles + 4 * (s_tile_base[k] + t);
            const int start = s_base[k] + t * spt;
            const int cnt = min(spt, s_base[k] + s_hist[k] - start);
            e[0] = k; e[1] = start; e[2] = cnt; e[3] = 0;
        }
        if (tiles2) {
            const int nt2 = s_tile_base2[k + 1] - s_tile_base2[k];
            for (int t = tid; t < nt2; t += 1024) {
                int32_t* e = tiles2 + 4 * (s_tile_base2[k] + t);
                const int start = s_base[k] + t * spt2;
                const int cnt = min(spt2, s_base[k] + s_hist[k] - start);
                e[0] = k; e[1] = start; e[2] = cnt; e[3] = 0;
            }
        }
    }
    // stable placement, 1024 samples at a time
    for (int b0 = 0; b0 < B; b0 += 1024) {
        for (int k = lane; k < n_dom; k += 32) s_cnt[wid][k] = 0;
        __syncthreads();
        const int b = b0 + tid;
        int g = -1;
        if (b < B) {
            long long gg = load_id(x, KIND, (long long)b * ld + dom_col);
            g = (gg < 0 || gg >= n_dom) ? 0 : (int)gg;
        }
        const unsigned same = __match_any_sync(0xffffffffu, g);
        const int rank_in_warp = __popc(same & ((1u << lane) - 1u));
        if (g >= 0 && rank_in_warp == 0) s_cnt[wid][g] = __popc(same);
        __syncthreads();
        if (g >= 0) {
            int before = 0;
            for (int w2 = 0; w2 < wid; ++w2) before += s_cnt[w2][g];
            const int pos = s_base[g] + before + rank_in_warp;
            perm[pos] = b;
            dom[pos] = g;
        }
        __syncthreads();
        if (tid < n_dom) {
            int tot = 0;
            for (int w2 = 0; w2 < 32; ++w2) tot += s_cnt[w2][tid];
            s_base[tid] += tot;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Encoder input u[key][enc_in]:
//   EN : u = relu(Dom[g])                                        reference satrans.py:204,213,218
//   ENP: u = relu([relu(Dom[g]) | Lid[l] + Qkv[q]])              reference satrans.py:226-232
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void key_decode(int key, int L, bool pos, int& g, int& l, int& q) {
    if (pos) { q = key & 1; l = (key >> 1) % L; g = (key >> 1) / L; } else { g = key; l = 0; q = 0; }
}
__device__ __forceinline__ float enc_input(const float* __restrict__ dense, const SatModel& m, int key, int i,
                                           float* pre = nullptr) {
    const bool pos = m.flags & SAT_FLAG_POS;
    int g, l, q;
    key_decode(key, m.L, pos, g, l, q);
    float v;
    if (i < D) v = dense[m.off_dom + (long long)g * D + i];
    else v = dense[m.off_lid + (long long)l * D + (i - D)] + dense[m.off_qkv + (long long)q * D + (i - D)];
    if (pre) *pre = v;
    return fmaxf(v, 0.f);
}

constexpr int ENC_THREADS = 256;
constexpr int ENC_MAX_KEYS = 64;     // keys staged per pass through shared memory

// theta[key][o] = b[o] + sum_i u[key][i] * W[o][i].  One thread per output o, weight row in registers,
// loop over keys (u staged in smem).  grid = 4096 / 256.
__global__ void __launch_bounds__(ENC_THREADS) k_encoder_fwd(SatModel m, int n_keys, float* __restrict__ theta) {
    __shared__ float s_u[ENC_MAX_KEYS][2 * D];
    const int o = blockIdx.x * ENC_THREADS + threadIdx.x;
    const int ein = m.enc_in;
    const bool live = o < SAT_META_PSZ(m.flags, m.H);       // gate / bilinear variants generate fewer weights: the rest of theta is 0
    float w[2 * D];
#pragma unroll
    for (int i = 0; i < 2 * D; ++i) w[i] = (live && i < ein) ? m.dense[m.off_encw + (long long)o * ein + i] : 0.f;
    const float bias = live ? m.dense[m.off_encb + o] : 0.f;
    for (int k0 = 0; k0 < n_keys; k0 += ENC_MAX_KEYS) {
        const int nk = min(ENC_MAX_KEYS, n_keys - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < nk * 2 * D; t += ENC_THREADS) {
            const int k = t / (2 * D), i = t % (2 * D);
            s_u[k][i] = (i < ein) ? enc_input(m.dense, m, k0 + k, i) : 0.f;
        }
        __syncthreads();
        for (int k = 0; k < nk; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * D; ++i) acc = fmaf(s_u[k][i], w[i], acc);
            theta[(long long)(k0 + k) * PSZ + o] = acc + bias;
        }
    }
}

// Encoder backward, part 1: dW[o][i] = sum_k dtheta[k][o] u[k][i], db[o] = sum_k dtheta[k][o] (each o owned
// by one thread -> plain stores), and du[k][i] = sum_o dtheta[k][o] W[o][i]: every block reduces its 256 outputs in a fixed
// order and stores the partial into du_part[block][k][64]; k_encoder_bwd_emb adds the blocks in order (no atomics).
__global__ void __launch_bounds__(ENC_THREADS)
k_encoder_bwd(SatModel m, int n_keys, const float* __restrict__ dtheta, float* __restrict__ dgrad,
              float* __restrict__ du) {
    __shared__ float s_u[ENC_MAX_KEYS][2 * D];
    __shared__ float s_red[ENC_THREADS / 32][2 * D];
    const int o = blockIdx.x * ENC_THREADS + threadIdx.x;
    const int ein = m.enc_in;
    const bool live = o < SAT_META_PSZ(m.flags, m.H);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float w[2 * D], dw[2 * D];
#pragma unroll
    for (int i = 0; i < 2 * D; ++i) { w[i] = (live && i < ein) ? m.dense[m.off_encw + (long long)o * ein + i] : 0.f; dw[i] = 0.f; }
    float db = 0.f;
    for (int k0 = 0; k0 < n_keys; k0 += ENC_MAX_KEYS) {
        const int nk = min(ENC_MAX_KEYS, n_keys - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < nk * 2 * D; t += ENC_THREADS) {
            const int k = t / (2 * D), i = t % (2 * D);
            s_u[k][i] = (i < ein) ? enc_input(m.dense, m, k0 + k, i) : 0.f;
        }
        __syncthreads();
        for (int k = 0; k < nk; ++k) {
            const float g = live ? dtheta[(long long)(k0 + k) * PSZ + o] : 0.f;
            db += g;
            // du[k][i]: reduce g*w[i] over the 256 outputs of this block
#pragma unroll
            for (int i = 0; i < 2 * D; ++i) {
                dw[i] = fmaf(g, s_u[k][i], dw[i]);
                float t = g * w[i];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                if (lane == 0) s_red[wid][i] = t;
            }
            __syncthreads();
            if (threadIdx.x < ein) {
                float t = 0.f;
#pragma unroll
                for (int w2 = 0; w2 < ENC_THREADS / 32; ++w2) t += s_red[w2][threadIdx.x];
                du[((long long)blockIdx.x * n_keys + (k0 + k)) * 2 * D + threadIdx.x] = t;
            }
            __syncthreads();
        }
    }
    if (!live) return;
#pragma unroll
    for (int i = 0; i < 2 * D; ++i)
        if (i < ein) dgrad[m.off_encw + (long long)o * ein + i] = dw[i];
    dgrad[m.off_encb + o] = db;
}

// Encoder backward, part 2: through the ReLUs into the domain / layer-id / qkv-id embedding gradients.
__global__ void k_encoder_bwd_emb(SatModel m, int n_keys, const float* __restrict__ du_part, int n_part,
                                  float* __restrict__ dgrad) {
    const bool pos = m.flags & SAT_FLAG_POS;
    const int i = threadIdx.x;          // 0 .. enc_in-1
    if (i >= m.enc_in) return;
    for (int key = 0; key < n_keys; ++key) {
        int g, l, q;
        key_decode(key, m.L, pos, g, l, q);
        float pre;
        enc_input(m.dense, m, key, i, &pre);
        float du = 0.f;
        for (int pb = 0; pb < n_part; ++pb) du += du_part[((long long)pb * n_keys + key) * 2 * D + i];
        const float d = pre > 0.f ? du : 0.f;
        if (i < D) {
            dgrad[m.off_dom + (long long)g * D + i] += d;          // thread i owns column i: no race
        } else {
            dgrad[m.off_lid + (long long)l * D + (i - D)] += d;
            dgrad[m.off_qkv + (long long)q * D + (i - D)] += d;
        }
    }
}

// 'onlyemb' (reference satrans.py:173-176): no encoder, theta[key] = relu(domain_embeddings[g(key)]) with a [n_dom][4096] table
// (with 'pos' the layer / qkv id embeddings sit behind the first 4096 entries of the concatenation and are sliced away, :63,70).
__global__ void __launch_bounds__(ENC_THREADS) k_onlyemb_fwd(SatModel m, int n_keys, float* __restrict__ theta) {
    const int o = blockIdx.x * ENC_THREADS + threadIdx.x;
    const bool pos = m.flags & SAT_FLAG_POS;
    for (int key = 0; key < n_keys; ++key) {
        int g, l, q;
        key_decode(key, m.L, pos, g, l, q);
        theta[(long long)key * PSZ + o] = fmaxf(m.dense[m.off_dom + (long long)g * PSZ + o], 0.f);
    }
}
// ... and its backward: thread o owns column o of every row; the keys of one scenario are added in key order
__global__ void __launch_bounds__(ENC_THREADS) k_onlyemb_bwd(SatModel m, int n_keys, const float* __restrict__ dtheta,
                                                             float* __restrict__ dgrad) {
    const int o = blockIdx.x * ENC_THREADS + threadIdx.x;
    const bool pos = m.flags & SAT_FLAG_POS;
    for (int key = 0; key < n_keys; ++key) {
        int g, l, q;
        key_decode(key, m.L, pos, g, l, q);
        const long long e = m.off_dom + (long long)g * PSZ + o;
        if (m.dense[e] > 0.f) dgrad[e] += dtheta[(long long)key * PSZ + o];
    }
}

}  // namespace sat

using namespace sat;

static int n_keys_of(const SatModel* m) {
    return (m->flags & SAT_FLAG_POS) ? m->n_dom * m->L * 2 : m->n_dom;
}

extern "C" int sat_bucket_by_domain(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w,
                                    int32_t samples_per_tile, void* stream) {
    if (!m || !x || !w) return sat_host::fail(SAT_EINVAL, "sat_bucket_by_domain: null argument");
    if (m->n_dom > MAX_KEYS_HIST) return sat_host::fail(SAT_EINVAL, "sat_bucket_by_domain: %d scenario ids > %d", m->n_dom, MAX_KEYS_HIST);
    if (samples_per_tile <= 0) samples_per_tile = NT / m->F;
    cudaStream_t s = (cudaStream_t)stream;
    switch (id_kind) {
        case SAT_IDS_F32:
            k_bucket<SAT_IDS_F32><<<1, 1024, 0, s>>>(x, ld, m->dom_col, w->B, m->n_dom, samples_per_tile, w->perm, w->dom, w->tiles, w->counters, NT2 / m->F, w->tiles2);
            break;
        case SAT_IDS_I32:
            k_bucket<SAT_IDS_I32><<<1, 1024, 0, s>>>(x, ld, m->dom_col, w->B, m->n_dom, samples_per_tile, w->perm, w->dom, w->tiles, w->counters, NT2 / m->F, w->tiles2);
            break;
        case SAT_IDS_I64:
            k_bucket<SAT_IDS_I64><<<1, 1024, 0, s>>>(x, ld, m->dom_col, w->B, m->n_dom, samples_per_tile, w->perm, w->dom, w->tiles, w->counters, NT2 / m->F, w->tiles2);
            break;
        default:
            return sat_host::fail(SAT_EINVAL, "sat_bucket_by_domain: id_kind %d", id_kind);
    }
    return sat_host::check_launch("k_bucket");
}

extern "C" int sat_encoder_fwd(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->theta) return sat_host::fail(SAT_EINVAL, "sat_encoder_fwd: null argument");
    if (m->flags & SAT_FLAG_ONLYEMB) {
        k_onlyemb_fwd<<<PSZ / ENC_THREADS, ENC_THREADS, 0, (cudaStream_t)stream>>>(*m, n_keys_of(m), w->theta);
        return sat_host::check_launch("k_onlyemb_fwd");
    }
    k_encoder_fwd<<<PSZ / ENC_THREADS, ENC_THREADS, 0, (cudaStream_t)stream>>>(*m, n_keys_of(m), w->theta);
    return sat_host::check_launch("k_encoder_fwd");
}

extern "C" int sat_encoder_bwd(const SatModel* m, SatWork* w, void* stream) {
    if (!m || !w || !w->dtheta || !w->dgrad || !w->enc_du) return sat_host::fail(SAT_EINVAL, "sat_encoder_bwd: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int nk = n_keys_of(m);
    if (m->flags & SAT_FLAG_ONLYEMB) {
        k_onlyemb_bwd<<<PSZ / ENC_THREADS, ENC_THREADS, 0, s>>>(*m, nk, w->dtheta, w->dgrad);
        return sat_host::check_launch("k_onlyemb_bwd");
    }
    k_encoder_bwd<<<PSZ / ENC_THREADS, ENC_THREADS, 0, s>>>(*m, nk, w->dtheta, w->dgrad, w->enc_du);
    k_encoder_bwd_emb<<<1, 64, 0, s>>>(*m, nk, w->enc_du, PSZ / ENC_THREADS, w->dgrad);
    return sat_host::check_launch("k_encoder_bwd");
}
