// Field self-attention layers of the baselines that share SATrans' attention front end ("next" rows of SURVEY.md section 8f,
// config 5 of BASELINE.json), through the same per-sample field-attention code as SATrans:
//   SelfAttention_Layer  (reference submodules.py:178-238, star_trans.py:335-396):
//       Q,K,V = X W_{Q,K,V};  A = dropout(softmax(Q K^T / sqrt(dh)));  R = dropout(A V);  out = LayerNorm(relu(R + X W_Res))
//       (W_Out is declared by the reference but never used)
//   MDR_InteractingLayer (reference star_trans.py:146-242): the same with ONE W_Query AND ONE LayerNorm PER SCENARIO (the merge
//       'sum' / 'dot' with the shared W_Query is a 32x32 elementwise op done by the caller); a sample whose scenario id is
//       outside 1..n keeps a zero output
//   InteractingLayer     (deepctr-torch 0.2.9, AutoInt; not vendored -> parity unpinned): no 1/sqrt(dh) scaling by default,
//       no dropout, no LayerNorm:  out = relu(softmax(Q K^T) V + X W_Res)
// One thread = one field token, a CTA = floor(192/F) whole samples; the forward (inference and train-mode dropout; backward: selfatt_bwd.cu).
#include "fwd_common.cuh"

namespace sat {

constexpr int FA_MAX_WQ = 8;           // scenario-specific W_Query / LayerNorm sets held in shared memory

struct SelfAttArgs {
    const float* x;        // [B][F][32]
    float* out;            // [B][F][32]
    const float* wq;       // [n_wq][32][32], used as x @ W: one matrix, or one per scenario
    const float* wk; const float* wv; const float* wres;   // [32][32] each
    const float* ln_g; const float* ln_b;                  // [n_wq][32] each, or NULL (no LayerNorm)
    const int32_t* dom;    // [B] scenario ids (sample uses set dom-1; outside 1..n_wq: zero output), or NULL (set 0)
    int B, F, use_res, training, n_wq, scaling;
    float p_drop;
    unsigned long long seed;
};

template <int H>
__global__ void __launch_bounds__(NT, 2) k_selfatt_fwd(const SelfAttArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                       // W_Key | W_Value | W_Res | W_Query[n_wq]
    float* sK = sW + (3 + a.n_wq) * 1024;   // [NT][36]
    float* sV = sK + NT * KV_LD;            // [NT][36]
    float* sLN = sV + NT * KV_LD;           // [n_wq] x (gamma | beta)
    const int tid = threadIdx.x, F = a.F;
    const int spt = NT / F;
    const int s = tid / F, f = tid - s * F;
    copy_to_smem(sW, a.wk, 1024, tid, NT);
    copy_to_smem(sW + 1024, a.wv, 1024, tid, NT);
    if (a.use_res) copy_to_smem(sW + 2048, a.wres, 1024, tid, NT);
    copy_to_smem(sW + 3072, a.wq, a.n_wq * 1024, tid, NT);
    if (a.ln_g)
        for (int i = tid; i < a.n_wq * 32; i += NT) { sLN[(i >> 5) * 64 + (i & 31)] = a.ln_g[i]; sLN[(i >> 5) * 64 + 32 + (i & 31)] = a.ln_b[i]; }
    constexpr int DH = D / H;
    const float qscale = a.scaling ? 1.f : sqrtf((float)DH);      // attention_fwd multiplies the scores by 1/sqrt(dh)
    const uint32_t seed32 = fold_seed(a.seed);
    const int n_tiles = (a.B + spt - 1) / spt;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int b = t * spt + s;
        const bool valid = s < spt && b < a.B;
        const long long tok = ((long long)b * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        int set = 0;
        if (valid && a.dom) set = a.dom[b] - 1;
        const bool active = valid && set >= 0 && set < a.n_wq;     // (inactive samples: zero output, star_trans.py:208,239)
        float x[D], q[D];
        zero(x);
        __syncthreads();
        if (valid) {
            load_row(a.x + tok, x);
            float k[D], v[D];
            zero(q); zero(k); zero(v);
            if (active) vm_outer<D, D, D>(sW + 3072 + set * 1024, x, q);
            vm_outer<D, D, D>(sW, x, k);
            vm_outer<D, D, D>(sW + 1024, x, v);
            sts_row(sK + tid * KV_LD, k);
            sts_row(sV + tid * KV_LD, v);
            if (!a.scaling) {
#pragma unroll
                for (int i = 0; i < D; ++i) q[i] *= qscale;
            }
        }
        __syncthreads();
        if (valid) {
            float r[D];
            zero(r);
            if (active) {
                attention_fwd<H>(q, sK, sV, s * F, F, make_drop(seed32, token_gid, 0, SITE_ATT, a.p_drop, a.training), r);
                const DropCtx dc = make_drop(seed32, token_gid, 0, SITE_OUT, a.p_drop, a.training);
                drop_apply32(dc, r);                                               // result = self.dropout(result)
                if (a.use_res) vm_outer<D, D, D>(sW + 2048, x, r);                  // result += inputs @ W_Res
#pragma unroll
                for (int i = 0; i < D; ++i) r[i] = fmaxf(r[i], 0.f);                // F.relu
                if (a.ln_g) {
                    float mean, rstd;
                    layernorm32(r, sLN + set * 64, sLN + set * 64 + 32, mean, rstd);
                }
            }
            store_row(a.out + tok, r);
        }
    }
}

}  // namespace sat

using namespace sat;

extern "C" int sat_field_attention_fwd(const float* x, float* out, const float* wq, int32_t n_wq, const float* wk, const float* wv,
                                       const float* wres, const float* ln_g, const float* ln_b, const int32_t* dom, int32_t B,
                                       int32_t F, int32_t H, int32_t use_res, int32_t scaling, int32_t training, float dropout_p,
                                       uint64_t seed, void* stream) {
    if (!x || !out || !wq || !wk || !wv || (use_res && !wres) || ((ln_g == nullptr) != (ln_b == nullptr)))
        return sat_host::fail(SAT_EINVAL, "sat_field_attention_fwd: null argument");
    if (F <= 0 || F > SAT_MAXF) return sat_host::fail(SAT_EINVAL, "sat_field_attention_fwd: F=%d not in 1..%d", F, SAT_MAXF);
    if (n_wq < 1 || n_wq > FA_MAX_WQ) return sat_host::fail(SAT_EINVAL, "sat_field_attention_fwd: %d W_Query sets not in 1..%d", n_wq, FA_MAX_WQ);
    if (dropout_p < 0.f || dropout_p >= 1.f) return sat_host::fail(SAT_EINVAL, "sat_field_attention_fwd: dropout_p=%f", dropout_p);
    if (B <= 0) return SAT_OK;
    SelfAttArgs a;
    a.x = x; a.out = out; a.wq = wq; a.wk = wk; a.wv = wv; a.wres = wres; a.ln_g = ln_g; a.ln_b = ln_b; a.dom = dom;
    a.B = B; a.F = F; a.use_res = use_res; a.training = training; a.n_wq = n_wq; a.scaling = scaling; a.p_drop = dropout_p; a.seed = seed;
    const int spt = NT / F;
    const int n_tiles = (B + spt - 1) / spt;
    const int grid = n_tiles < 148 * 2 ? n_tiles : 148 * 2;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = ((3 + n_wq) * 1024 + 2 * NT * KV_LD + n_wq * 64) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_selfatt_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_field_attention_fwd: cannot raise dynamic shared memory to %zu", smem);
        configured = smem;
    }
    switch (H) {
        case 1: k_selfatt_fwd<1><<<grid, NT, smem, s>>>(a); break;
        case 2: k_selfatt_fwd<2><<<grid, NT, smem, s>>>(a); break;
        case 4: k_selfatt_fwd<4><<<grid, NT, smem, s>>>(a); break;
        case 8: k_selfatt_fwd<8><<<grid, NT, smem, s>>>(a); break;
        default: return sat_host::fail(SAT_EINVAL, "sat_field_attention_fwd: heads=%d not in {1,2,4,8}", H);
    }
    return sat_host::check_launch("k_selfatt_fwd");
}

extern "C" int sat_selfatt_fwd(const float* x, float* out, const float* wq, const float* wk, const float* wv,
                               const float* wres, const float* ln_g, const float* ln_b, int32_t B, int32_t F, int32_t H,
                               int32_t use_res, int32_t training, float dropout_p, uint64_t seed, void* stream) {
    if (!ln_g || !ln_b) return sat_host::fail(SAT_EINVAL, "sat_selfatt_fwd: null argument");
    return sat_field_attention_fwd(x, out, wq, 1, wk, wv, wres, ln_g, ln_b, nullptr, B, F, H, use_res, 1, training, dropout_p, seed, stream);
}

// ---- logit head of the baselines: y[b] = sigmoid(bias + sum_i feat[b][i] * w[i] (+ sum_k dense[b][k] * w[n + k])) -------------
__global__ void __launch_bounds__(256) k_head_fwd(const float* __restrict__ feat, int n, const float* __restrict__ dense, int n_dense,
                                                  int ld_dense, const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                  float* __restrict__ y, int apply_sigmoid) {
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * 256 + threadIdx.x) >> 5; b < B; b += (gridDim.x * 256) >> 5) {      // one warp per sample
        float acc = 0.f;
        for (int i = lane; i < n; i += 32) acc = fmaf(feat[(long long)b * n + i], __ldg(w + i), acc);
        for (int k = lane; k < n_dense; k += 32) acc = fmaf(dense[(long long)b * ld_dense + k], __ldg(w + n + k), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const float z = acc + (bias ? __ldg(bias) : 0.f);
            y[b] = apply_sigmoid ? 1.f / (1.f + expf(-z)) : z;
        }
    }
}

extern "C" int sat_head_fwd(const float* feat, int32_t n, const float* dense, int32_t n_dense, int32_t ld_dense, const float* w,
                            const float* bias, int32_t B, float* y, int32_t apply_sigmoid, void* stream) {
    if (!feat || !w || !y || n <= 0 || (n_dense > 0 && !dense)) return sat_host::fail(SAT_EINVAL, "sat_head_fwd: bad argument");
    if (B <= 0) return SAT_OK;
    int blocks = (B + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_head_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(feat, n, dense, n_dense, ld_dense, w, bias, B, y, apply_sigmoid);
    return sat_host::check_launch("k_head_fwd");
}
