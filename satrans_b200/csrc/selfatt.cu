// SelfAttention_Layer forward (reference submodules.py:178-238) through the same per-sample field-attention code
// as SATrans: "next" row 1 of SURVEY.md §8f (the AutoInt / STAR-Trans style baselines of config 5).
//   Q,K,V = X W_{Q,K,V};  A = dropout(softmax(Q K^T / sqrt(dh)));  R = dropout(A V);
//   out = LayerNorm(relu(R + X W_Res))        (W_Out is declared by the reference but never used)
// One thread = one field token, a CTA = floor(192/F) whole samples; forward only (inference and train-mode dropout).
#include "fwd_common.cuh"

namespace sat {

struct SelfAttArgs {
    const float* x;        // [B][F][32]
    float* out;            // [B][F][32]
    const float* wq; const float* wk; const float* wv; const float* wres;   // [32][32] each, used as x @ W
    const float* ln_g; const float* ln_b;
    int B, F, use_res, training;
    float p_drop;
    unsigned long long seed;
};

template <int H>
__global__ void __launch_bounds__(NT, 2) k_selfatt_fwd(const SelfAttArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                       // W_Query | W_Key | W_Value | W_Res
    float* sK = sW + 4 * 1024;              // [NT][36]
    float* sV = sK + NT * KV_LD;            // [NT][36]
    float* sLN = sV + NT * KV_LD;           // gamma | beta
    const int tid = threadIdx.x, F = a.F;
    const int spt = NT / F;
    const int s = tid / F, f = tid - s * F;
    copy_to_smem(sW, a.wq, 1024, tid, NT);
    copy_to_smem(sW + 1024, a.wk, 1024, tid, NT);
    copy_to_smem(sW + 2048, a.wv, 1024, tid, NT);
    if (a.use_res) copy_to_smem(sW + 3072, a.wres, 1024, tid, NT);
    if (tid < 32) { sLN[tid] = a.ln_g[tid]; sLN[32 + tid] = a.ln_b[tid]; }
    const int n_tiles = (a.B + spt - 1) / spt;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int b = t * spt + s;
        const bool valid = s < spt && b < a.B;
        const long long tok = ((long long)b * F + f) * D;
        const uint32_t token_gid = (uint32_t)b * (uint32_t)F + (uint32_t)f;
        float x[D], q[D];
        zero(x);
        __syncthreads();
        if (valid) {
            load_row(a.x + tok, x);
            float k[D], v[D];
            zero(q); zero(k); zero(v);
            vm_outer<D, D, D>(sW, x, q);
            vm_outer<D, D, D>(sW + 1024, x, k);
            vm_outer<D, D, D>(sW + 2048, x, v);
            sts_row(sK + tid * KV_LD, k);
            sts_row(sV + tid * KV_LD, v);
        }
        __syncthreads();
        if (valid) {
            float r[D];
            attention_fwd<H>(q, sK, sV, s * F, F, make_drop(fold_seed(a.seed), token_gid, 0, SITE_ATT, a.p_drop, a.training), r);
            const DropCtx dc = make_drop(fold_seed(a.seed), token_gid, 0, SITE_OUT, a.p_drop, a.training);
#pragma unroll
            drop_apply32(dc, r);                                               // result = self.dropout(result)
            if (a.use_res) vm_outer<D, D, D>(sW + 3072, x, r);                  // result += inputs @ W_Res
#pragma unroll
            for (int i = 0; i < D; ++i) r[i] = fmaxf(r[i], 0.f);                // F.relu
            float mean, rstd;
            layernorm32(r, sLN, sLN + 32, mean, rstd);
            store_row(a.out + tok, r);
        }
    }
}

}  // namespace sat

using namespace sat;

extern "C" int sat_selfatt_fwd(const float* x, float* out, const float* wq, const float* wk, const float* wv,
                               const float* wres, const float* ln_g, const float* ln_b, int32_t B, int32_t F, int32_t H,
                               int32_t use_res, int32_t training, float dropout_p, uint64_t seed, void* stream) {
    if (!x || !out || !wq || !wk || !wv || !ln_g || !ln_b || (use_res && !wres))
        return sat_host::fail(SAT_EINVAL, "sat_selfatt_fwd: null argument");
    if (F <= 0 || F > SAT_MAXF) return sat_host::fail(SAT_EINVAL, "sat_selfatt_fwd: F=%d not in 1..%d", F, SAT_MAXF);
    if (B <= 0) return SAT_OK;
    SelfAttArgs a;
    a.x = x; a.out = out; a.wq = wq; a.wk = wk; a.wv = wv; a.wres = wres; a.ln_g = ln_g; a.ln_b = ln_b;
    a.B = B; a.F = F; a.use_res = use_res; a.training = training; a.p_drop = dropout_p; a.seed = seed;
    const int spt = NT / F;
    const int n_tiles = (B + spt - 1) / spt;
    const int grid = n_tiles < 148 * 2 ? n_tiles : 148 * 2;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (4 * 1024 + 2 * NT * KV_LD + 64) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_selfatt_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaFuncSetAttribute(k_selfatt_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return sat_host::fail(SAT_ECUDA, "sat_selfatt_fwd: cannot raise dynamic shared memory to %zu", smem);
        configured = true;
    }
    switch (H) {
        case 1: k_selfatt_fwd<1><<<grid, NT, smem, s>>>(a); break;
        case 2: k_selfatt_fwd<2><<<grid, NT, smem, s>>>(a); break;
        case 4: k_selfatt_fwd<4><<<grid, NT, smem, s>>>(a); break;
        case 8: k_selfatt_fwd<8><<<grid, NT, smem, s>>>(a); break;
        default: return sat_host::fail(SAT_EINVAL, "sat_selfatt_fwd: heads=%d not in {1,2,4,8}", H);
    }
    return sat_host::check_launch("k_selfatt_fwd");
}
