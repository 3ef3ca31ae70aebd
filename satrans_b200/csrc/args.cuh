// Argument blocks shared by the per-head-count instantiations of the token kernels and their dispatchers.
#pragma once
#include "common.cuh"

namespace sat {

struct FwdArgs {
    SatModel m;
    const void* x;
    int id_kind, ld;
    const float* y;            // labels in original sample order, or NULL
    const int32_t* perm;
    const int32_t* tiles;
    const int32_t* ntiles_ptr; // number of tiles in `tiles`
    int32_t* counters;
    const float* theta;
    float* acts;               // [L][5][act_stride] + [act_stride]
    long long act_stride;      // cap_B * F * 32 floats
    float* dz;
    double* loss_part;
    float* y_pred;
    float p_drop;
    unsigned long long seed;
    const unsigned long long* seed_ptr;   // != NULL: the seed lives in SatWork.step_state (CUDA-graph replay)
    float* dx;                 // tensor-core forward with labels: receives dLoss/dX_L (the logit head's backward is folded in)
    float* gpart;              // ... and the head's parameter-gradient partials (gpart_layout in bwd_common.cuh)
};

// saved per layer and token (training): layer input x, pre-LayerNorm MetaNet outputs z_q / z_k, v, attention output r, and the
// pre-LayerNorm layer output y = drop(act(r W_out^T)) + x (so that the backward does not recompute the output projection)
enum { ACT_X = 0, ACT_ZQ = 1, ACT_ZK = 2, ACT_V = 3, ACT_R = 4, ACT_Y = 5, ACT_PER_LAYER = SAT_ACT_PER_LAYER };
static_assert(ACT_PER_LAYER == 6, "activation slots");

// ReLU decisions of the MetaNet hidden layer (64 bits per token, layer and side), saved by the forward so that the
// backward differentiates exactly the function the forward evaluated (the forward may run on the tensor cores).
// They live behind the activations: acts + (L*ACT_PER_LAYER+1)*act_stride, as uint32 [L][2 sides][tokens][2].
__device__ __forceinline__ uint32_t* relu_mask_ptr(float* acts, long long act_stride, int L, int layer, int side, long long tok) {
    uint32_t* base = reinterpret_cast<uint32_t*>(acts + ((long long)L * ACT_PER_LAYER + 1) * act_stride);
    return base + (((long long)layer * 2 + side) * (act_stride / 32) + tok) * 2;
}

struct BwdArgs {
    SatModel m;
    const void* x;
    int id_kind, ld;
    const int32_t* perm;
    const int32_t* tiles;
    int32_t* counters;
    const float* theta;
    float* dtheta;
    const float* acts;
    long long act_stride;
    const float* dz;
    float* dx;                 // [tok][32] gradient w.r.t. the current layer's output / input (in place)
    float* dqkv;               // [3][act_stride]
    float* dgrad;
    const uint32_t* bitmap;
    const int32_t* word_rank;
    float* row_grad;
    float* gpart;              // per-CTA partial gradient buffers (layout: gpart_layout in bwd_common.cuh)
    int layer;
    float p_drop;
    unsigned long long seed;
    const unsigned long long* seed_ptr;   // != NULL: the seed lives in SatWork.step_state (CUDA-graph replay)
};

}  // namespace sat

// one translation unit per head count H (compiled with -DSAT_INST_H=H); defined in forward.cu / backward_attn.cu
#define SAT_DECL_H(H)                                                                          \
    int sat_launch_forward_h##H(const sat::FwdArgs& a, int grid, int training, cudaStream_t s); \
    int sat_launch_forward_tc_h##H(const sat::FwdArgs& a, const float* wpack, int grid, int training, int passes, cudaStream_t s); \
    int sat_launch_bwd_attn_h##H(const sat::BwdArgs& a, int grid, cudaStream_t s);
SAT_DECL_H(1)
SAT_DECL_H(2)
SAT_DECL_H(4)
SAT_DECL_H(8)
#undef SAT_DECL_H
int sat_launch_pack_weights(const SatModel* m, int n_keys, const float* theta, float* wpack, int bf16, cudaStream_t s);
