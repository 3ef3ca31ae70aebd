// C-ABI plumbing: error reporting and workspace sizing.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "bwd_common.cuh"

long long sat_touched_cap(const SatModel* m, long long cap_tok);      // rows.cu
long long sat_idx_stride(const SatModel* m, long long cap_tok);

namespace sat_host {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SAT_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return SAT_OK;
}

}  // namespace sat_host

extern "C" int sat_abi_version(void) { return SAT_ABI_VERSION; }
extern "C" const char* sat_last_error(void) { return sat_host::g_err; }
extern "C" int sat_struct_sizes(int64_t* sz) {
    if (!sz) return sat_host::fail(SAT_EINVAL, "sat_struct_sizes: null argument");
    sz[0] = sizeof(SatModel); sz[1] = sizeof(SatWork); sz[2] = sizeof(SatAdam);
    return SAT_OK;
}

extern "C" int sat_workspace_sizes(const SatModel* m, int32_t cap_B, int32_t training, int64_t* sz, int32_t* max_tiles_host) {
    if (!m || !sz || cap_B <= 0 || m->F <= 0 || m->F > SAT_MAXF || m->n_shards < 1 || m->n_shards > SAT_MAX_SHARDS)
        return sat_host::fail(SAT_EINVAL, "sat_workspace_sizes: bad argument");
    const int64_t spt = sat::NT / m->F;
    const int64_t max_tiles = (cap_B + spt - 1) / spt + m->n_dom;
    const int64_t spt2 = sat::NT2 / m->F;
    const int64_t max_tiles2 = (cap_B + spt2 - 1) / spt2 + m->n_dom;
    const int64_t n_keys = (m->flags & SAT_FLAG_POS) ? (int64_t)m->n_dom * m->L * 2 : m->n_dom;
    const int64_t tok = (int64_t)cap_B * m->F;
    const int64_t words = (m->local_rows + 31) / 32;
    const int64_t touched_cap = sat_touched_cap(m, tok);                  // rows of this shard any process may touch
    const int64_t n_chunks = (tok + SAT_ROWS_CHUNK - 1) / SAT_ROWS_CHUNK;
    const int64_t sort_blocks = (tok + SAT_SORT_TILE - 1) / SAT_SORT_TILE;
    const bool sharded = m->n_shards > 1;
    int i = 0;
    sz[i++] = cap_B * 4;                                   // perm
    sz[i++] = cap_B * 4;                                   // dom
    sz[i++] = max_tiles * 16;                              // tiles
    sz[i++] = 64 * 4;                                      // counters
    sz[i++] = n_keys * SAT_P * 4;                          // theta
    sz[i++] = training ? n_keys * SAT_P * 4 : 0;           // dtheta
    sz[i++] = training ? 16 * n_keys * 64 * 4 : 0;         // enc_du: one partial per block of k_encoder_bwd (4096 / 256)
    sz[i++] = training ? ((int64_t)m->L * SAT_ACT_PER_LAYER + 1) * tok * SAT_D * 4 + (int64_t)m->L * 2 * tok * 2 * 4 : 0;   // acts + ReLU masks
    sz[i++] = cap_B * 4;                                   // dz
    sz[i++] = training ? tok * SAT_D * 4 : 0;              // dx
    sz[i++] = training ? 3 * tok * SAT_D * 4 : 0;          // dqkv
    sz[i++] = max_tiles2 * 8;                              // loss_part (sized for the finer tiling)
    sz[i++] = training ? m->n_dense_params * 4 : 0;        // dgrad
    // bitmap; row-sharded: [final bitmap of the own shard | staging region per destination shard], sat_stage_words() words each
    sz[i++] = training ? (sharded ? (int64_t)(m->n_shards + 1) * sat_stage_words(m) : words) * 4 : 0;
    sz[i++] = training ? (sharded ? (int64_t)(m->n_shards + 1) * sat_stage_words(m) : words) * 4 : 0;   // word_rank (same regions)
    sz[i++] = training ? ((words + 1023) / 1024 + 1) * 4 : 0;   // block_sums
    sz[i++] = training ? touched_cap * 4 : 0;              // uniq_rows
    sz[i++] = training ? touched_cap * SAT_D * 4 : 0;      // row_grad
    sz[i++] = SAT_REG_PARTS * 8;                           // reg_part
    sz[i++] = 8 * 4;                                       // scalars
    sz[i++] = 2 * ((int64_t)m->L + n_keys) * 8192 * 4;     // wpack (tensor-core B operands: forward + backward blocks)
    sz[i++] = max_tiles2 * 16;                             // tiles2
    sz[i++] = training ? sat::gpart_layout(*m).total * 4 : 0;          // gpart
    sz[i++] = training ? 2 * tok * 4 : 0;                  // sort_keys
    sz[i++] = training ? 2 * tok * 4 : 0;                  // sort_vals
    sz[i++] = training ? (int64_t)(1 << SAT_SORT_RADIX_BITS) * sort_blocks * 4 : 0;   // sort_hist
    sz[i++] = training ? n_chunks * 16 : 0;                // chunk_info
    sz[i++] = training ? n_chunks * 2 * SAT_D * 4 : 0;     // pieces
    sz[i++] = training && sharded ? (int64_t)m->n_shards * tok * SAT_D * 4 : 0;               // recv
    sz[i++] = training && sharded ? (int64_t)m->n_shards * sat_idx_stride(m, tok) * 4 : 0;    // recv_idx (the caller fills it with -1)
    sz[i++] = training ? 32 * 4 : 0;                       // peer_flags
    sz[i++] = 32 * 4;                                      // step_state
    if (max_tiles_host) *max_tiles_host = (int32_t)max_tiles;
    return SAT_OK;
}

// ---- sat_forward: argument marshalling + dispatch on the head count ---------------------------------
extern "C" int sat_forward(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, const float* y, SatWork* w,
                           float* y_pred, int32_t training, float dropout_p, uint64_t seed, void* stream) {
    using namespace sat;
    if (!m || !x || !w || !y_pred) return sat_host::fail(SAT_EINVAL, "sat_forward: null argument");
    if (m->F <= 0 || m->F > SAT_MAXF) return sat_host::fail(SAT_EINVAL, "sat_forward: F=%d not in 1..%d", m->F, SAT_MAXF);
    if (m->n_dense > 64) return sat_host::fail(SAT_EINVAL, "sat_forward: n_dense=%d > 64", m->n_dense);
    if (training && !w->acts) return sat_host::fail(SAT_EINVAL, "sat_forward: training needs w->acts");
    if (dropout_p < 0.f || dropout_p >= 1.f) return sat_host::fail(SAT_EINVAL, "sat_forward: dropout_p=%f", dropout_p);
    if (w->B == 0) return SAT_OK;
    if (w->B > w->cap_B) return sat_host::fail(SAT_EINVAL, "sat_forward: B=%d > workspace capacity %d", w->B, w->cap_B);
    FwdArgs a;
    a.m = *m; a.x = x; a.id_kind = id_kind; a.ld = ld; a.y = y; a.perm = w->perm; a.tiles = w->tiles;
    a.ntiles_ptr = w->counters + CNT_NTILES;
    a.counters = w->counters; a.theta = w->theta; a.acts = w->acts;
    a.act_stride = (long long)w->cap_B * m->F * D;
    a.dz = w->dz; a.loss_part = w->loss_part; a.y_pred = y_pred; a.p_drop = dropout_p; a.seed = seed;
    a.seed_ptr = w->device_step ? reinterpret_cast<const unsigned long long*>(w->step_state) : nullptr;
    a.dx = nullptr; a.gpart = nullptr;
    w->head_bwd_done = 0;
    const int spt = NT / m->F;
    const int max_tiles = (w->B + spt - 1) / spt + m->n_dom;
    const int grid = max_tiles < 148 * 2 ? max_tiles : 148 * 2;
    cudaStream_t s = (cudaStream_t)stream;
    switch (m->H) {
        case 1: return sat_launch_forward_h1(a, grid, training, s);
        case 2: return sat_launch_forward_h2(a, grid, training, s);
        case 4: return sat_launch_forward_h4(a, grid, training, s);
        case 8: return sat_launch_forward_h8(a, grid, training, s);
        default: return sat_host::fail(SAT_EINVAL, "sat_forward: heads=%d not in {1,2,4,8}", m->H);
    }
}

// ---- CUDA IPC: map a peer process' allocation into THIS device's address space (NVLink P2P) ---------------
extern "C" int sat_ipc_open(const void* handle64_host, void** out_ptr_host) {
    if (!handle64_host || !out_ptr_host) return sat_host::fail(SAT_EINVAL, "sat_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return sat_host::fail(SAT_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    *out_ptr_host = p;
    return SAT_OK;
}
extern "C" int sat_ipc_close(void* ptr) {
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return sat_host::fail(SAT_ECUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e)); }
    return SAT_OK;
}

// ---- tensor-core forward: pack the B operands, then the tcgen05 kernel ---------------------------------------
extern "C" int sat_pack_weights(const SatModel* m, SatWork* w, int32_t passes, void* stream) {
    if (!m || !w || !w->theta || !w->wpack) return sat_host::fail(SAT_EINVAL, "sat_pack_weights: null argument");
    if (passes < 1 || passes > 3) return sat_host::fail(SAT_EINVAL, "sat_pack_weights: passes must be 1, 2 or 3, got %d", passes);
    const int n_keys = (m->flags & SAT_FLAG_POS) ? m->n_dom * m->L * 2 : m->n_dom;
    return sat_launch_pack_weights(m, n_keys, w->theta, w->wpack, passes == 2, (cudaStream_t)stream);
}

extern "C" int sat_forward_tc(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, const float* y, SatWork* w,
                              float* y_pred, int32_t training, float dropout_p, uint64_t seed, int32_t passes, void* stream) {
    using namespace sat;
    if (!m || !x || !w || !y_pred || !w->wpack || !w->tiles2) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: null argument");
    if (m->flags & (SAT_FLAG_GATE_Q | SAT_FLAG_GATE_K | SAT_FLAG_BILINEAR))
        return sat_host::fail(SAT_EINVAL, "sat_forward_tc: the gate / bilinear variants run in sat_forward (CUDA-core kernels) only");
    if (m->F <= 0 || m->F > SAT_MAXF) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: F=%d not in 1..%d", m->F, SAT_MAXF);
    if (m->n_dense > 64) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: n_dense=%d > 64", m->n_dense);
    if (training && !w->acts) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: training needs w->acts");
    if (dropout_p < 0.f || dropout_p >= 1.f) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: dropout_p=%f", dropout_p);
    if (w->B == 0) return SAT_OK;
    if (w->B > w->cap_B) return sat_host::fail(SAT_EINVAL, "sat_forward_tc: B=%d > workspace capacity %d", w->B, w->cap_B);
    FwdArgs a;
    a.m = *m; a.x = x; a.id_kind = id_kind; a.ld = ld; a.y = y; a.perm = w->perm; a.tiles = w->tiles2;
    a.ntiles_ptr = w->counters + CNT_NTILES2;
    a.counters = w->counters; a.theta = w->theta; a.acts = w->acts;
    a.act_stride = (long long)w->cap_B * m->F * D;
    a.dz = w->dz; a.loss_part = w->loss_part; a.y_pred = y_pred; a.p_drop = dropout_p; a.seed = seed;
    a.seed_ptr = w->device_step ? reinterpret_cast<const unsigned long long*>(w->step_state) : nullptr;
    // training with labels: the logit head's backward runs in this kernel's epilogue (dLoss/dX_L -> w->dx, head partials)
    const bool fold = training && y && w->dx && w->gpart;
    a.dx = fold ? w->dx : nullptr; a.gpart = fold ? w->gpart : nullptr;
    w->head_bwd_done = fold ? 1 : 0;
    const int spt = 128 / m->F;
    const int max_tiles = (w->B + spt - 1) / spt + m->n_dom;
    const int grid = max_tiles < 148 * 2 ? max_tiles : 148 * 2;
    cudaStream_t s = (cudaStream_t)stream;
    switch (m->H) {
        case 1: return sat_launch_forward_tc_h1(a, w->wpack, grid, training, passes, s);
        case 2: return sat_launch_forward_tc_h2(a, w->wpack, grid, training, passes, s);
        case 4: return sat_launch_forward_tc_h4(a, w->wpack, grid, training, passes, s);
        case 8: return sat_launch_forward_tc_h8(a, w->wpack, grid, training, passes, s);
        default: return sat_host::fail(SAT_EINVAL, "sat_forward_tc: heads=%d not in {1,2,4,8}", m->H);
    }
}
