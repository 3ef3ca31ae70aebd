/*
 * satrans_b200.h — C ABI of the B200-native SATrans hot path (libsatrans_b200.so).
 *
 * The reference (qwerfdsaplking/SATrans) is 100 % Python/PyTorch: it has no FFI / plugin layer,
 * its drop-in boundary is the Python class API.  This header is the native
 * boundary *underneath* that API: every entry point replaces one group of ATen dispatches of the
 * reference's eager path, cited as `reference file:line` on each declaration.  The Python host
 * code in satrans_b200/ binds these symbols with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add) — no torch types cross this boundary: plain device /
 * host pointers, sizes, and a CUDA stream handle passed as `void*`.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SAT_E* code otherwise;
 *     `sat_last_error()` returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - `stream` is a `cudaStream_t` cast to `void*` (NULL = legacy default stream).
 *   - no entry point allocates or frees memory, synchronises the device, or throws.
 *   - fp32 everywhere; embedding dimension d = 32, MetaNet hidden units (64, 32)  [the reference's
 *     "best performance" configurations, readme.md:17-24]; F <= 32 sparse fields; heads in {1,2,4,8}.
 */
#ifndef SATRANS_B200_H
#define SATRANS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAT_ABI_VERSION 4

#define SAT_OK 0
#define SAT_EINVAL (-1)      /* bad argument / unsupported shape            */
#define SAT_ECUDA (-2)       /* a CUDA runtime call or kernel launch failed */

#define SAT_D 32             /* embedding dim                               */
#define SAT_HID 64           /* MetaNet hidden width                        */
#define SAT_P 4096           /* generated weights per key: 32*64 + 64*32    */
#define SAT_MAXF 32          /* max sparse fields (tokens) per sample       */
#define SAT_LAYER_FLOATS 4288 /* packed per-layer parameter block, see SatModel */
#define SAT_REG_PARTS (148 * 256 + 148 * 4) /* doubles in SatWork.reg_part */
#define SAT_WORK_BUFFERS 32  /* pointer members of SatWork */
#define SAT_ACT_PER_LAYER 6  /* activation rows saved per layer and token by a training forward (SatWork.acts) */
#define SAT_GRID_MAX (148 * 2) /* most CTAs any token kernel is launched with: rows of the per-CTA partial buffers */
#define SAT_SORT_RADIX_BITS 9  /* digit width of the touched-row radix sort */
#define SAT_SORT_TILE 2048     /* sorted entries per CTA and radix pass */
#define SAT_ROWS_CHUNK 32      /* sorted entries one 8-lane group reduces */
#define SAT_MAX_SHARDS 8     /* processes (GPUs) of one node sharing the row-sharded tables */
#define SAT_ERR_BAD_ID 1     /* scalars[3] / counters[1] bit: a sparse id was outside its table     */
#define SAT_ERR_BAD_DOMAIN 2 /* ... bit: a scenario id was outside [0, n_dom)                       */

/* id_kind: how sample ids reach the kernels */
#define SAT_IDS_F32 0        /* the reference's (B, n_cols) float32 input matrix; ids = trunc(X[b, col]) */
#define SAT_IDS_I32 1        /* (B, n_cols) int32 matrix                                                */
#define SAT_IDS_I64 2        /* (B, n_cols) int64 matrix                                                */

/* flags for SatModel.flags */
#define SAT_FLAG_POS 1       /* 'pos' in flag: ENP, per-(layer,q/k) generated weights, separate K LayerNorm */
#define SAT_FLAG_META_Q 2    /* 'Q' in meta_mode                                                          */
#define SAT_FLAG_META_K 4    /* 'K' in meta_mode                                                          */
#define SAT_FLAG_RELU 8      /* 'relu' in flag: ReLU on the output projection                             */
#define SAT_FLAG_RES 16      /* att_res: residual connection                                              */
#define SAT_FLAG_LOSS_MSE 32 /* loss = F.mse_loss(y_pred, y, 'sum') instead of binary cross-entropy (meta_basemodel.py:627-640) */
#define SAT_FLAG_LOSS_MAE 64 /* loss = F.l1_loss(y_pred, y, 'sum')                                         */
/* The reference's ablation variants of the scenario-specific projection (satrans.py:59-81,158-176).  GATE_* and BILINEAR run
 * in the CUDA-core token kernels only (sat_forward / sat_backward; sat_forward_tc / sat_backward_tc refuse them); they replace
 * the MetaNet of that side (SAT_FLAG_META_Q / _K are then clear), so Q' / K' carry no MetaNet LayerNorm and no dropout. */
#define SAT_FLAG_GATE_Q 128  /* 'gate' in flag and 'Q' in meta_mode: Q' = (x W_Q) * theta_Q[0:32] * 2     (satrans.py:61-62) */
#define SAT_FLAG_GATE_K 256  /* 'gate' in flag and 'K' in meta_mode: K' = (x W_K) * theta_K[0:32] * 2     (satrans.py:68-69) */
#define SAT_FLAG_BILINEAR 512 /* 'bilinear' in flag: per head h, Q'_h = (x W_Q)_h @ B_h with B_h[i][j] = theta_Q[(h*dh + i)*dh + j],
                                dh = 32 / H; K' = x W_K                                                  (satrans.py:63-64,79-81) */
#define SAT_FLAG_ONLYEMB 1024 /* 'onlyemb' in flag: no scenario encoder; domain_embeddings.weight is [n_dom][4096] and
                                theta[key] = relu(domain_embeddings[g(key)])                             (satrans.py:173-176) */
/* generated weights per key that are live: 32 (gate), 1024 / H (bilinear), 4096 (MetaNet); rows of theta keep the stride SAT_P */
#define SAT_META_PSZ(flags, H) (((flags) & (SAT_FLAG_GATE_Q | SAT_FLAG_GATE_K)) ? 32 : ((flags) & SAT_FLAG_BILINEAR) ? 1024 / (H) : SAT_P)

/*
 * Model description.  All dense (non-table) parameters live in ONE flat fp32 buffer `dense`
 * (gradients / Adam moments use the same layout), at these float offsets:
 *
 *   off_dom   : domain_embeddings.weight            [n_dom][32]          reference satrans.py:145
 *                                                   [n_dom][4096] with SAT_FLAG_ONLYEMB (then off_encw / off_encb are unused)
 *   off_lid   : layerid_embeddings.weight           [L][32]   (POS only) reference satrans.py:169
 *   off_qkv   : qkvid_embeddings.weight             [3][32]   (POS only) reference satrans.py:170
 *   off_encw  : domain_map_dnn_Q.linears.0.weight   [P][enc_in]          reference satrans.py:178, submodules.py:31
 *   off_encb  : domain_map_dnn_Q.linears.0.bias     [P]                  P = SAT_META_PSZ(flags, H): 4096 unless gate / bilinear
 *   off_layers: L blocks of SAT_LAYER_FLOATS floats:                     reference satrans.py:31-47
 *                 +0    W_Query [32][32]  (x @ W)
 *                 +1024 W_Key   [32][32]
 *                 +2048 W_Value [32][32]
 *                 +3072 Out_linear.weight [32 out][32 in]  (x @ W^T)
 *                 +4096 layer_norm.weight[32]  +4128 layer_norm.bias[32]
 *                 +4160 Q_meta_mlp.ffn_layer_norm.weight[32]  +4192 .bias[32]
 *                 +4224 K_meta_mlp.ffn_layer_norm.weight[32]  +4256 .bias[32]   (used only with POS;
 *                        without POS the K side reads/accumulates the Q block — reference satrans.py:44-46)
 *   off_headw : dnn_linear.weight [F*32 + n_dense]                        reference satrans.py:183
 *   off_headb : dnn_linear.bias   [1]
 */
typedef struct SatModel {
    int32_t F;            /* sparse fields = tokens per sample                              */
    int32_t L;            /* domain_att_layer_num                                           */
    int32_t H;            /* att_head_num                                                   */
    int32_t n_dom;        /* rows of domain_embeddings = num_domains + 1                    */
    int32_t n_dense;      /* DenseFeat columns appended to the head input                   */
    int32_t n_cols;       /* columns of the input matrix                                    */
    int32_t flags;        /* SAT_FLAG_*                                                     */
    int32_t enc_in;       /* 32, or 64 with POS                                             */
    int32_t dom_col;      /* column of the scenario id in the input matrix                  */
    int32_t pad0;
    int64_t n_dense_params; /* floats in the flat dense buffer                              */
    int64_t off_dom, off_lid, off_qkv, off_encw, off_encb, off_layers, off_headw, off_headb;
    int64_t total_rows;   /* rows of the concatenated embedding table (all shards)          */
    const int64_t* row_off;   /* [F+1] first row of each field's table in `tables`          */
    const int32_t* sparse_cols; /* [F] input-matrix column of each sparse field             */
    const int32_t* dense_cols;  /* [n_dense] input-matrix column of each dense feature      */
    float* tables;        /* [local_rows][32] this process' rows of the concatenated table  */
    float* dense;         /* flat dense parameter buffer                                    */
    /* Row sharding across the R processes of one node (R = 1: everything local).  Global row g of the
     * concatenated table lives on shard g % R at local row g / R.  Every shard_* array holds the R peer-mapped
     * base pointers (CUDA IPC, NVLink P2P); entry [shard_rank] is this process' own buffer. */
    int32_t n_shards;     /* R                                                              */
    int32_t shard_rank;   /* this process                                                   */
    int64_t local_rows;   /* rows held by this process: ceil((total_rows - shard_rank) / R) */
    float* shard_tables[SAT_MAX_SHARDS];
    uint32_t* shard_bitmap[SAT_MAX_SHARDS];     /* SatWork.bitmap of every shard            */
    int32_t* shard_word_rank[SAT_MAX_SHARDS];   /* SatWork.word_rank of every shard         */
    float* shard_row_grad[SAT_MAX_SHARDS];      /* SatWork.row_grad of every shard          */
    float* shard_recv[SAT_MAX_SHARDS];          /* SatWork.recv of every shard              */
    int32_t* shard_recv_idx[SAT_MAX_SHARDS];    /* SatWork.recv_idx of every shard          */
    uint32_t* shard_flags[SAT_MAX_SHARDS];      /* SatWork.peer_flags of every shard        */
} SatModel;

/* Per-batch workspace, caller-allocated (sizes from sat_workspace_sizes).  */
typedef struct SatWork {
    int32_t B;            /* samples in this batch                                          */
    int32_t cap_B;        /* capacity the buffers were sized for                            */
    int32_t device_step;  /* != 0: the dropout seed and the optimiser's step-dependent constants come from step_state
                             (advanced on the device by sat_step_begin) instead of the by-value arguments, so that a
                             whole training step can be captured ONCE in a CUDA graph and replayed            */
    int32_t defer_finish; /* != 0: sat_backward* stop after the layer kernels; the caller runs sat_backward_finish (fixed-order
                             reduction of the per-CTA partials + encoder backward) later, e.g. on a second stream beside the
                             HBM-bound table sweep                                                                    */
    int32_t pad2;
    int32_t head_bwd_done; /* set by sat_forward_tc when it also ran the logit head's backward (training, labels given); read by
                             the next sat_backward*, which then skips its own head kernel; cleared by sat_forward      */
    int32_t* perm;        /* [cap_B] sorted position -> sample index (bucketed by scenario) */
    int32_t* dom;         /* [cap_B] scenario id per sorted position                        */
    int32_t* tiles;       /* [max_tiles][4] (key/domain, start, count, pad)                 */
    int32_t* counters;    /* [64] ints: [0]=n_tiles [1]=error flags [2]=n_unique [3]=n_tiles2 [8..] histogram */
    float* theta;         /* [n_keys][4096] generated MetaNet weights per key               */
    float* dtheta;        /* [n_keys][4096] their gradient                                  */
    float* enc_du;        /* [n_keys][64] gradient w.r.t. the encoder inputs                */
    float* acts;          /* training only: [L][6][cap_B*F][32] saved x, zq, zk, v, r, y + [cap_B*F][32] final x + ReLU masks */
    float* dz;            /* [cap_B] dLoss/dlogit per sorted position                       */
    float* dx;            /* [2][cap_B*F][32] ping-pong activation gradients                */
    float* dqkv;          /* [3][cap_B*F][32] dq', dk', dv between the two backward kernels */
    double* loss_part;    /* [max_tiles] per-tile BCE partial sums                          */
    float* dgrad;         /* flat dense-parameter gradient buffer (n_dense_params)          */
    /* touched-row bookkeeping for the embedding update */
    uint32_t* bitmap;     /* [ceil(local_rows/32)] one bit per table row; row-sharded: (n_shards + 1) regions of
                             sat_stage_words() words: the shard's final bitmap, then this process' marks per destination shard */
    int32_t* word_rank;   /* [ceil(local_rows/32)] exclusive prefix popcount; row-sharded: (n_shards + 1) regions like the
                             bitmap: the own ranks, then mirrors of every shard's ranks (and, in the bitmap's regions, of every
                             shard's final bitmap) pulled with bulk NVLink loads before the token sort                  */
    int32_t* block_sums;  /* [ceil(words/1024)+1]                                           */
    int32_t* uniq_rows;   /* [cap_B*F] sorted unique touched rows                           */
    float* row_grad;      /* [cap_B*F][32] accumulated dE per unique touched row            */
    double* reg_part;     /* [SAT_REG_PARTS] partial sums of squares of the tables          */
    float* scalars;       /* [8] device results: [0]=bce sum [1]=reg loss [2]=n_unique rows [3]=error flags */
    float* wpack;         /* [(L + n_keys)][8192] weights packed as tcgen05 B operands (hi | lo)         */
    int32_t* tiles2;      /* tile list of the tensor-core kernels (128 tokens per tile)                  */
    /* deterministic gradient accumulation (no floating-point atomics anywhere): */
    float* gpart;         /* per-CTA partial dense gradients, reduced in CTA order by sat_backward*:
                             [L][SAT_GRID_MAX][1088] dW_out + d layer_norm | [L][SAT_GRID_MAX][3200] dW_q/k/v + d MetaNet-LN |
                             [L][SAT_GRID_MAX][1 or 2][4096] d theta | [SAT_GRID_MAX][head_ld] d dnn_linear        */
    uint32_t* sort_keys;  /* [2][cap_B*F] radix-sort ping-pong: key = (owner shard << slot_bits) | touched-row slot  */
    int32_t* sort_vals;   /* [2][cap_B*F] ... value = token (sorted position * F + field)                           */
    int32_t* sort_hist;   /* [2^SAT_SORT_RADIX_BITS][ceil(cap_B*F / SAT_SORT_TILE)] digit histogram / offsets       */
    int32_t* chunk_info;  /* [ceil(cap_B*F / SAT_ROWS_CHUNK)][4] run bookkeeping of the segmented reduce            */
    float* pieces;        /* [chunks][2][32] partial sums of runs that cross a chunk boundary                        */
    /* row-sharded mode only (n_shards > 1): every source shard pushes ONE pre-reduced gradient per unique row */
    float* recv;          /* [n_shards][cap_B*F][32] rows pushed by source shard s (filled by the peers)             */
    int32_t* recv_idx;    /* [n_shards][touched rows cap] slot -> index into recv[s], -1 = nothing (filled by peers) */
    uint32_t* peer_flags; /* [32] words: [0..8) epoch flags written by the peers, [8] own epoch, [9] != 0: sat_peer_barrier gave
                             up waiting for peer [9]-1 after 20 s, [16..24) push cursors                                  */
    float* step_state;    /* [32] words, see sat_step_begin: [0..1] u64 seed, [2] i32 step, [4..7] f32 lr, beta1, beta2, eps (the
                             caller copies them in, e.g. from pinned memory inside the graph), [8..15] derived constants  */
} SatWork;

/* words per region of the row-sharded SatWork.bitmap: the largest shard's rows / 32, the same number on every process */
#ifdef __CUDACC__
__host__ __device__
#endif
static inline int64_t sat_stage_words(const SatModel* m) {
    return ((((m->total_rows + m->n_shards - 1) / m->n_shards + 31) / 32) + 3) & ~(int64_t)3;      /* (multiple of 16 bytes) */
}

/* optimiser kinds (torch.optim defaults; the strings BaseModel._get_optim accepts, reference meta_basemodel.py:612-625) */
#define SAT_OPT_ADAM 0      /* m, v */
#define SAT_OPT_SGD 1       /* p -= lr*g                                  (no state)               */
#define SAT_OPT_ADAGRAD 2   /* v += g*g; p -= lr*g/(sqrt(v)+eps)          (v = state_sum)          */
#define SAT_OPT_RMSPROP 3   /* v = a*v+(1-a)*g*g; p -= lr*g/(sqrt(v)+eps) (v = square_avg, a=beta2) */

typedef struct SatAdam {
    float lr, beta1, beta2, eps;
    float l2_emb;         /* l2_reg_embedding (reference satrans.py:117, meta_basemodel.py:179) */
    int32_t step;         /* 1-based step index of THIS update                                  */
    int32_t dense_mode;   /* 1 = reference semantics (every table row gets L2 + Adam every step,
                             reference meta_basemodel.py:577-593 + main.py:343);
                             0 = touched rows only (lazy/sparse Adam)                           */
    int32_t kind;         /* SAT_OPT_*                                                           */
    float* tab_m; float* tab_v;       /* [local_rows][32] optimiser state of the tables (unused ones may be NULL) */
    float* dense_m; float* dense_v;   /* [n_dense_params]                                       */
    /* deferred table sweep (sat_defer_*; NULL when unused): */
    uint8_t* row_upto;    /* [local_rows] steps of the current sweep epoch already applied to the row             */
    float* defer_state;   /* [SAT_DEFER_STATE_WORDS] [0] i32 steps begun in this epoch, [1] i32 overflow flag,
                             [8 + 16 j ..] the optimiser constants of the epoch's step j                        */
    double* defer_reg;    /* [SAT_DEFER_REG_DOUBLES] per-step partial sums of squares, see sat_defer_sweep        */
} SatAdam;

/* ---- library info ------------------------------------------------------------------------- */
int sat_abi_version(void);
const char* sat_last_error(void);
/* sizes_host[0..2] = sizeof(SatModel), sizeof(SatWork), sizeof(SatAdam) as THIS library was compiled (a binding checks its
 * struct mirrors against them). */
int sat_struct_sizes(int64_t* sizes_host);
/* bytes needed for each SatWork buffer, in the order of the struct's pointer members
 * (SAT_WORK_BUFFERS entries; 0 = not needed for inference-only use). */
int sat_workspace_sizes(const SatModel* m, int32_t cap_B, int32_t training, int64_t* sizes_host, int32_t* max_tiles_host);

/* ---- embedding gather: replaces BaseModel.input_from_feature_columns (reference
 *      meta_basemodel.py:519-545: F slices + F float->long casts + F nn.Embedding calls) and the
 *      torch.cat at satrans.py:211.  out[b][f][:] = tables[row_off[f] + trunc(ids[b][col_f])][:].
 *      `x` is the (B, ld) input matrix in the layout given by id_kind.  Out-of-range ids set bit 0
 *      of *err_flag (device int, may be NULL) and read row 0 of that field. ------------------- */
int sat_gather_fwd(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, int32_t B,
                   float* out, int32_t* err_flag, void* stream);

/* ---- scenario bucketing: samples grouped by trunc(X[b, dom_col]) (reference satrans.py:203) so that
 *      every tile of the fused kernels shares one set of generated MetaNet weights. ------------- */
int sat_bucket_by_domain(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w,
                         int32_t samples_per_tile, void* stream);

/* ---- scenario encoder: replaces domain_embeddings + ReLU + DNN_v2 per sample (reference
 *      satrans.py:203-234, submodules.py:47-61) by one evaluation per distinct key. ------------- */
int sat_encoder_fwd(const SatModel* m, SatWork* w, void* stream);
int sat_encoder_bwd(const SatModel* m, SatWork* w, void* stream);

/* ---- fused forward: gather + L x Meta_Transformer_Layer + logit head + sigmoid (+ BCE and dLogit
 *      when y != NULL).  Replaces SATrans.forward (reference satrans.py:197-256), Meta_Transformer_Layer
 *      .forward (satrans.py:50-100), MetaNet.forward (submodules.py:77-103) and, with y, the BCE of
 *      meta_basemodel.py:317.  y_pred is (B,) in ORIGINAL sample order.  training != 0 saves the
 *      activations the backward needs and applies dropout(p) with a counter-based RNG keyed by
 *      (seed, sample, field, layer, site).  Requires sat_bucket_by_domain + sat_encoder_fwd first. */
int sat_forward(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, const float* y,
                SatWork* w, float* y_pred, int32_t training, float dropout_p, uint64_t seed, void* stream);

/* ---- fused backward of the L layers + head; writes the dense-parameter gradient into w->dgrad,
 *      dtheta, and leaves dLoss/dE0 per token in w->dx (sat_rows_reduce sums it per touched table row).
 *      Every reduction has a fixed order: CTAs write partial sums into w->gpart and one kernel adds
 *      them in CTA order, so two runs give bit-identical gradients.  Replaces autograd over the graph
 *      of satrans.py:197-256. ------------------------------------------------------------------- */
int sat_backward(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w,
                 float dropout_p, uint64_t seed, void* stream);

/* the tail of sat_backward / sat_backward_tc when w->defer_finish was set: per-CTA partials -> w->dgrad and w->dtheta, then the
 * scenario-encoder backward.  passes: 0 after sat_backward, the value given to sat_backward_tc otherwise. */
int sat_backward_finish(const SatModel* m, SatWork* w, int32_t passes, void* stream);

/* ---- touched-row bookkeeping.  sat_rows_clear zeroes this shard's bitmap; sat_rows_mark sets one bit per
 *      row the batch reads IN THE OWNING SHARD's bitmap (remote atomics over NVLink when R > 1);
 *      sat_rows_rank turns this shard's bitmap into ranks + the sorted unique row list.  With R > 1 the caller
 *      separates the three calls with a cross-process barrier; sat_rows_mark_and_rank does all three (R = 1). */
int sat_rows_clear(const SatModel* m, SatWork* w, void* stream);
int sat_rows_mark(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, void* stream);
int sat_rows_rank(const SatModel* m, SatWork* w, void* stream);
int sat_rows_mark_and_rank(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, void* stream);
/* ---- sorted-index scatter-add of the table-row gradients (replaces embedding_dense_backward of the reference's
 *      autograd, meta_basemodel.py:533-535 backward).  sat_rows_sort (after sat_bucket_by_domain and the touched-row
 *      ranks; ids only, so it can run beside the forward) radix-sorts the batch's tokens by (owner shard, touched-row
 *      slot), stable in the token index.  sat_rows_reduce (after sat_backward*) sums w->dx over each run of equal
 *      keys in that order -- fixed chunks of SAT_ROWS_CHUNK sorted entries, chunk partials added in chunk order --
 *      into w->row_grad[slot] (n_shards == 1) or pushes ONE row per unique (source shard, row) into the owner's
 *      w->recv over NVLink; sat_rows_combine (row-sharded mode, after a cross-process barrier) adds the pushed rows
 *      in source-shard order into w->row_grad.  No floating-point atomics: results do not depend on timing. ------- */
int sat_rows_sort(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, void* stream);
int sat_rows_reduce(const SatModel* m, SatWork* w, void* stream);
int sat_rows_combine(const SatModel* m, SatWork* w, void* stream);
/* Stream-ordered barrier across the n_shards processes through the peer-mapped w->peer_flags (one tiny kernel: bump the
 * own epoch, store it into every peer's flag word with system scope, spin until every peer's epoch arrived).  Every
 * process must call it the same number of times. */
int sat_peer_barrier(const SatModel* m, SatWork* w, void* stream);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device (idempotent). */
int sat_enable_peer_access(int32_t peer_device);
/* Map a peer process' allocation (64-byte cudaIpcMemHandle_t, host memory) into the CURRENT device's address
 * space with peer access over NVLink; *out_ptr_host receives the base pointer of the exporter's allocation. */
int sat_ipc_open(const void* handle64_host, void** out_ptr_host);
int sat_ipc_close(void* ptr);

/* ---- device-side step state (CUDA-graph replay).  One 1-thread kernel: step += 1; seed = seed_base + step; the optimiser
 *      constants of THIS step (bias corrections in double precision, like torch.optim.Adam) are derived from the
 *      hyper-parameters lr, beta1, beta2, eps found in w->step_state[4..7].  Kernels launched with w->device_step != 0 read
 *      them from there.  sat_step_set writes the step counter and the hyper-parameters from the host (stream-ordered). ---- */
int sat_step_begin(SatWork* w, uint64_t seed_base, void* stream);
int sat_step_set(SatWork* w, int32_t step, float lr, float beta1, float beta2, float eps, void* stream);

/* ---- optimiser: replaces get_regularization_loss (reference meta_basemodel.py:577-593), the dense
 *      embedding_dense_backward and torch.optim.Adam.step (main.py:343, meta_basemodel.py:328). ---- */
/* which = SAT_ADAM_TOUCHED (rows read by the batch: g = dE + 2*l2*w), SAT_ADAM_UNTOUCHED (every other row:
 * g = 2*l2*w; skipped when dense_mode == 0) or both.  The two kernels write disjoint rows, so they may be
 * issued on different streams and run concurrently with each other and with forward/backward. */
#define SAT_ADAM_TOUCHED 1
#define SAT_ADAM_UNTOUCHED 2
int sat_tables_adam(const SatModel* m, SatWork* w, const SatAdam* a, int32_t which, void* stream);

/* ---- deferred table sweep: the reference's every-row-every-step update (meta_basemodel.py:577-593 + main.py:343) with the
 *      rows NOT read by a batch brought up to date in batches of up to SAT_DEFER_MAX steps.  A row the batch does not read gets
 *      g = 2*l2*w, which depends on nothing but the row itself and the step's constants, so applying its steps later -- one
 *      after the other, same arithmetic, the constants of every step kept in a ring -- gives bit-identical (p, m, v); the table
 *      state then crosses HBM once per sweep epoch instead of once per step.  Order inside one training step:
 *        sat_defer_begin                     ring[n] = this step's constants, n += 1            (before anything else)
 *        sat_defer_catchup                   after sat_rows_rank, BEFORE the forward: the rows this batch reads (w->uniq_rows)
 *                                            get the steps of the epoch they have not seen yet
 *        sat_tables_adam(SAT_ADAM_TOUCHED | SAT_ADAM_DEFERRED)   as before, and marks those rows as up to date
 *      and after at most SAT_DEFER_MAX steps (or before anything else reads the tables)
 *        sat_defer_sweep                     every row up to date, epoch closed (n = 0); a->defer_reg[0..n) receives
 *                                            l2 * sum(table^2) as each step of the epoch saw it (the regulariser term of
 *                                            its loss), [SAT_DEFER_MAX] their running total, [SAT_DEFER_MAX + 1] the count.
 *      No floating-point atomics: every partial sum has its own slot and is added in a fixed order. ---------------------- */
#define SAT_DEFER_MAX 16
#define SAT_ADAM_DEFERRED 4
#define SAT_DEFER_STATE_WORDS (8 + 16 * SAT_DEFER_MAX)
#define SAT_DEFER_ROWS_BLOCKS (148 * 4)
#define SAT_DEFER_SWEEP_BLOCKS (148 * 64)
/* doubles: [0..18) results | touch [MAX][ROWS_BLOCKS] | catch-up [MAX launches][ROWS_BLOCKS][MAX] | sweep [SWEEP_BLOCKS][MAX] */
#define SAT_DEFER_REG_DOUBLES (32 + SAT_DEFER_MAX * SAT_DEFER_ROWS_BLOCKS + SAT_DEFER_MAX * SAT_DEFER_ROWS_BLOCKS * SAT_DEFER_MAX + \
                               SAT_DEFER_SWEEP_BLOCKS * SAT_DEFER_MAX)
int sat_defer_begin(const SatModel* m, SatWork* w, const SatAdam* a, void* stream);
int sat_defer_catchup(const SatModel* m, SatWork* w, const SatAdam* a, void* stream);
int sat_defer_sweep(const SatModel* m, const SatAdam* a, void* stream);
int sat_dense_adam(const SatModel* m, SatWork* w, const SatAdam* a, void* stream);
/* sum of squares of all table rows into w->reg_part (regulariser without an optimiser step). */
int sat_tables_sumsq(const SatModel* m, SatWork* w, void* stream);
/* w->scalars[0] = BCE(sum) of the batch, [1] = l2_emb * sum(table^2), [3] = error flags.
 * tile_set: 0 if the last forward was sat_forward, 1 if it was sat_forward_tc. */
int sat_finalize_loss(const SatModel* m, SatWork* w, float l2_emb, int32_t tile_set, void* stream);

/* ---- tensor-core forward (tcgen05 / TMEM): same contract as sat_forward.  sat_pack_weights (after
 *      sat_encoder_fwd) packs the layer weights and the generated MetaNet weights as UMMA B operands;
 *      passes = 3: error-compensated 3xTF32 (fp32-level parity), passes = 1: plain TF32, passes = 2: bf16 operands with fp32
 *      accumulation (kind::f16; the packing must have been done with the same value). ---------------------------- */
int sat_pack_weights(const SatModel* m, SatWork* w, int32_t passes, void* stream);
int sat_forward_tc(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, const float* y, SatWork* w,
                   float* y_pred, int32_t training, float dropout_p, uint64_t seed, int32_t passes, void* stream);
/* sat_backward with the MetaNet / QKV data-gradient projections on the tensor cores (same contract, same outputs). */
int sat_backward_tc(const SatModel* m, const void* x, int32_t id_kind, int32_t ld, SatWork* w, float dropout_p,
                    uint64_t seed, int32_t passes, void* stream);

/* ---- "next" row: SelfAttention_Layer.forward (reference submodules.py:178-238) through the same
 *      per-sample field-attention code.  x, out: [B][F][32]; W_*: [32][32] used as x @ W; forward.    ---------- */
int sat_selfatt_fwd(const float* x, float* out, const float* wq, const float* wk, const float* wv, const float* wres,
                    const float* ln_g, const float* ln_b, int32_t B, int32_t F, int32_t H, int32_t use_res,
                    int32_t training, float dropout_p, uint64_t seed, void* stream);

/* ---- the general form: SelfAttention_Layer (n_wq = 1, ln given, scaling = 1), MDR_InteractingLayer (reference
 *      star_trans.py:146-242: wq / ln_g / ln_b hold n_wq <= 8 scenario-specific sets, dom[b] in 1..n_wq selects set dom-1, any
 *      other id leaves a zero output; the 'sum' / 'dot' merge with the shared W_Query is done by the caller) and deepctr-torch's
 *      InteractingLayer (AutoInt: ln_g = ln_b = NULL, scaling = 0, dropout_p = 0).  Forward. --------------------------------- */
int sat_field_attention_fwd(const float* x, float* out, const float* wq, int32_t n_wq, const float* wk, const float* wv,
                            const float* wres, const float* ln_g, const float* ln_b, const int32_t* dom, int32_t B, int32_t F,
                            int32_t H, int32_t use_res, int32_t scaling, int32_t training, float dropout_p, uint64_t seed,
                            void* stream);
/* ---- backward of the same layer with ONE W_Query / LayerNorm set (SelfAttention_Layer: ln_g given, scaling = 1;
 *      InteractingLayer: ln_g = NULL): replaces torch autograd over reference submodules.py:208-238.  Only x is needed from the
 *      forward (it is recomputed per tile with the same seed, so the dropout masks are the forward's).  sample_idx = NULL: the
 *      launch works on samples 0..B-1.  sample_idx = B sample indices (device int32): it works on those samples only -- rows of
 *      x / dy / dx and the dropout counters are those of sample sample_idx[i], every other row of dx is left untouched.  That is
 *      one scenario of an MDR_InteractingLayer (reference star_trans.py:218-239 under autograd): the caller launches once per
 *      scenario with that scenario's W_Query / LayerNorm and adds the shared-parameter gradients in scenario order.  dx rows
 *      and the parameter gradients are OVERWRITTEN (not accumulated); every sum has a fixed order.
 *      scratch: sat_field_attention_bwd_scratch(B, F) bytes. ------------------------------------------------------------- */
int64_t sat_field_attention_bwd_scratch(int32_t B, int32_t F);
int sat_field_attention_bwd(const float* x, const float* dy, float* dx, const float* wq, const float* wk, const float* wv,
                            const float* wres, const float* ln_g, const int32_t* sample_idx, int32_t B, int32_t F, int32_t H,
                            int32_t use_res, int32_t scaling, int32_t training, float dropout_p, uint64_t seed, float* dwq,
                            float* dwk, float* dwv, float* dwres, float* dln_g, float* dln_b, float* scratch, void* stream);
/* ---- logit head of the baselines (reference star_trans.py:762-775, autoint.py:107-121): y[b] = sigmoid(bias + feat[b] . w[0:n]
 *      + dense[b] . w[n:n+n_dense]); feat: [B][n], dense: [B][ld_dense] or NULL, bias may be NULL, apply_sigmoid = 0 -> logit. */
int sat_head_fwd(const float* feat, int32_t n, const float* dense, int32_t n_dense, int32_t ld_dense, const float* w,
                 const float* bias, int32_t B, float* y, int32_t apply_sigmoid, void* stream);

/* ---- per-batch training metrics of fit(verbose > 0) (reference meta_basemodel.py:330-337,352-356; metric table :655-671).
 *      y_pred, y: device arrays; batch b covers [batch_off[b], batch_off[b+1]) (device int32, n_batches + 1 entries);
 *      out[b][4] = { log_loss, roc_auc_score, mean_squared_error, accuracy(p > 0.5) } as sklearn computes them in float64
 *      (NaN logloss / auc for a single-class batch).  One CTA per batch, max_batch <= 16384 samples. ---------------------- */
int sat_batch_metrics(const float* y_pred, const float* y, const int32_t* batch_off, int32_t n_batches, int32_t max_batch,
                      double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SATRANS_B200_H */
