// Shared device helpers for the SATrans sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "satrans_b200.h"

namespace sat {

constexpr int D = SAT_D;          // embedding dim
constexpr int HID = SAT_HID;      // MetaNet hidden width
constexpr int PSZ = SAT_P;        // generated weights per key
constexpr int NT = 192;           // threads per CTA of the token kernels (one thread = one field token)
constexpr int KV_LD = 36;         // padded row stride (floats) of [token][36] smem tiles: conflict-free STS.128
constexpr int LAYER_FLOATS = SAT_LAYER_FLOATS;
constexpr int OFF_WQ = 0, OFF_WK = 1024, OFF_WV = 2048, OFF_WO = 3072;
constexpr int OFF_LN_G = 4096, OFF_LN_B = 4128, OFF_QLN_G = 4160, OFF_QLN_B = 4192, OFF_KLN_G = 4224,
              OFF_KLN_B = 4256;
constexpr float LN_EPS = 1e-6f;   // nn.LayerNorm(d, eps=1e-6): reference satrans.py:40, submodules.py:75

// counters[] slots of SatWork
constexpr int CNT_NTILES = 0, CNT_ERR = 1, CNT_NUNIQ = 2, CNT_NTILES2 = 3, CNT_HIST = 8;
constexpr int NT2 = 128;          // threads per CTA of the tensor-core forward (= TMEM lanes)
constexpr int MAX_KEYS_HIST = 48;  // scenario ids supported by the single-block bucketing kernel

// dropout sites inside one Meta_Transformer_Layer (reference submodules.py:99, satrans.py:87,92-94)
constexpr uint32_t SITE_Q = 0, SITE_K = 1, SITE_ATT = 2, SITE_OUT = 3;

// ------------------------------------------------------------------------------------------------
// ids: integers stored in the (B, ld) input matrix as f32 / i32 / i64 (reference meta_basemodel.py:
// 533-535 casts the fp32 column with .long(): truncation toward zero).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long load_id(const void* __restrict__ x, int kind, long long idx) {
    if (kind == SAT_IDS_F32) return __float2ll_rz(__ldg(reinterpret_cast<const float*>(x) + idx));
    if (kind == SAT_IDS_I32) return (long long)__ldg(reinterpret_cast<const int32_t*>(x) + idx);
    return __ldg(reinterpret_cast<const long long*>(x) + idx);
}
__device__ __forceinline__ float load_dense(const void* __restrict__ x, int kind, long long idx) {
    if (kind == SAT_IDS_F32) return __ldg(reinterpret_cast<const float*>(x) + idx);
    if (kind == SAT_IDS_I32) return (float)__ldg(reinterpret_cast<const int32_t*>(x) + idx);
    return (float)__ldg(reinterpret_cast<const long long*>(x) + idx);
}

// ------------------------------------------------------------------------------------------------
// counter-based RNG for dropout: two 16-bit uniforms per 32-bit hash.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
struct DropCtx {
    uint32_t key;     // per (seed, token, layer, site)
    uint32_t thr;     // drop if u16 < thr
    float scale;      // 1/(1-p)
    bool on;
};
// 64-bit step seed -> the 32 bits the per-token keys are derived from (once per kernel)
__device__ __forceinline__ uint32_t fold_seed(uint64_t seed) { return (uint32_t)seed ^ (uint32_t)(seed >> 32) * 0x9E3779B1U; }
// words of SatWork.step_state (device-side step state, see sat_step_begin)
constexpr int SS_SEED = 0, SS_STEP = 2, SS_HYPER = 4, SS_CONST = 8;
// the step's seed: by value, or from the device-side step state when the step is replayed from a CUDA graph
__device__ __forceinline__ uint32_t step_seed32(uint64_t seed, const unsigned long long* __restrict__ seed_ptr) {
    return fold_seed(seed_ptr ? *seed_ptr : seed);
}
__device__ __forceinline__ DropCtx make_drop(uint32_t seed32, uint32_t token_gid, uint32_t layer, uint32_t site,
                                             float p, bool training) {
    DropCtx c;
    c.on = training && p > 0.f;
    c.thr = (uint32_t)(p * 65536.f + 0.5f);
    c.scale = 1.f / (1.f - p);
    c.key = mix32(seed32 + token_gid * 0x9E3779B1U) ^ ((layer * 4u + site + 1u) * 0x85EBCA77U);
    return c;
}
// keep-mask factor (0 or 1/(1-p)) of element e
__device__ __forceinline__ float drop_factor(const DropCtx& c, uint32_t e) {
    if (!c.on) return 1.f;
    uint32_t r = mix32(c.key + (e >> 1) * 0xC2B2AE35U);
    uint32_t u = (e & 1u) ? (r >> 16) : (r & 0xFFFFu);
    return u < c.thr ? 0.f : c.scale;
}
// keep-mask factors of elements 2*pair and 2*pair + 1 from ONE hash (same values as drop_factor)
__device__ __forceinline__ void drop_pair(const DropCtx& c, uint32_t pair, float& f0, float& f1) {
    if (!c.on) { f0 = 1.f; f1 = 1.f; return; }
    const uint32_t r = mix32(c.key + pair * 0xC2B2AE35U);
    f0 = (r & 0xFFFFu) < c.thr ? 0.f : c.scale;
    f1 = (r >> 16) < c.thr ? 0.f : c.scale;
}
// v[i] *= keep-mask factor of element i, i < D
__device__ __forceinline__ void drop_apply32(const DropCtx& c, float (&v)[SAT_D]) {
#pragma unroll
    for (int i = 0; i < SAT_D; i += 2) {
        float f0, f1;
        drop_pair(c, i >> 1, f0, f1);
        v[i] *= f0; v[i + 1] *= f1;
    }
}

// Softmax exponential of (score - row max) <= 0: one multiply by log2(e) and one MUFU.EX2 instead of expf's range reduction
// (~7 instructions).  The difference is formed in the natural domain first, so the rounding of the log2(e) multiply is
// relative to |score - max| (where exp is tiny when that is large) and not to |score|: scaling the scores themselves into the
// log2 domain was measurably noisier in the K-side gradients.  y_pred error vs the reference's golden vectors stays 6e-8.
#ifndef SAT_FAST_EXP2
#define SAT_FAST_EXP2 1
#endif
__device__ __forceinline__ float softmax_exp(float x) {
#if SAT_FAST_EXP2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
#else
    return expf(x);
#endif
}

// Packed fp32 FMA (sm_100a FFMA2): c0 += a * g0, c1 += a * g1 in ONE instruction; each lane is an IEEE fp32 FMA, so the
// results are bit-identical to two fmaf().  ptxas folds the packing moves: the scalar is read as a broadcast operand and
// (g0, g1) / (c0, c1) as aligned register pairs.
#ifndef SAT_FFMA2
#define SAT_FFMA2 1
#endif
__device__ __forceinline__ void fma2(float& c0, float& c1, float a, float g0, float g1) {
#if SAT_FFMA2
    unsigned long long A, G, Cc, R;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(G) : "f"(g0), "f"(g1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(A), "l"(G), "l"(Cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(R));
#else
    c0 = fmaf(a, g0, c0); c1 = fmaf(a, g1, c1);
#endif
}
// (c0, c1) = (a0 + b0, a1 + b1) and (a0 * b0, a1 * b1): FADD2 / FMUL2, each lane an IEEE fp32 operation
__device__ __forceinline__ void add2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
#if SAT_FFMA2
    unsigned long long A, B, R;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(R));
#else
    c0 = a0 + b0; c1 = a1 + b1;
#endif
}
__device__ __forceinline__ void mul2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
#if SAT_FFMA2
    unsigned long long A, B, R;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(R) : "l"(A), "l"(B));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(R));
#else
    c0 = a0 * b0; c1 = a1 * b1;
#endif
}
// c0 += a0 * g0, c1 += a1 * g1
__device__ __forceinline__ void fma2v(float& c0, float& c1, float a0, float a1, float g0, float g1) {
#if SAT_FFMA2
    unsigned long long A, G, Cc, R;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(G) : "f"(g0), "f"(g1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(A), "l"(G), "l"(Cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(R));
#else
    c0 = fmaf(a0, g0, c0); c1 = fmaf(a1, g1, c1);
#endif
}

// two softmax exponentials: s = exp(s - mx), same arithmetic per lane as softmax_exp(s - mx)
__device__ __forceinline__ void softmax_exp_pair(float& s0, float& s1, float mx) {
#if SAT_FAST_EXP2
    float t0, t1;
    add2(t0, t1, s0, s1, -mx, -mx);
    mul2(t0, t1, t0, t1, 1.4426950408889634f, 1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(t1));
#else
    s0 = expf(s0 - mx); s1 = expf(s1 - mx);
#endif
}

// ------------------------------------------------------------------------------------------------
// thread-local small GEMV against a matrix in shared memory (all lanes read the same address:
// LDS.128 broadcasts).  a[] and y[] are registers.
//   vm_outer : y[j] += sum_i a[i] * M[i*LD + j]        (M is [NIN][LD], "x @ W")
//   vm_dot   : y[j] += sum_i a[i] * M[j*LD + i]        (M is [NOUT][LD], "x @ W^T")
// ------------------------------------------------------------------------------------------------
// rotate a register array left by R positions (compile-time indices only: pure register moves)
template <int N, int R>
__device__ __forceinline__ void rotate_regs(float (&v)[N]) {
    float t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = v[i];
#pragma unroll
    for (int i = 0; i < N - R; ++i) v[i] = v[i + R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[N - R + i] = t[i];
}

// The GEMVs are ROLLED loops over chunks of CH inputs (vm_outer) / CH outputs (vm_dot); the register array that
// would need a run-time index is rotated by CH after every chunk instead (N/CH rotations restore it).  A fully
// unrolled layer body is ~30 k SASS instructions and was instruction-fetch bound (ncu: stall_no_instruction #1).
template <int NIN, int NOUT, int LD>
__device__ __forceinline__ void vm_outer(const float* __restrict__ M, float (&a)[NIN], float (&y)[NOUT]) {
    constexpr int CH = (NIN * NOUT >= 512) ? ((NOUT >= 32) ? 8 : 16) : NIN;
    static_assert(NIN % CH == 0, "chunking");
#pragma unroll 1
    for (int c = 0; c < NIN / CH; ++c) {
        const float* Mc = M + c * CH * LD;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
#pragma unroll
            for (int j = 0; j < NOUT; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(Mc + i * LD + j);
                fma2(y[j + 0], y[j + 1], a[i], w.x, w.y);
                fma2(y[j + 2], y[j + 3], a[i], w.z, w.w);
            }
        }
        if (CH < NIN) rotate_regs<NIN, CH>(a);
    }
}
template <int NIN, int NOUT, int LD>
__device__ __forceinline__ void vm_dot(const float* __restrict__ M, const float (&a)[NIN], float (&y)[NOUT]) {
    constexpr int CH = (NIN * NOUT >= 512) ? ((NIN >= 32) ? 8 : 16) : NOUT;
    static_assert(NOUT % CH == 0, "chunking");
#pragma unroll 1
    for (int c = 0; c < NOUT / CH; ++c) {
        const float* Mc = M + c * CH * LD;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int i = 0; i < NIN; i += 4) {
                const float4 w = *reinterpret_cast<const float4*>(Mc + j * LD + i);
                fma2v(acc0, acc1, a[i + 0], a[i + 1], w.x, w.y);
                fma2v(acc0, acc1, a[i + 2], a[i + 3], w.z, w.w);
            }
            y[j] += acc0 + acc1;
        }
        if (CH < NOUT) rotate_regs<NOUT, CH>(y);
    }
}

template <int N>
__device__ __forceinline__ void zero(float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.f;
}

// LayerNorm over the 32 channels a thread owns.  g/b in shared memory.  Returns mean / rstd for the backward.
// sum / sum of squared deviations over the 32 channels with four interleaved partial sums (a single chain of 32 dependent
// adds is 128 cycles of latency per reduction for a warp that has one or two neighbours to hide it)
__device__ __forceinline__ float sum32(const float (&v)[D]) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < D; i += 4) { add2(s0, s1, s0, s1, v[i], v[i + 1]); add2(s2, s3, s2, s3, v[i + 2], v[i + 3]); }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float sumsq_dev32(const float (&v)[D], float mean) {
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < D; i += 4) {
        float t0, t1, t2, t3;
        add2(t0, t1, v[i], v[i + 1], -mean, -mean); add2(t2, t3, v[i + 2], v[i + 3], -mean, -mean);
        fma2v(q0, q1, t0, t1, t0, t1); fma2v(q2, q3, t2, t3, t2, t3);
    }
    return (q0 + q1) + (q2 + q3);
}

__device__ __forceinline__ void layernorm32(float (&v)[D], const float* __restrict__ g, const float* __restrict__ b,
                                            float& mean, float& rstd) {
    mean = sum32(v) * (1.f / D);
    const float q = sumsq_dev32(v, mean);
    rstd = 1.f / sqrtf(q * (1.f / D) + LN_EPS);
#pragma unroll
    for (int i = 0; i < D; i += 2) {                     // v = fmaf((v - mean) * rstd, g, b), two channels per instruction
        float t0, t1, o0 = b[i], o1 = b[i + 1];
        add2(t0, t1, v[i], v[i + 1], -mean, -mean);
        mul2(t0, t1, t0, t1, rstd, rstd);
        fma2v(o0, o1, t0, t1, g[i], g[i + 1]);
        v[i] = o0; v[i + 1] = o1;
    }
}

// 128-byte row load/store helpers (global, rows are 128 B aligned).  A thread owns a whole row, so the lanes of a warp touch 32
// different lines: with 16-byte accesses every 32-byte sector is visited by two instructions (ncu: "only 16 of the 32 bytes
// transmitted per sector are utilized", half of the L1/L2 sector traffic of the token kernels is excess).  sm_100 has 256-bit
// global loads / stores (LDG.E.256 / STG.E.256): one full sector per lane and instruction, half the LSU instructions.
#ifndef SAT_LDST256
#define SAT_LDST256 1
#endif
__device__ __forceinline__ void ldg256_nc(const float* p, float* v) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_cg(const float* p, float* v) {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&v)[D]) {
#if SAT_LDST256
#pragma unroll
    for (int i = 0; i < D; i += 8) ldg256_nc(p + i, &v[i]);
#else
#pragma unroll
    for (int i = 0; i < D; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(p + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
#endif
}
// the same through L2 only (ld.global.cg): for a row this kernel has written earlier.  load_row's non-coherent path
// (LDG.CONSTANT) may return the stale L1 copy of such a row.
__device__ __forceinline__ void load_row_cg(const float* p, float (&v)[D]) {
#if SAT_LDST256
#pragma unroll
    for (int i = 0; i < D; i += 8) ldg256_cg(p + i, &v[i]);
#else
#pragma unroll
    for (int i = 0; i < D; i += 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p + i));
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
#endif
}
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&v)[D]) {
#if SAT_LDST256
#pragma unroll
    for (int i = 0; i < D; i += 8) stg256(p + i, &v[i]);
#else
#pragma unroll
    for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
#endif
}

// 128-byte row <-> shared memory ([token][36] tiles: conflict-free 16 B accesses)
__device__ __forceinline__ void sts_row(float* __restrict__ dst, const float (&v)[D]) {
#pragma unroll
    for (int i = 0; i < D; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
__device__ __forceinline__ void lds_row(const float* __restrict__ src, float (&v)[D]) {
#pragma unroll
    for (int i = 0; i < D; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + i);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
}

// cooperative global -> shared copy of n floats (n % 4 == 0, both 16 B aligned)
__device__ __forceinline__ void copy_to_smem(float* __restrict__ dst, const float* __restrict__ src, int n, int tid,
                                             int nthreads) {
    for (int i = tid * 4; i < n; i += nthreads * 4)
        *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
}

// address of global table row `row` (row-sharded: shard = row % R, local row = row / R)
__device__ __forceinline__ const float* table_row(const SatModel& m, long long row) {
    if (m.n_shards == 1) return m.tables + row * D;
    const int owner = (int)(row % m.n_shards);
    return m.shard_tables[owner] + (row / m.n_shards) * D;
}

// key of the generated-weight table row used by (domain g, layer l, q/k side)
__device__ __forceinline__ int theta_key(int g, int l, int side, int L, bool pos) {
    return pos ? ((g * L + l) * 2 + side) : g;
}

}  // namespace sat

// host-side error plumbing (defined in capi.cu)
namespace sat_host {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
}  // namespace sat_host
