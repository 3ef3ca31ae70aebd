"""ctypes binding of libsatrans_b200.so (the C ABI declared in include/satrans_b200.h).

There is no CPU fallback: `lib()` raises if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C satrans_b200/csrc`).
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SAT_LIB') or os.path.join(HERE, 'lib', 'libsatrans_b200.so')   # SAT_LIB: A/B builds of the same ABI
CSRC = os.path.join(HERE, 'csrc')

SAT_D, SAT_HID, SAT_P, SAT_MAXF, SAT_LAYER_FLOATS = 32, 64, 4096, 32, 4288
SAT_REG_PARTS = 148 * 256 + 148 * 4
SAT_WORK_BUFFERS = 32
SAT_ACT_PER_LAYER = 6      # activation rows saved per layer and token by a training forward
SAT_GRID_MAX, SAT_SORT_RADIX_BITS, SAT_SORT_TILE, SAT_ROWS_CHUNK = 148 * 2, 9, 2048, 32
SAT_ABI_VERSION = 4
SAT_MAX_SHARDS = 8
IDS_F32, IDS_I32, IDS_I64 = 0, 1, 2
FLAG_POS, FLAG_META_Q, FLAG_META_K, FLAG_RELU, FLAG_RES = 1, 2, 4, 8, 16
FLAG_LOSS_MSE, FLAG_LOSS_MAE = 32, 64
FLAG_GATE_Q, FLAG_GATE_K, FLAG_BILINEAR, FLAG_ONLYEMB = 128, 256, 512, 1024
OPT_ADAM, OPT_SGD, OPT_ADAGRAD, OPT_RMSPROP = 0, 1, 2, 3
ERR_BAD_ID, ERR_BAD_DOMAIN = 1, 2
ADAM_TOUCHED, ADAM_UNTOUCHED, ADAM_DEFERRED = 1, 2, 4
SAT_DEFER_MAX = 16
SAT_DEFER_STATE_WORDS = 8 + 16 * SAT_DEFER_MAX
SAT_DEFER_REG_DOUBLES = 32 + SAT_DEFER_MAX * 148 * 4 + SAT_DEFER_MAX * 148 * 4 * SAT_DEFER_MAX + 148 * 64 * SAT_DEFER_MAX

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class SatModel(C.Structure):
    _fields_ = [('F', i32), ('L', i32), ('H', i32), ('n_dom', i32), ('n_dense', i32), ('n_cols', i32),
                ('flags', i32), ('enc_in', i32), ('dom_col', i32), ('pad0', i32),
                ('n_dense_params', i64),
                ('off_dom', i64), ('off_lid', i64), ('off_qkv', i64), ('off_encw', i64), ('off_encb', i64),
                ('off_layers', i64), ('off_headw', i64), ('off_headb', i64),
                ('total_rows', i64),
                ('row_off', vp), ('sparse_cols', vp), ('dense_cols', vp), ('tables', vp), ('dense', vp),
                ('n_shards', i32), ('shard_rank', i32), ('local_rows', i64),
                ('shard_tables', vp * 8), ('shard_bitmap', vp * 8), ('shard_word_rank', vp * 8),
                ('shard_row_grad', vp * 8), ('shard_recv', vp * 8), ('shard_recv_idx', vp * 8), ('shard_flags', vp * 8)]


WORK_BUFFERS = ['perm', 'dom', 'tiles', 'counters', 'theta', 'dtheta', 'enc_du', 'acts', 'dz', 'dx', 'dqkv',
                'loss_part', 'dgrad', 'bitmap', 'word_rank', 'block_sums', 'uniq_rows', 'row_grad', 'reg_part',
                'scalars', 'wpack', 'tiles2', 'gpart', 'sort_keys', 'sort_vals', 'sort_hist', 'chunk_info', 'pieces',
                'recv', 'recv_idx', 'peer_flags', 'step_state']
assert len(WORK_BUFFERS) == SAT_WORK_BUFFERS


class SatWork(C.Structure):
    _fields_ = [('B', i32), ('cap_B', i32), ('device_step', i32), ('defer_finish', i32), ('pad2', i32), ('head_bwd_done', i32)] + [(n, vp) for n in WORK_BUFFERS]


class SatAdam(C.Structure):
    _fields_ = [('lr', f32), ('beta1', f32), ('beta2', f32), ('eps', f32), ('l2_emb', f32),
                ('step', i32), ('dense_mode', i32), ('kind', i32),
                ('tab_m', vp), ('tab_v', vp), ('dense_m', vp), ('dense_v', vp),
                ('row_upto', vp), ('defer_state', vp), ('defer_reg', vp)]


# name -> (restype, argtypes); must list every symbol include/satrans_b200.h declares
PROTOTYPES = {
    'sat_abi_version': (C.c_int, []),
    'sat_last_error': (C.c_char_p, []),
    'sat_struct_sizes': (C.c_int, [C.POINTER(i64)]),
    'sat_workspace_sizes': (C.c_int, [C.POINTER(SatModel), i32, i32, C.POINTER(i64), C.POINTER(i32)]),
    'sat_gather_fwd': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, i32, vp, vp, vp]),
    'sat_bucket_by_domain': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), i32, vp]),
    'sat_encoder_fwd': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_encoder_bwd': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_forward': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, vp, C.POINTER(SatWork), vp, i32, f32, C.c_uint64, vp]),
    'sat_backward': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), f32, C.c_uint64, vp]),
    'sat_backward_tc': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), f32, C.c_uint64, i32, vp]),
    'sat_backward_finish': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), i32, vp]),
    'sat_rows_clear': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_rows_mark': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), vp]),
    'sat_rows_rank': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_rows_mark_and_rank': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), vp]),
    'sat_enable_peer_access': (C.c_int, [i32]),
    'sat_rows_sort': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, C.POINTER(SatWork), vp]),
    'sat_rows_reduce': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_rows_combine': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_peer_barrier': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_ipc_open': (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    'sat_ipc_close': (C.c_int, [vp]),
    'sat_step_begin': (C.c_int, [C.POINTER(SatWork), C.c_uint64, vp]),
    'sat_step_set': (C.c_int, [C.POINTER(SatWork), i32, f32, f32, f32, f32, vp]),
    'sat_tables_adam': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), C.POINTER(SatAdam), i32, vp]),
    'sat_defer_begin': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), C.POINTER(SatAdam), vp]),
    'sat_defer_catchup': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), C.POINTER(SatAdam), vp]),
    'sat_defer_sweep': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatAdam), vp]),
    'sat_dense_adam': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), C.POINTER(SatAdam), vp]),
    'sat_tables_sumsq': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), vp]),
    'sat_finalize_loss': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), f32, i32, vp]),
    'sat_selfatt_fwd': (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, C.c_uint64, vp]),
    'sat_field_attention_fwd': (C.c_int, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, C.c_uint64, vp]),
    'sat_field_attention_bwd_scratch': (C.c_int64, [i32, i32]),
    'sat_field_attention_bwd': (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, C.c_uint64,
                                          vp, vp, vp, vp, vp, vp, vp, vp]),
    'sat_head_fwd': (C.c_int, [vp, i32, vp, i32, i32, vp, vp, i32, vp, i32, vp]),
    'sat_pack_weights': (C.c_int, [C.POINTER(SatModel), C.POINTER(SatWork), i32, vp]),
    'sat_batch_metrics': (C.c_int, [vp, vp, vp, i32, i32, vp, vp]),
    'sat_forward_tc': (C.c_int, [C.POINTER(SatModel), vp, i32, i32, vp, C.POINTER(SatWork), vp, i32, f32, C.c_uint64, i32, vp]),
}

_lib = None


class SatError(RuntimeError):
    pass


def build(verbose=False, jobs=None):
    """Compile satrans_b200/csrc/*.cu for sm_100a into satrans_b200/lib/libsatrans_b200.so (make)."""
    jobs = jobs or os.cpu_count() or 4
    r = subprocess.run(['make', '-C', CSRC, '-j%d' % jobs], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise SatError('building libsatrans_b200.so failed')
    return LIB_PATH


def lib():
    """The loaded shared library with prototypes attached; raises SatError if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SatError('%s not found: build it first (make -C satrans_b200/csrc); there is no CPU fallback'
                           % LIB_PATH)
        import torch  # noqa: F401  (makes sure libcudart is already resolvable / loaded in-process)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.sat_abi_version() != SAT_ABI_VERSION:
            raise SatError('libsatrans_b200.so ABI version mismatch')
        sz = (i64 * 3)()
        l.sat_struct_sizes(sz)
        if list(sz) != [C.sizeof(SatModel), C.sizeof(SatWork), C.sizeof(SatAdam)]:
            raise SatError('struct layout mismatch between _lib.py and libsatrans_b200.so: %s' % list(sz))
        _lib = l
    return _lib


_DEBUG_SYNC = bool(os.environ.get('SAT_DEBUG_SYNC'))


def check(rc, what=''):
    if rc != 0:
        raise SatError('%s failed (%d): %s' % (what, rc, lib().sat_last_error().decode()))
    if _DEBUG_SYNC:          # debugging aid: surface asynchronous kernel faults at the call that caused them
        import torch
        try:
            torch.cuda.synchronize()
        except Exception as e:
            raise SatError('%s: asynchronous CUDA failure: %s' % (what, e))
