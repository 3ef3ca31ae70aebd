"""Host-side driver of the fused SATrans kernels: owns the C-ABI descriptors, per-batch workspaces and
stream choreography.  PyTorch is used only for device memory, streams and events.

One training step (reference meta_basemodel.py:311-328: forward, BCE(sum), dense L2, backward,
dense Adam) becomes, on two streams:

    main : bucket-by-scenario -> encoder -> pack weights -> fused forward(+BCE, dLogit) -> head bwd
           -> L x (attention bwd, MetaNet/QKV bwd) -> fixed-order reduce of the per-CTA partial gradients -> encoder bwd
           --(join)--> sorted segmented reduce of dE0 per touched row -> L2 + Adam sweep over the untouched rows -> finalize
    side : mark + rank the touched rows -> radix sort of the tokens by touched-row slot (ids only: runs beside the
           forward) ... later: Adam on the touched rows + Adam on the flat dense buffer (beside the HBM-bound sweep)

No floating-point atomics anywhere: every sum has a fixed order, so two runs of a step are bit-identical.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import SatAdam, SatModel, SatWork, check, lib

_BUF_DTYPES = {'perm': torch.int32, 'dom': torch.int32, 'tiles': torch.int32, 'counters': torch.int32,
               'loss_part': torch.float64, 'reg_part': torch.float64, 'bitmap': torch.int32,
               'word_rank': torch.int32, 'block_sums': torch.int32, 'uniq_rows': torch.int32,
               'sort_keys': torch.int32, 'sort_vals': torch.int32, 'sort_hist': torch.int32, 'chunk_info': torch.int32,
               'recv_idx': torch.int32, 'peer_flags': torch.int32}


def _id_kind(X):
    if X.dtype == torch.float32:
        return _lib.IDS_F32
    if X.dtype == torch.int32:
        return _lib.IDS_I32
    if X.dtype == torch.int64:
        return _lib.IDS_I64
    raise TypeError('input matrix must be float32, int32 or int64, got %s' % X.dtype)


class Workspace:
    """Caller-owned device buffers of one SatWork (sized by sat_workspace_sizes)."""

    def __init__(self, engine, cap_B, training):
        self.cap_B, self.training = int(cap_B), bool(training)
        sizes = (C.c_int64 * _lib.SAT_WORK_BUFFERS)()
        mt = C.c_int32()
        check(lib().sat_workspace_sizes(C.byref(engine.desc), self.cap_B, int(training), sizes, C.byref(mt)),
              'sat_workspace_sizes')
        self.max_tiles = mt.value
        self.dev_step = None       # value of the device-side step counter (step_state), None = unknown
        self.bufs = {}
        self.c = SatWork()
        self.c.B, self.c.cap_B = 0, self.cap_B
        dev = engine.device
        for name, nbytes in zip(_lib.WORK_BUFFERS, sizes):
            if nbytes == 0:
                setattr(self.c, name, None)
                continue
            dt = _BUF_DTYPES.get(name, torch.float32)
            n = (nbytes + dt.itemsize - 1) // dt.itemsize
            t = torch.zeros(n, dtype=dt, device=dev)
            if name == 'recv_idx':
                t.fill_(-1)            # "no row pushed" (sat_rows_combine re-arms the entries it consumes)
            self.bufs[name] = t
            setattr(self.c, name, t.data_ptr())

    def __getitem__(self, k):
        return self.bufs[k]


class Engine:
    def __init__(self, model):
        self.model = model
        self.device = None
        self.desc = SatModel()
        self._bound_ptrs = None
        self._work = {}
        self._adam = None
        self._sweep_stream = None
        self.side_streams = True          # run the independent short launch chains of a step on a second stream
        self._ev_marked = None
        self._ev_swept = None
        self.step_count = 0
        self.launches = 0          # kernels launched by this engine (counted per C-ABI call)
        self.timing = None         # set to {} to collect CUDA-event pairs per phase: {name: [(start, end), ...]}
        self.dist = None           # satrans_b200.dist.DistContext when the tables are row-sharded over ranks
        self.use_cuda_graph = True # train_on_batch: replay the step from a CUDA graph once a batch shape has been seen twice
        self._graphs = {}          # key -> captured step (see graph_train_step)
        self._graph_seen = {}
        # projection arithmetic: 'tf32x3' (default) = tcgen05 tensor cores with the error-compensated 3xTF32 split (fp32-level
        # parity: y_pred within 1e-7 of the CUDA-core path); 'fp32' = CUDA-core FFMA kernels (bit-tight parity tests use it);
        # 'tf32' = tcgen05 plain TF32 (fast mode)
        self.precision = 'tf32x3'
        if model._layout.get('flags', 0) & (_lib.FLAG_GATE_Q | _lib.FLAG_GATE_K | _lib.FLAG_BILINEAR):
            self.precision = 'fp32'       # the gate / bilinear ablation variants exist in the CUDA-core token kernels only
        # backward of the MetaNet/QKV block: None = follow `precision`; 'fp32' forces the CUDA-core kernel
        self.backward_precision = None
        self._peer_tables = None
        # deferred table sweep (sat_defer_*): steps begun since the last sweep, the SatAdam of the last deferred step
        self.defer_pending = 0
        self._defer_adam = None

    def attach_dist(self, ctx):
        """Row-sharded multi-GPU mode (collective: every rank calls it).  The model must have been built with
        table_shard=(ctx.rank, ctx.world)."""
        if tuple(self.model._shard) != (ctx.rank, ctx.world):
            raise ValueError('model.table_shard %s does not match the process group (%d, %d)'
                             % (self.model._shard, ctx.rank, ctx.world))
        self.dist = ctx
        if ctx.world > 1:           # replicated dense parameters start identical on every rank
            import torch.distributed as tdist
            self.model._ensure_flat()
            tdist.broadcast(self.model._dense, src=0, group=ctx.group)
        self._bound_ptrs = None
        self.bind()

    def prepare_training(self, cap_B):
        """Allocate the training workspace for batches of up to cap_B samples per rank and (sharded mode) map
        the peers' touched-row buffers.  Collective in sharded mode."""
        self.bind()
        w = self._work.get(True)
        if w is None or w.cap_B < cap_B:
            w = Workspace(self, int(cap_B), True)
            self._work[True] = w
            self._wire_shards(w)
        return w

    def _wire_shards(self, w):
        d = self.desc
        r = d.shard_rank
        if self.dist is None or self.dist.world == 1:
            d.shard_bitmap[r] = w['bitmap'].data_ptr()
            d.shard_word_rank[r] = w['word_rank'].data_ptr()
            d.shard_row_grad[r] = w['row_grad'].data_ptr()
            return
        for name, field in (('bitmap', d.shard_bitmap), ('word_rank', d.shard_word_rank), ('row_grad', d.shard_row_grad),
                            ('recv', d.shard_recv), ('recv_idx', d.shard_recv_idx), ('peer_flags', d.shard_flags)):
            for k, ptr in enumerate(self.dist.exchange_pointers(w[name])):
                field[k] = ptr
        torch.cuda.synchronize()
        self.dist.host_barrier()        # every peer's buffers are initialised before anybody's kernels touch them

    def _tic(self, name, stream=None):
        if self.timing is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream())
        return (name, ev, stream)

    def _toc(self, tok):
        if tok is None:
            return
        name, ev0, stream = tok
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record(stream if stream is not None else torch.cuda.current_stream())
        self.timing.setdefault(name, []).append((ev0, ev1))

    def timing_ms(self):
        """{phase: mean milliseconds} of the collected event pairs (synchronises)."""
        torch.cuda.synchronize()
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in (self.timing or {}).items()}

    # ------------------------------------------------------------------------------------------
    def bind(self):
        """(Re)build the SatModel descriptor if the model's buffers moved."""
        m = self.model
        m._ensure_flat()
        ptrs = (m._tables.data_ptr(), m._dense.data_ptr())
        if self._bound_ptrs == ptrs:
            return
        if self.defer_pending:
            raise _lib.SatError('the parameter buffers moved while deferred table updates were pending: call '
                                'model.sync_tables() before moving / replacing the parameters')
        if not m._tables.is_cuda:
            raise _lib.SatError('the SATrans B200 path needs the model on a CUDA device (got %s); there is no '
                                'CPU fallback' % m._tables.device)
        lib()
        self.device = m._tables.device
        lay = m._layout
        d = self.desc
        d.F, d.L, d.H = lay['F'], lay['L'], lay['H']
        d.n_dom, d.n_dense, d.n_cols = lay['n_dom'], lay['n_dense'], lay['n_cols']
        loss_kind = getattr(m, '_loss_kind', 'binary_crossentropy')
        d.flags = lay['flags'] | {'binary_crossentropy': 0, 'mse': _lib.FLAG_LOSS_MSE, 'mae': _lib.FLAG_LOSS_MAE}[loss_kind]
        d.enc_in, d.dom_col = lay['enc_in'], lay['dom_col']
        d.n_dense_params = lay['n_dense_params']
        for k in ('off_dom', 'off_lid', 'off_qkv', 'off_encw', 'off_encb', 'off_layers', 'off_headw', 'off_headb'):
            setattr(d, k, lay[k])
        d.total_rows = lay['total_rows']
        self._row_off = torch.tensor(lay['row_off'], dtype=torch.int64, device=self.device)
        self._sparse_cols = torch.tensor(lay['sparse_cols'], dtype=torch.int32, device=self.device)
        self._dense_cols = torch.tensor(lay['dense_cols'] or [0], dtype=torch.int32, device=self.device)
        d.row_off, d.sparse_cols = self._row_off.data_ptr(), self._sparse_cols.data_ptr()
        d.dense_cols = self._dense_cols.data_ptr()
        d.tables, d.dense = ptrs
        rank, world = m._shard
        d.n_shards, d.shard_rank = world, rank
        d.local_rows = m._tables.shape[0]
        for k in range(_lib.SAT_MAX_SHARDS):
            d.shard_tables[k] = None
            d.shard_bitmap[k] = None
            d.shard_word_rank[k] = None
            d.shard_row_grad[k] = None
            d.shard_recv[k] = None
            d.shard_recv_idx[k] = None
            d.shard_flags[k] = None
        if world == 1:
            d.shard_tables[0] = ptrs[0]
        else:
            if self.dist is None:
                raise _lib.SatError('model built with table_shard=%s: call engine.attach_dist(DistContext()) first'
                                    % (m._shard,))
            for k, ptr in enumerate(self.dist.exchange_pointers(m._tables)):
                d.shard_tables[k] = ptr
        self._bound_ptrs = ptrs
        self._work = {}
        self._graphs, self._graph_seen = {}, {}      # captured steps hold the old pointers
        if self._adam is not None and self._adam['tab_v'] is not None and \
                (self._adam['tab_v'].device != self.device or self._adam['tab_v'].shape != m._tables.shape):
            # the parameters moved (model.cpu() / .cuda() around a checkpoint, as reference main.py does): the optimiser
            # state follows them, like torch.optim's state tensors follow param.data only if re-created -- here it is kept
            for k in ('tab_m', 'tab_v', 'dense_m', 'dense_v', 'row_upto', 'defer_state', 'defer_reg'):
                if self._adam[k] is not None:
                    self._adam[k] = self._adam[k].to(self.device)

    def workspace(self, B, training):
        key = bool(training)
        w = self._work.get(key)
        if w is None or w.cap_B < B:
            if training and self.dist is not None and self.dist.world > 1:
                raise _lib.SatError('sharded mode: call engine.prepare_training(cap_B) (collective) before stepping '
                                    'with a batch of %d samples' % B)
            cap = max(int(B), w.cap_B if w else 0)
            w = Workspace(self, cap, training)
            self._work[key] = w
            if training:
                self._wire_shards(w)
        w.c.B = int(B)
        return w

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _check_x(self, X):
        if not (X.is_cuda and X.dim() == 2 and X.is_contiguous() and X.shape[1] == self.desc.n_cols):
            raise ValueError('expected a contiguous CUDA (B, %d) input matrix, got %s on %s'
                             % (self.desc.n_cols, tuple(X.shape), X.device))
        return _id_kind(X)

    # ------------------------------------------------------------------------------------------
    def gather(self, X):
        """(B, F, 32) field embeddings of the batch — input_from_feature_columns + cat."""
        self.sync_tables()
        self.bind()
        kind = self._check_x(X)
        B = X.shape[0]
        out = torch.empty((B, self.desc.F, _lib.SAT_D), dtype=torch.float32, device=self.device)
        w = self.workspace(max(B, 1), False)
        check(lib().sat_gather_fwd(C.byref(self.desc), X.data_ptr(), kind, X.shape[1], B, out.data_ptr(),
                                   w['counters'].data_ptr() + 4, self._stream()), 'sat_gather_fwd')
        self.launches += 1
        return out

    def _prepare(self, X, w, kind):
        s = self._stream()
        l = lib()
        # counters[1] (error flags) is sticky until check_errors() reads and clears it
        check(l.sat_bucket_by_domain(C.byref(self.desc), X.data_ptr(), kind, X.shape[1], C.byref(w.c), 0, s),
              'sat_bucket_by_domain')
        check(l.sat_encoder_fwd(C.byref(self.desc), C.byref(w.c), s), 'sat_encoder_fwd')
        if self.precision != 'fp32':
            check(l.sat_pack_weights(C.byref(self.desc), C.byref(w.c), self._passes(), s), 'sat_pack_weights')

    def _passes(self):
        try:
            return {'fp32': 0, 'tf32x3': 3, 'tf32': 1, 'bf16': 2}[self.precision]
        except KeyError:
            raise ValueError("precision must be 'fp32', 'tf32x3', 'tf32' or 'bf16', got %r" % (self.precision,))

    def _forward_call(self, X, kind, y_ptr, w, y_pred, training, p, seed, s):
        l, d = lib(), C.byref(self.desc)
        passes = self._passes()
        if passes == 0:
            check(l.sat_forward(d, X.data_ptr(), kind, X.shape[1], y_ptr, C.byref(w.c), y_pred.data_ptr(),
                                int(training), float(p), seed, s), 'sat_forward')
        else:
            check(l.sat_forward_tc(d, X.data_ptr(), kind, X.shape[1], y_ptr, C.byref(w.c), y_pred.data_ptr(),
                                   int(training), float(p), seed, passes, s), 'sat_forward_tc')

    def _backward_passes(self):
        return {'fp32': 0, 'tf32x3': 3, 'tf32': 1, 'bf16': 2}[self.backward_precision or self.precision]

    def _backward_call(self, X, kind, w, p, seed, s, defer_finish=False):
        """defer_finish: stop after the layer kernels; the caller runs _backward_finish (partial-row reduction + encoder
        backward) later, e.g. on the second stream beside the table sweep."""
        l, d = lib(), C.byref(self.desc)
        w.c.defer_finish = int(bool(defer_finish))
        bp = self.backward_precision or self.precision
        passes = {'fp32': 0, 'tf32x3': 3, 'tf32': 1, 'bf16': 2}[bp]
        if passes and passes != self._passes() and self.precision != 'fp32':
            raise ValueError("the tensor-core backward reads the weights packed for the forward: same precision required")
        if passes and self.precision == 'fp32':
            raise ValueError("a tensor-core backward needs the packed weights of a tensor-core forward precision")
        if passes == 0:
            check(l.sat_backward(d, X.data_ptr(), kind, X.shape[1], C.byref(w.c), float(p), seed, s), 'sat_backward')
        else:
            check(l.sat_backward_tc(d, X.data_ptr(), kind, X.shape[1], C.byref(w.c), float(p), seed, passes, s),
                  'sat_backward_tc')

    def _backward_finish(self, w, stream):
        check(lib().sat_backward_finish(C.byref(self.desc), C.byref(w.c), self._backward_passes(), stream), 'sat_backward_finish')

    def _barrier(self, w, stream=None):
        """Stream-ordered cross-process barrier through the peer-mapped flags (one tiny kernel, no NCCL call)."""
        check(lib().sat_peer_barrier(C.byref(self.desc), C.byref(w.c), self._stream() if stream is None else stream),
              'sat_peer_barrier')

    def _mark_rows(self, X, w, kind, catchup=None):
        """Touched-row bitmap + ranks.  Sharded: clear | mark (local staging regions) | barrier | merge + rank (| catch-up) | barrier.
        catchup: the SatAdam of a deferred-sweep step -- this shard's rows that ANY rank's batch reads are brought up to date
        before the last barrier, i.e. before a peer's forward can gather them."""
        l, s, d = lib(), self._stream(), C.byref(self.desc)
        if self.dist is None or self.dist.world == 1:
            check(l.sat_rows_mark_and_rank(d, X.data_ptr(), kind, X.shape[1], C.byref(w.c), s), 'sat_rows_mark_and_rank')
            if catchup is not None:
                check(l.sat_defer_catchup(d, C.byref(w.c), C.byref(catchup), s), 'sat_defer_catchup')
            return
        # (no barrier between clear and mark: the marks are local; the peers finished reading this process' staging regions
        #  before the previous step's last barrier)
        check(l.sat_rows_clear(d, C.byref(w.c), s), 'sat_rows_clear')
        check(l.sat_rows_mark(d, X.data_ptr(), kind, X.shape[1], C.byref(w.c), s), 'sat_rows_mark')
        self._barrier(w)
        check(l.sat_rows_rank(d, C.byref(w.c), s), 'sat_rows_rank')
        if catchup is not None:
            check(l.sat_defer_catchup(d, C.byref(w.c), C.byref(catchup), s), 'sat_defer_catchup')
        self._barrier(w)

    def _sort_rows(self, X, w, kind):
        """Tokens sorted by (owner shard, touched-row slot): needs the scenario permutation and the ranks."""
        check(lib().sat_rows_sort(C.byref(self.desc), X.data_ptr(), kind, X.shape[1], C.byref(w.c), self._stream()),
              'sat_rows_sort')

    def _reduce_rows(self, w):
        """dE0 per token (w['dx']) -> w['row_grad'][slot]: sorted segmented reduce; sharded: push + barrier + combine."""
        l, s, d = lib(), self._stream(), C.byref(self.desc)
        check(l.sat_rows_reduce(d, C.byref(w.c), s), 'sat_rows_reduce')
        if self.dist is not None and self.dist.world > 1:
            self._barrier(w)
            check(l.sat_rows_combine(d, C.byref(w.c), s), 'sat_rows_combine')

    def forward(self, X, y=None, training=False, dropout_p=0.1, seed=0):
        """y_pred (B,) in sample order.  With y: also leaves BCE partials / dLogit in the workspace."""
        self.sync_tables()
        self.bind()
        kind = self._check_x(X)
        B = X.shape[0]
        y_pred = torch.empty(B, dtype=torch.float32, device=self.device)
        if B == 0:
            return y_pred
        w = self.workspace(B, training)
        self._prepare(X, w, kind)
        self._forward_call(X, kind, y.data_ptr() if y is not None else None, w, y_pred, training, dropout_p,
                           int(seed) & (2 ** 64 - 1), self._stream())
        self.launches += 3 + (1 if self.precision != 'fp32' else 0)
        return y_pred

    def check_errors(self, w):
        """Raise like the reference's nn.Embedding would for ids outside their table (device->host sync)."""
        if self.dist is not None and self.dist.world > 1 and 'peer_flags' in w.bufs:
            stuck = int(w['peer_flags'][9].item())
            if stuck:
                raise _lib.SatError('cross-process barrier timed out waiting for rank %d: the ranks no longer run the same '
                                    'sequence of steps' % (stuck - 1))
        flags = int(w['counters'][1].item())
        if flags:
            w['counters'][1] = 0
        if flags & _lib.ERR_BAD_ID:
            raise IndexError('a sparse feature id is outside its embedding table (index out of range in self)')
        if flags & _lib.ERR_BAD_DOMAIN:
            raise IndexError('a scenario id is outside [0, num_domains] (index out of range in self)')

    # ------------------------------------------------------------------------------------------
    def loss_and_grads(self, X, y, training=True, dropout_p=0.1, seed=0, l2_emb=0.0, with_reg=True):
        """forward + BCE(sum) + backward.  Afterwards: w['dgrad'] = flat dense-parameter gradient,
        w['row_grad'][:n_unique] / w['uniq_rows'][:n_unique] = dBCE/d(table row) of the touched rows
        (the 2*l2*w regulariser term is added inside the optimiser), w['scalars'][0:2] = (bce, reg)."""
        self.sync_tables()
        self.bind()
        kind = self._check_x(X)
        B = X.shape[0]
        w = self.workspace(B, True)
        l, s, d = lib(), self._stream(), C.byref(self.desc)
        self._prepare(X, w, kind)
        self._mark_rows(X, w, kind)
        self._sort_rows(X, w, kind)
        y_pred = torch.empty(B, dtype=torch.float32, device=self.device)
        self._forward_call(X, kind, y.data_ptr(), w, y_pred, 1, float(dropout_p if training else 0.0),
                           int(seed) & (2 ** 64 - 1), s)
        self._backward_call(X, kind, w, float(dropout_p if training else 0.0), int(seed) & (2 ** 64 - 1), s)
        self._reduce_rows(w)
        self.launches += self._launch_count(False)
        if self.dist is not None and self.dist.world > 1:
            self.dist.all_reduce_sum(w['dgrad'])
        if with_reg:
            check(l.sat_tables_sumsq(d, C.byref(w.c), s), 'sat_tables_sumsq')
            self.launches += 1
        check(l.sat_finalize_loss(d, C.byref(w.c), float(l2_emb), int(self.precision != 'fp32'), s), 'sat_finalize_loss')
        if self.dist is not None and self.dist.world > 1:
            self.dist.all_reduce_sum(w['scalars'][0:2])
        self.launches += 1
        return y_pred, w

    # ------------------------------------------------------------------------------------------
    def _adam_state(self, kind=0):
        """Optimiser state buffers: Adam m+v, Adagrad / RMSprop v only, SGD none."""
        if self._adam is None or self._adam.get('kind') != kind:
            self.sync_tables()
            m = self.model
            self._graphs, self._graph_seen = {}, {}
            need_m, need_v = kind == _lib.OPT_ADAM, kind != _lib.OPT_SGD
            self._adam = {
                'kind': kind,
                'tab_m': torch.zeros_like(m._tables) if need_m else None,
                'tab_v': torch.zeros_like(m._tables) if need_v else None,
                'dense_m': torch.zeros_like(m._dense) if need_m else None,
                'dense_v': torch.zeros_like(m._dense) if need_v else None,
                'row_upto': None, 'defer_state': None, 'defer_reg': None,
            }
        return self._adam

    def _defer_buffers(self, st):
        if st['row_upto'] is None:
            st['row_upto'] = torch.zeros(self.model._tables.shape[0], dtype=torch.uint8, device=self.device)
            st['defer_state'] = torch.zeros(_lib.SAT_DEFER_STATE_WORDS, dtype=torch.float32, device=self.device)
            st['defer_reg'] = torch.zeros(_lib.SAT_DEFER_REG_DOUBLES, dtype=torch.float64, device=self.device)
        return st

    def sync_tables(self):
        """Close the current sweep epoch of the deferred table updates: every table row (and its optimiser state) is brought up to
        date, exactly as if the every-row sweep had run after each step.  A no-op when nothing is pending.  Every engine entry
        point that reads the tables calls it first; call it yourself before touching model._tables / the optimiser state
        directly.  Stream-ordered on the current stream (no host synchronisation)."""
        if self.defer_pending == 0:
            return
        t = self._tic('table_sweep')
        check(lib().sat_defer_sweep(C.byref(self.desc), C.byref(self._defer_adam), self._stream()), 'sat_defer_sweep')
        self._toc(t)
        if self.dist is not None and self.dist.world > 1:
            # row-sharded: every rank closes its shard's epoch at the same point of the (SPMD) call sequence; no peer may gather
            # a row before its owner's sweep has finished
            self._barrier(self._work[True])
        self.defer_pending = 0
        self.launches += 2

    def deferred_reg_total(self):
        """Sum of the regulariser terms of every step of every closed sweep epoch so far (Python float; synchronises)."""
        if self._adam is None or self._adam.get('defer_reg') is None:
            return 0.0
        t = self._adam['defer_reg'][_lib.SAT_DEFER_MAX].clone()
        if self.dist is not None and self.dist.world > 1:      # every shard holds the terms of its own rows (collective)
            self.dist.all_reduce_sum(t)
        return float(t.item())

    def deferred_reg(self):
        """(per-step regulariser terms of the last closed sweep epoch, running total over all closed epochs) as device tensors
        (float64); valid after sync_tables()."""
        r = self._adam['defer_reg']
        n = int(r[_lib.SAT_DEFER_MAX + 1].item())
        if int(r[_lib.SAT_DEFER_MAX + 2].item()):
            raise _lib.SatError('deferred table sweep: more than %d steps were begun without a sweep' % _lib.SAT_DEFER_MAX)
        return r[0:n], r[_lib.SAT_DEFER_MAX]

    def _launch_count(self, with_update, dense_mode=True):
        """Kernels of THIS library launched by one loss_and_grads / train_step (memsets and NCCL calls not counted)."""
        L, tc = self.desc.L, self.precision != 'fp32'
        sharded = self.dist is not None and self.dist.world > 1
        cap = self._work[True].cap_B * self.desc.F if True in self._work else 1
        idx = min(cap * self.desc.n_shards, -(-self.desc.total_rows // self.desc.n_shards))
        key_bits = max(idx, 1).bit_length() + (self.desc.n_shards - 1).bit_length()
        passes = -(-key_bits // _lib.SAT_SORT_RADIX_BITS)
        n = 2 + (1 if tc else 0)              # bucket, encoder, pack
        n += 3 + (3 if sharded else 0)        # mark, rank x2 (+ barriers)
        n += 1 + 3 * passes                   # sort keys + radix passes
        n += 1                                # forward
        n += 1 + 2 * L + 1 + 2                # head bwd, per layer attn + meta, dense reduce, encoder bwd x2
        n += 2 + (2 if sharded else 0)        # rows reduce + fixup (+ barrier + combine)
        if with_update:
            n += 1 + (1 if dense_mode else 0) + 1      # adam rows, sweep, dense adam
        return n

    def train_step(self, X, y, lr, betas=(0.9, 0.999), eps=1e-8, l2_emb=1e-5, dense_mode=True, training=True,
                   dropout_p=0.1, seed=0, overlap=False, opt_kind=0, _capturing=False, defer=1, loss_fn=None):
        """One optimiser step with the reference's semantics (dense_mode=True) or lazy/sparse table updates
        (dense_mode=False).  Returns (y_pred, workspace); scalars[0:2] = (bce sum, reg loss) after the step's
        finalize kernel (read them with .item() only when needed: that synchronises).

        defer = K > 1 (single GPU, dense_mode): the every-row table sweep runs once per K steps instead of after every step
        (include/satrans_b200.h, sat_defer_*): the rows this batch reads are brought up to date before the forward, all others
        when the epoch closes (after K steps, or at sync_tables()) -- parameters and optimiser state are bit-identical to
        defer=1.  scalars[1] of a deferred step is NOT its regulariser term (that needs every row): read deferred_reg() after
        sync_tables().

        loss_fn: a custom loss callable `loss_fn(y_pred, y, reduction='sum')` on torch tensors (reference
        meta_basemodel.py:627-640 accepts any callable).  The fused kernels then run the forward without labels, torch
        evaluates the loss and d loss / d y_pred on the (B,) prediction vector, and the backward starts from that
        (dLoss/dlogit = dLoss/dy_pred * y_pred * (1 - y_pred), written in scenario-sorted order); scalars[0] = the loss value."""
        self.bind()
        kind = self._check_x(X)
        B = X.shape[0]
        w = self.workspace(B, True)
        deferred = bool(int(defer) > 1 and dense_mode and not overlap)
        if not _capturing and self.defer_pending and (not deferred or self.defer_pending >= min(int(defer), _lib.SAT_DEFER_MAX)):
            self.sync_tables()
        st = self._adam_state(opt_kind)
        if not _capturing:                       # (a captured step takes its step index from the device-side state)
            self.step_count += 1
        a = SatAdam()
        a.kind = int(opt_kind)
        a.lr, a.beta1, a.beta2, a.eps, a.l2_emb = float(lr), float(betas[0]), float(betas[1]), float(eps), float(l2_emb)
        a.step, a.dense_mode = max(self.step_count, 1), int(bool(dense_mode))
        a.tab_m = st['tab_m'].data_ptr() if st['tab_m'] is not None else None
        a.tab_v = st['tab_v'].data_ptr() if st['tab_v'] is not None else None
        a.dense_m = st['dense_m'].data_ptr() if st['dense_m'] is not None else None
        a.dense_v = st['dense_v'].data_ptr() if st['dense_v'] is not None else None
        if deferred:
            self._defer_buffers(st)
            a.row_upto, a.defer_state, a.defer_reg = (st['row_upto'].data_ptr(), st['defer_state'].data_ptr(),
                                                      st['defer_reg'].data_ptr())
            self._defer_adam = a
        l, d = lib(), C.byref(self.desc)
        main = torch.cuda.current_stream()
        s = main.cuda_stream
        if self._sweep_stream is None:
            self._sweep_stream = torch.cuda.Stream(device=self.device)
            # the short update chain beside the sweep: high priority, so that its CTAs are placed as soon as sweep CTAs retire
            self._tail_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._ev_marked = torch.cuda.Event()
            self._ev_swept = torch.cuda.Event()
            self._ev_fork = torch.cuda.Event()
            self._ev_bucket = torch.cuda.Event()
            self._ev_caught = torch.cuda.Event()
        side = self._sweep_stream
        sd = int(seed) & (2 ** 64 - 1)
        p = float(dropout_p if training else 0.0)
        single = self.dist is None or self.dist.world == 1
        use_side = bool(self.side_streams)
        join_before_forward = deferred or not single     # the forward must not gather rows before the catch-up / the peers' barrier
        t = self._tic('prepare')
        if deferred:                                   # this step's optimiser constants enter the epoch's ring
            check(l.sat_defer_begin(d, C.byref(w.c), C.byref(a), s), 'sat_defer_begin')
        if use_side:
            # the touched-row bitmap / ranks and the token sort depend on the ids only: they run on the second stream beside
            # bucketing + encoder + weight packing + the forward, and are joined before the row-gradient reduce
            self._ev_fork.record(main)
            side.wait_event(self._ev_fork)
            with torch.cuda.stream(side):
                self._mark_rows(X, w, kind, a if deferred else None)   # (the rows this batch reads catch up BEFORE the forward)
                if join_before_forward:
                    self._ev_caught.record(side)
            self._prepare(X, w, kind)
            self._ev_bucket.record(main)
            side.wait_event(self._ev_bucket)           # the sort reads the scenario permutation
            with torch.cuda.stream(side):
                self._sort_rows(X, w, kind)
            self._ev_marked.record(side)
            if join_before_forward:
                main.wait_event(self._ev_caught)
        else:
            self._prepare(X, w, kind)
            self._mark_rows(X, w, kind, a if deferred else None)
            self._sort_rows(X, w, kind)
        self._toc(t)
        sweep_on_side = bool(dense_mode and overlap and single)
        if sweep_on_side:
            if not use_side:
                self._ev_marked.record(main)
            side.wait_event(self._ev_marked)
            t = self._tic('adam_sweep', side)
            check(l.sat_tables_adam(d, C.byref(w.c), C.byref(a), _lib.ADAM_UNTOUCHED, side.cuda_stream), 'sat_tables_adam')
            self._toc(t)
            self._ev_swept.record(side)
        y_pred = torch.empty(B, dtype=torch.float32, device=self.device)
        t = self._tic('forward')
        self._forward_call(X, kind, y.data_ptr() if loss_fn is None else None, w, y_pred, 1, p, sd, s)
        loss_val = None
        if loss_fn is not None:
            with torch.enable_grad():
                yp = y_pred.detach().clone().requires_grad_(True)
                loss_t = loss_fn(yp, y.reshape(-1).to(yp.dtype), reduction='sum')
                (g_yp,) = torch.autograd.grad(loss_t, yp)
            dz = g_yp * y_pred * (1.0 - y_pred)                     # through the sigmoid of the logit head
            w['dz'][:B] = dz.index_select(0, w['perm'][:B].long())  # sorted position p -> sample perm[p]
            loss_val = loss_t.detach()
        self._toc(t)
        tail_on_side = bool(single and dense_mode and not sweep_on_side and self.side_streams and not deferred)
        # deferred sweep, one GPU: the dense-parameter chain (fixed-order reduction of the per-CTA partials, encoder backward, dense
        # Adam) runs on the second stream beside the table-row chain (sorted segmented reduce, touched-row Adam): disjoint data
        split_tail = bool(single and deferred and self.side_streams)
        t = self._tic('backward')
        self._backward_call(X, kind, w, p, sd, s, defer_finish=tail_on_side or split_tail)
        self._toc(t)
        if tail_on_side:
            # Everything that only the (short) touched-row and dense-parameter updates need -- the fixed-order reduction of the
            # per-CTA partial gradients, the encoder backward, the sorted segmented reduce of dE0 -- runs on the second stream
            # beside the HBM-bound sweep over the untouched rows (disjoint data; the sweep only needs the touched-row bitmap,
            # which the second stream produced before: same-stream order).
            tail = self._tail_stream
            self._ev_fork.record(main)
            tail.wait_event(self._ev_fork)
            tail.wait_event(self._ev_marked)           # sorted order + touched-row ranks (second stream)
            main.wait_event(self._ev_marked)           # (the sweep reads the bitmap)
            with torch.cuda.stream(tail):
                t = self._tic('finish+rows+adam (side stream)', tail)
                self._backward_finish(w, tail.cuda_stream)
                self._reduce_rows(w)
                check(l.sat_tables_adam(d, C.byref(w.c), C.byref(a), _lib.ADAM_TOUCHED, tail.cuda_stream), 'sat_tables_adam')
                check(l.sat_dense_adam(d, C.byref(w.c), C.byref(a), tail.cuda_stream), 'sat_dense_adam')
                self._toc(t)
            self._ev_swept.record(tail)
            t = self._tic('adam_sweep')
            check(l.sat_tables_adam(d, C.byref(w.c), C.byref(a), _lib.ADAM_UNTOUCHED, s), 'sat_tables_adam')
            self._toc(t)
            main.wait_event(self._ev_swept)
        else:
            if split_tail:
                tail = self._tail_stream
                self._ev_fork.record(main)
                tail.wait_event(self._ev_fork)
                with torch.cuda.stream(tail):
                    t = self._tic('finish+adam_dense (side stream)', tail)
                    self._backward_finish(w, tail.cuda_stream)
                    check(l.sat_dense_adam(d, C.byref(w.c), C.byref(a), tail.cuda_stream), 'sat_dense_adam')
                    self._toc(t)
                self._ev_swept.record(tail)
            if use_side:
                main.wait_event(self._ev_marked)       # sorted order + touched-row ranks are ready
            t = self._tic('rows_reduce')
            self._reduce_rows(w)
            self._toc(t)
            if not single:
                t = self._tic('allreduce')
                self.dist.all_reduce_sum(w['dgrad'])
                self._toc(t)
            t = self._tic('adam_rows')
            # (lazy mode: ADAM_UNTOUCHED only clears the sweep's slice of the sum-of-squares partials)
            check(l.sat_tables_adam(d, C.byref(w.c), C.byref(a),
                                    _lib.ADAM_TOUCHED | (0 if dense_mode else _lib.ADAM_UNTOUCHED) |
                                    (_lib.ADAM_DEFERRED if deferred else 0), s), 'sat_tables_adam')
            self._toc(t)
            if dense_mode and not sweep_on_side and not deferred:
                t = self._tic('adam_sweep')
                check(l.sat_tables_adam(d, C.byref(w.c), C.byref(a), _lib.ADAM_UNTOUCHED, s), 'sat_tables_adam')
                self._toc(t)
            if split_tail:
                main.wait_event(self._ev_swept)
            else:
                t = self._tic('adam_dense')
                check(l.sat_dense_adam(d, C.byref(w.c), C.byref(a), s), 'sat_dense_adam')
                self._toc(t)
            if sweep_on_side:
                main.wait_event(self._ev_swept)
        check(l.sat_finalize_loss(d, C.byref(w.c), float(l2_emb), int(self.precision != 'fp32'), s), 'sat_finalize_loss')
        if loss_val is not None:
            w['scalars'][0] = loss_val.to(torch.float32)
        if not single:
            self.dist.all_reduce_sum(w['scalars'][0:2])     # global BCE sum and L2 term
        if not _capturing:
            self.launches += self._launch_count(True, dense_mode) + 1 + (1 if deferred else 0)    # (begin + catch-up - sweep)
            if deferred:
                self.defer_pending += 1
        return y_pred, w

    # ------------------------------------------------------------------------------------------
    def graph_train_step(self, Xh, yh, lr, betas, eps, opt_kind, l2_emb, dense_mode, dropout_p, seed0, defer=1):
        """One optimiser step on a host (or device-resident) batch, replayed from a CUDA graph (single-GPU mode).

        The whole step -- H2D copies of the batch and of the four hyper-parameters from pinned staging buffers,
        sat_step_begin (step counter, dropout seed = seed0 + step - 1 and Adam's bias corrections advance ON THE DEVICE) and
        every kernel of train_step on its two streams -- is captured once per (batch shape, dtype, l2, mode, dropout,
        optimiser kind, precision) and then costs one cudaGraphLaunch per step instead of ~45 launches and their host-side
        marshalling.  lr / betas / eps are data (a scheduler can change them every step without re-capturing).
        Returns (workspace, pinned host tensor): the last node of the graph copies w['scalars'][0:4] = (bce, reg, n_unique, error
        flags) into the pinned tensor; it is valid after a synchronisation of the current stream."""
        Xh = torch.as_tensor(Xh)
        yh = torch.as_tensor(yh).reshape(-1)
        deferred = bool(int(defer) > 1 and dense_mode)
        key = (tuple(Xh.shape), Xh.dtype, float(l2_emb), bool(dense_mode), float(dropout_p), int(opt_kind), self.precision, deferred)
        if self.defer_pending and (not deferred or self.defer_pending >= min(int(defer), _lib.SAT_DEFER_MAX)):
            self.sync_tables()                     # (outside the graph: the host decides when a sweep epoch closes)
        G = self._graphs.get(key)
        if G is None:
            G = self._capture_step(key, Xh, lr, betas, eps, opt_kind, l2_emb, dense_mode, dropout_p, seed0, defer)
        # H2D copies of the batch: stream-ordered in front of the replay, straight from the caller's buffers when they are pinned
        # (no host-side staging pass), through the pinned staging buffers otherwise
        if Xh.is_cuda or Xh.is_pinned():
            G['Xs'].copy_(Xh, non_blocking=True)
        else:
            G['x_pin'].copy_(Xh)
            G['Xs'].copy_(G['x_pin'], non_blocking=True)
        if yh.is_cuda or (yh.dtype == torch.float32 and yh.is_pinned()):
            G['ys'].copy_(yh, non_blocking=True)
        else:
            G['y_pin'].copy_(yh)
            G['ys'].copy_(G['y_pin'], non_blocking=True)
        hp = G['hyper_pin']
        hp[0], hp[1], hp[2], hp[3] = float(lr), float(betas[0]), float(betas[1]), float(eps)
        w = G['w']
        if w.dev_step != self.step_count:          # eager steps ran in between: the device-side counter continues from the host's
            check(lib().sat_step_set(C.byref(w.c), int(self.step_count), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                     self._stream()), 'sat_step_set')
        G['graph'].replay()
        self.step_count += 1
        if deferred:
            self.defer_pending += 1
        w.dev_step = self.step_count
        self.launches += G['launches']
        self.last_y_pred = G['y_pred']             # (B,) predictions of this step: a static buffer, overwritten by the next replay
        return G['w'], G['sc_pin']

    def _capture_step(self, key, Xh, lr, betas, eps, opt_kind, l2_emb, dense_mode, dropout_p, seed0, defer=1):
        self.bind()
        B = Xh.shape[0]
        w = self.workspace(B, True)
        self._adam_state(opt_kind)
        G = {'w': w,
             'x_pin': torch.empty(tuple(Xh.shape), dtype=Xh.dtype).pin_memory(),
             'y_pin': torch.empty(B, dtype=torch.float32).pin_memory(),
             'hyper_pin': torch.zeros(4, dtype=torch.float32).pin_memory(),
             'sc_pin': torch.zeros(4, dtype=torch.float32).pin_memory(),
             'Xs': torch.empty(tuple(Xh.shape), dtype=Xh.dtype, device=self.device),
             'ys': torch.empty(B, dtype=torch.float32, device=self.device)}
        l = lib()
        torch.cuda.synchronize()
        w.dev_step = None                        # (graph_train_step writes the counter before the first replay)
        timing, self.timing = self.timing, None
        graph = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=self.device)
        w.c.device_step = 1
        try:
            with torch.cuda.graph(graph, stream=cap):
                w['step_state'][4:8].copy_(G['hyper_pin'], non_blocking=True)
                check(l.sat_step_begin(C.byref(w.c), (int(seed0) - 1) & (2 ** 64 - 1), self._stream()), 'sat_step_begin')
                G['y_pred'], _ = self.train_step(G['Xs'], G['ys'], lr=lr, betas=betas, eps=eps, l2_emb=l2_emb,
                                                 dense_mode=dense_mode, training=True, dropout_p=dropout_p, seed=0,
                                                 opt_kind=opt_kind, _capturing=True, defer=defer)
                G['sc_pin'].copy_(w['scalars'][0:4], non_blocking=True)
        finally:
            w.c.device_step = 0
            self.timing = timing
        G['graph'] = graph
        G['launches'] = self._launch_count(True, dense_mode) + 2 + (1 if key[-1] else 0)
        self._graphs[key] = G
        return G

    def adam_state(self):
        self.sync_tables()
        return self._adam_state(self._adam['kind'] if self._adam else 0)

    def reset_optimizer(self):
        """Fresh optimiser state (moments and step counter), as a new torch.optim instance would have."""
        self.sync_tables()
        self._adam = None
        self.step_count = 0
        self._graphs, self._graph_seen = {}, {}      # captured steps hold the old state pointers
