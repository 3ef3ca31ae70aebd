"""BaseModel: the trainer / scorer half of the drop-in API.

Mirrors `models/meta_basemodel.py:124-686` of the reference (identical public surface to
`models/basemodel.py:125-128`): constructor arguments, `compile`, `fit`, `evaluate`, `predict`,
`input_from_feature_columns`, `compute_input_dim`, `add_regularization_weight`,
`get_regularization_loss`, `add_auxiliary_loss`, and the attributes `feature_index`,
`embedding_dict`, `history`, `metrics_names`, `stop_training`.

What differs underneath: `fit`/`predict` keep the input matrix resident in HBM, run each batch through the
fused sm_100a kernels (satrans_b200/csrc) instead of ~4.7 k ATen dispatches, fuse L2 + Adam into the
kernels, and compute the per-step host metrics from asynchronously copied predictions.
"""
import time

import numpy as np
import torch
import torch.nn as nn

from .inputs import DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features


class History(object):
    """keras-style History callback (the reference takes it from deepctr_torch.callbacks)."""

    def __init__(self):
        self.epoch, self.history, self.model = [], {}, None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        self.epoch, self.history = [], {}

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


class CallbackList(object):
    """The hooks of keras' CallbackList that `fit` drives (reference meta_basemodel.py:283-294,379-383)."""

    def __init__(self, callbacks=None):
        self.callbacks, self.model = list(callbacks or []), None

    def set_model(self, model):
        self.model = model
        for c in self.callbacks:
            if hasattr(c, 'set_model'):
                c.set_model(model)

    def _call(self, name, *a):
        for c in self.callbacks:
            fn = getattr(c, name, None)
            if fn is not None:
                fn(*a)

    def on_train_begin(self, logs=None):
        self._call('on_train_begin', logs)

    def on_train_end(self, logs=None):
        self._call('on_train_end', logs)

    def on_epoch_begin(self, epoch, logs=None):
        self._call('on_epoch_begin', epoch, logs)

    def on_epoch_end(self, epoch, logs=None):
        self._call('on_epoch_end', epoch, logs)


class PredictionLayer(nn.Module):
    """Bias + sigmoid (deepctr PredictionLayer).  SATrans never calls it (reference satrans.py:255 uses
    torch.sigmoid directly) but `out.bias` is part of the state dict."""

    def __init__(self, task='binary', use_bias=True):
        if task not in ["binary", "multiclass", "regression"]:
            raise ValueError("task must be binary,multiclass or regression")
        super().__init__()
        self.use_bias, self.task = use_bias, task
        if use_bias:
            self.bias = nn.Parameter(torch.zeros((1,)))

    def forward(self, X):
        out = X + self.bias if self.use_bias else X
        return torch.sigmoid(out) if self.task == "binary" else out


class Linear(nn.Module):
    """The per-feature dim-1 tables of the linear part (reference meta_basemodel.py:34-92).  SATrans.forward
    never evaluates it (use_linear is ignored, l2_reg_linear=0), there it only carries state-dict entries; the attention
    baselines with use_linear=True call forward() (plain torch lookups: F one-float rows per sample, not a hot op)."""

    def __init__(self, feature_columns, feature_index, init_std=0.0001, device='cpu'):
        super().__init__()
        self.feature_index = feature_index
        sparse = [f for f in feature_columns if isinstance(f, SparseFeat)]
        dense = [f for f in feature_columns if isinstance(f, DenseFeat)]
        self._sparse, self._dense_feats = sparse, dense
        self.embedding_dict = nn.ModuleDict({f.embedding_name: nn.Embedding(f.vocabulary_size, 1) for f in sparse})
        for t in self.embedding_dict.values():
            nn.init.normal_(t.weight, mean=0, std=init_std)
        if dense:
            self.weight = nn.Parameter(torch.empty(sum(f.dimension for f in dense), 1))
            nn.init.normal_(self.weight, mean=0, std=init_std)
        self.to(device)

    def forward(self, X):
        """linear_logit (B, 1) = sum_f w_f[id_f] + dense @ weight (reference basemodel.py:63-92)."""
        logit = torch.zeros((X.shape[0], 1), dtype=torch.float32, device=X.device)
        if self._sparse:
            logit = logit + torch.cat([self.embedding_dict[f.embedding_name](X[:, self.feature_index[f.name][0]].long())
                                       for f in self._sparse], dim=-1).sum(dim=-1, keepdim=True)
        if self._dense_feats:
            cols = [c for f in self._dense_feats for c in range(*self.feature_index[f.name])]
            logit = logit + X[:, cols].float().matmul(self.weight)
        return logit


def _auc(y_true, y_score):
    from sklearn.metrics import roc_auc_score
    return roc_auc_score(y_true, y_score)


def _logloss(y_true, y_score):
    from sklearn.metrics import log_loss
    return log_loss(y_true, y_score)


def _mse(y_true, y_score):
    return float(np.mean((np.asarray(y_true, dtype=np.float64).ravel() - np.asarray(y_score).ravel()) ** 2))


def _acc(y_true, y_score):
    return float(np.mean(np.asarray(y_true).ravel() == np.where(np.asarray(y_score).ravel() > 0.5, 1, 0)))


class _BatchFeeder(object):
    """fit()'s host -> device batch pipeline: the training matrix and labels stay in pinned host memory; batch i + 1 is copied
    (after a host-side gather into pinned staging when shuffling) on a copy stream while step i runs."""

    def __init__(self, X, y, batch_size, dev):
        self.Xh = torch.from_numpy(np.ascontiguousarray(X))
        self.yh = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        if not self.Xh.is_pinned():
            self.Xh = self.Xh.pin_memory()
        if not self.yh.is_pinned():
            self.yh = self.yh.pin_memory()
        self.n, self.B, self.dev = len(self.Xh), int(batch_size), dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.Xs = [torch.empty((self.B, self.Xh.shape[1]), dtype=self.Xh.dtype, device=dev) for _ in range(2)]
        self.ys = [torch.empty(self.B, dtype=torch.float32, device=dev) for _ in range(2)]
        self.x_pin = self.y_pin = None
        self.ready = [torch.cuda.Event() for _ in range(2)]      # batch landed in buffer k
        self.free = [torch.cuda.Event() for _ in range(2)]       # the step that read buffer k has been enqueued and finished with it
        self.copied = [torch.cuda.Event() for _ in range(2)]     # pinned staging k may be overwritten by the host
        self.bytes_h2d = 0

    def start_epoch(self, order_host):
        self.order = order_host
        if order_host is not None and self.x_pin is None:
            self.x_pin = [torch.empty((self.B, self.Xh.shape[1]), dtype=self.Xh.dtype).pin_memory() for _ in range(2)]
            self.y_pin = [torch.empty(self.B, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.i_next, self.i_cur = 0, 0
        self._issue()                                            # batch 0

    def _issue(self):
        """Enqueue the copy of batch i_next (if any) on the copy stream."""
        i = self.i_next
        s0 = i * self.B
        if s0 >= self.n:
            return
        s1 = min(s0 + self.B, self.n)
        k = i & 1
        if self.order is not None:
            if i >= 2:
                self.copied[k].synchronize()                     # (two steps old: long done)
            idx = self.order[s0:s1]
            torch.index_select(self.Xh, 0, idx, out=self.x_pin[k][:s1 - s0])
            torch.index_select(self.yh, 0, idx, out=self.y_pin[k][:s1 - s0])
            xs, ys_ = self.x_pin[k][:s1 - s0], self.y_pin[k][:s1 - s0]
        else:
            xs, ys_ = self.Xh[s0:s1], self.yh[s0:s1]
        with torch.cuda.stream(self.copy_stream):
            if i >= 2:
                self.copy_stream.wait_event(self.free[k])
            self.Xs[k][:s1 - s0].copy_(xs, non_blocking=True)
            self.ys[k][:s1 - s0].copy_(ys_, non_blocking=True)
            self.ready[k].record(self.copy_stream)
            self.copied[k].record(self.copy_stream)
        self.bytes_h2d += xs.numel() * xs.element_size() + ys_.numel() * 4
        self.i_next = i + 1

    def next_batch(self):
        i = self.i_cur
        k = i & 1
        n = min(self.B, self.n - i * self.B)
        torch.cuda.current_stream().wait_event(self.ready[k])
        self._issue()                                            # batch i + 1 crosses PCIe while step i runs
        return self.Xs[k][:n], self.ys[k][:n]

    def batch_consumed(self):
        self.free[self.i_cur & 1].record(torch.cuda.current_stream())
        self.i_cur += 1


class BaseModel(nn.Module):
    def __init__(self, linear_feature_columns, dnn_feature_columns, l2_reg_linear=1e-5, l2_reg_embedding=1e-5,
                 init_std=0.0001, seed=1024, task='binary', device='cpu', gpus=None,
                 flag=None, domain_column=None, num_domains=None, meta_dnn_hidden_units=(32, 64, 32),
                 table_shard=None):
        super().__init__()
        if any(isinstance(f, VarLenSparseFeat) for f in list(linear_feature_columns) + list(dnn_feature_columns)):
            raise NotImplementedError("VarLenSparseFeat is not on the SATrans path (reference main.py:103,146)")
        if domain_column:
            raise NotImplementedError("BaseModel(domain_column=...) is the 'metatrans' hook of the single-task "
                                      "baselines (reference basemodel.py:191-199), outside the SATrans path")
        self.flag = flag
        torch.manual_seed(int(seed))                                  # reference meta_basemodel.py:153
        self.dnn_feature_columns = dnn_feature_columns
        self.device = device
        self.gpus = gpus
        if gpus and str(self.gpus[0]) not in self.device:              # reference meta_basemodel.py:160-162
            raise ValueError("`gpus[0]` should be the same gpu with `device`")
        if gpus and len(gpus) > 1:
            raise NotImplementedError("single-process nn.DataParallel is replaced by one process per GPU: "
                                      "launch with torchrun and use satrans_b200.dist (see DESIGN.md)")
        self.feature_index = build_input_features(list(linear_feature_columns) + list(dnn_feature_columns))
        self.reg_loss = torch.zeros((1,), device=device)
        self.aux_loss = torch.zeros((1,), device=device)

        # ---- embedding tables: ONE [total_rows, d] buffer, every nn.Embedding.weight is a view of it -----
        sparse = [f for f in dnn_feature_columns if isinstance(f, SparseFeat)]
        if len(set(f.embedding_name for f in sparse)) != len(sparse):
            raise NotImplementedError("two SparseFeat columns sharing one embedding_name (one table looked up by two columns, "
                                      "deepctr's shared-embedding form) are not built: every field owns its slice of the table buffer")
        d = self.embedding_size
        self._sparse_feats = sparse
        self._row_off = np.concatenate([[0], np.cumsum([f.vocabulary_size for f in sparse])]).astype(np.int64)
        # table_shard = (rank, world): this process holds only rows g with g % world == rank (satrans_b200/dist.py)
        self._shard = (0, 1) if table_shard is None else (int(table_shard[0]), int(table_shard[1]))
        total_rows = int(self._row_off[-1])
        n_local = (total_rows - self._shard[0] + self._shard[1] - 1) // self._shard[1]
        self._tables = torch.empty((n_local, d), dtype=torch.float32, device=device)
        # Row-sharded: every rank draws ITS rows from its own stream (seed folded with the rank); with one shared stream all
        # ranks would draw the same values and global rows kR .. kR+R-1 would start identical instead of i.i.d.
        gen = None
        if self._shard[1] > 1:
            gen = torch.Generator(device=self._tables.device)
            gen.manual_seed((int(seed) * 1000003 + 7919 * (self._shard[0] + 1)) & (2 ** 63 - 1))
        std = init_std if not (flag and 'noembinit' in flag) else 1.0    # reference meta_basemodel.py:114-119
        self._tables.normal_(mean=0.0, std=std, generator=gen)
        self.embedding_dict = nn.ModuleDict()
        for i, f in enumerate(sparse):
            emb = nn.Embedding(f.vocabulary_size, d, device='meta')
            if self._shard[1] == 1:
                emb.weight = nn.Parameter(self._tables[int(self._row_off[i]):int(self._row_off[i + 1])])
            else:       # rows of one field are strided over the ranks: no per-field view exists (see gather_tables)
                emb.weight = nn.Parameter(torch.empty((0, d), dtype=torch.float32, device=device))
            self.embedding_dict[f.embedding_name] = emb

        self.linear_model = Linear(linear_feature_columns, self.feature_index, device=device)
        self.regularization_weight = []
        self.add_regularization_weight(self.embedding_dict.parameters(), l2=l2_reg_embedding)
        self.add_regularization_weight(self.linear_model.parameters(), l2=l2_reg_linear)
        self.l2_reg_embedding = l2_reg_embedding
        self.out = PredictionLayer(task).to(device)

        self._is_graph_network = True
        self._ckpt_saved_epoch = False
        self.history = History()
        self.stop_training = False
        self._engine = None

    # ---- introspection helpers kept from the reference ------------------------------------------------
    @property
    def embedding_size(self):
        sizes = set(f.embedding_dim for f in self.dnn_feature_columns if isinstance(f, SparseFeat))
        if len(sizes) > 1:
            raise ValueError("embedding_dim of SparseFeat and VarlenSparseFeat must be same in this model!")
        return list(sizes)[0]

    def compute_input_dim(self, feature_columns, include_sparse=True, include_dense=True, feature_group=False):
        sparse = [f for f in feature_columns if isinstance(f, SparseFeat)]
        dense = [f for f in feature_columns if isinstance(f, DenseFeat)]
        dense_dim = sum(f.dimension for f in dense)
        sparse_dim = len(sparse) if feature_group else sum(f.embedding_dim for f in sparse)
        return (sparse_dim if include_sparse else 0) + (dense_dim if include_dense else 0)

    def add_regularization_weight(self, weight_list, l1=0.0, l2=0.0):
        if isinstance(weight_list, torch.nn.parameter.Parameter):
            weight_list = [weight_list]
        else:
            weight_list = list(weight_list)
        self.regularization_weight.append((weight_list, l1, l2))

    def add_auxiliary_loss(self, aux_loss, alpha):
        self.aux_loss = aux_loss * alpha

    def get_regularization_loss(self):
        """l2 * sum(square(w)) over the registered weights as a (1,) tensor (reference
        meta_basemodel.py:577-593).  `fit` never calls this: the fused optimiser sweep produces the same
        number as a by-product; it is kept for API compatibility and evaluated with torch ops."""
        self.sync_tables()
        total = torch.zeros((1,), device=self._tables.device)
        for weight_list, l1, l2 in self.regularization_weight:
            for w in weight_list:
                p = w[1] if isinstance(w, tuple) else w
                if l1 > 0:
                    total = total + torch.sum(l1 * torch.abs(p))
                if l2 > 0:
                    total = total + torch.sum(l2 * torch.square(p))
        return total

    def input_from_feature_columns(self, X, feature_columns, embedding_dict, support_dense=True):
        """(list of F (B,1,d) embeddings, list of (B,k) dense slices) — reference meta_basemodel.py:519-545.
        The F lookups are ONE launch of the multi-field gather kernel; the list entries are views."""
        sparse = [f for f in feature_columns if isinstance(f, SparseFeat)]
        dense = [f for f in feature_columns if isinstance(f, DenseFeat)]
        if not support_dense and len(dense) > 0:
            raise ValueError("DenseFeat is not supported in dnn_feature_columns")
        if embedding_dict is not self.embedding_dict or [f.name for f in sparse] != [f.name for f in self._sparse_feats]:
            raise NotImplementedError("only the model's own embedding_dict / dnn_feature_columns are supported")
        E = self.engine.gather(self._as_device_matrix(X))
        embs = [E[:, i:i + 1, :] for i in range(E.shape[1])]
        dense_vals = [X[:, self.feature_index[f.name][0]:self.feature_index[f.name][1]] for f in dense]
        return embs, dense_vals

    # ---- row-sharded tables (multi-GPU) ---------------------------------------------------------------
    stream_host_batches = None   # fit(): None = keep the training matrix in HBM unless it exceeds 256 MB; True / False force it

    # ---- deferred table sweep (engine.sync_tables): anything that reads or moves the parameters closes the epoch first --------
    table_sweep_interval = 8     # fit(): the every-row table update runs once per this many steps (1 = after every step);
                                 # parameters and optimiser state are bit-identical either way (include/satrans_b200.h, sat_defer_*)

    def sync_tables(self):
        """Bring every table row up to date (no-op unless a deferred sweep epoch is open)."""
        eng = getattr(self, '_engine', None)
        if eng is not None:
            eng.sync_tables()

    def state_dict(self, *args, **kwargs):
        self.sync_tables()
        return super().state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.sync_tables()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.sync_tables()
        return super().load_state_dict(*args, **kwargs)

    def load_full_tables(self, full):
        """Copy this process' rows out of a full [total_rows, d] table (rows g with g % world == rank)."""
        rank, world = self._shard
        self.sync_tables()
        with torch.no_grad():
            self._tables.copy_(torch.as_tensor(full)[rank::world].to(self._tables.device))

    def gather_tables(self, group=None):
        """Full [total_rows, d] table on every rank (all-gather of the shards) — e.g. to write a reference-
        compatible checkpoint from a sharded run."""
        rank, world = self._shard
        self.sync_tables()
        if world == 1:
            return self._tables.detach().clone()
        import torch.distributed as tdist
        total = int(self._row_off[-1])
        n_max = (total + world - 1) // world
        pad = torch.zeros((n_max, self._tables.shape[1]), dtype=torch.float32, device=self._tables.device)
        pad[:self._tables.shape[0]] = self._tables.detach()
        parts = [torch.empty_like(pad) for _ in range(world)]
        tdist.all_gather(parts, pad, group=group)
        full = torch.empty((total, self._tables.shape[1]), dtype=torch.float32, device=self._tables.device)
        for r in range(world):
            n_r = (total - r + world - 1) // world
            full[r::world] = parts[r][:n_r]
        return full

    # ---- engine plumbing ------------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    @property
    def precision(self):
        """Arithmetic of the projection GEMMs: 'tf32x3' (default: tcgen05 tensor cores, error-compensated 3xTF32, y_pred within
        1e-7 of fp32; what bench.py times), 'fp32' (CUDA-core FFMA, tightest parity with the reference) or 'tf32' (plain TF32)."""
        return self.engine.precision

    @precision.setter
    def precision(self, value):
        if value not in ('fp32', 'tf32x3', 'tf32', 'bf16'):
            raise ValueError("precision must be 'fp32', 'tf32x3', 'tf32' or 'bf16', got %r" % (value,))
        from . import _lib
        if value != 'fp32' and self._layout.get('flags', 0) & (_lib.FLAG_GATE_Q | _lib.FLAG_GATE_K | _lib.FLAG_BILINEAR):
            raise NotImplementedError("the 'gate' / 'bilinear' flag variants run in the CUDA-core kernels only (precision='fp32')")
        self.engine.precision = value

    def _as_device_matrix(self, X):
        if not torch.is_tensor(X):
            X = torch.as_tensor(np.ascontiguousarray(self._normalise_matrix(X)))
        if X.dtype not in (torch.float32, torch.int32, torch.int64):
            X = X.long() if not X.dtype.is_floating_point else X.float()
        return X.to(self._tables.device, non_blocking=True).contiguous()

    # ---- compile -----------------------------------------------------------------------------------------
    def compile(self, optimizer, loss=None, metrics=None):
        """Same contract as the reference (meta_basemodel.py:598-610).  The optimiser may be a string or a
        torch.optim instance built over `model.parameters()`; its hyper-parameters are read and the update
        itself runs inside the fused kernels (Adam: tables + dense parameters)."""
        self.metrics_names = ["loss"]
        if getattr(self, '_engine', None) is not None:
            self._engine.reset_optimizer()          # a new optimiser starts from fresh state, like a new torch.optim instance
        self.optim = self._get_optim(optimizer)
        self.loss_func = self._get_loss_func(loss)
        self.metrics = self._get_metrics(metrics)

    def _get_optim(self, optimizer):
        """String or torch.optim instance -> hyper-parameters of the fused update (reference meta_basemodel.py:
        612-625: sgd lr=0.01, adam, adagrad, rmsprop with torch defaults).  Only the plain variants exist in the
        kernels (no momentum / weight decay / amsgrad / centering); anything else raises."""
        from . import _lib
        if isinstance(optimizer, str):
            table = {"sgd": dict(kind=_lib.OPT_SGD, lr=0.01, betas=(0.0, 0.0), eps=0.0),
                     "adam": dict(kind=_lib.OPT_ADAM, lr=1e-3, betas=(0.9, 0.999), eps=1e-8),
                     "adagrad": dict(kind=_lib.OPT_ADAGRAD, lr=0.01, betas=(0.0, 0.0), eps=1e-10),
                     "rmsprop": dict(kind=_lib.OPT_RMSPROP, lr=0.01, betas=(0.0, 0.99), eps=1e-8)}
            if optimizer not in table:
                raise NotImplementedError
            self._opt_hparams = table[optimizer]
            return optimizer
        g = optimizer.param_groups[0] if hasattr(optimizer, 'param_groups') else None
        if isinstance(optimizer, torch.optim.Adam):
            if g.get('amsgrad') or g.get('weight_decay', 0) != 0 or g.get('maximize'):
                raise NotImplementedError("Adam with amsgrad / weight_decay / maximize is not on the SATrans path")
            self._opt_hparams = dict(kind=_lib.OPT_ADAM, lr=g['lr'], betas=tuple(g['betas']), eps=g['eps'])
        elif isinstance(optimizer, torch.optim.SGD):
            if g.get('momentum', 0) != 0 or g.get('weight_decay', 0) != 0 or g.get('nesterov') or g.get('maximize'):
                raise NotImplementedError("SGD with momentum / weight_decay / nesterov is not built")
            self._opt_hparams = dict(kind=_lib.OPT_SGD, lr=g['lr'], betas=(0.0, 0.0), eps=0.0)
        elif isinstance(optimizer, torch.optim.Adagrad):
            if g.get('lr_decay', 0) != 0 or g.get('weight_decay', 0) != 0 or g.get('initial_accumulator_value', 0) != 0:
                raise NotImplementedError("Adagrad with lr_decay / weight_decay / initial accumulator is not built")
            self._opt_hparams = dict(kind=_lib.OPT_ADAGRAD, lr=g['lr'], betas=(0.0, 0.0), eps=g['eps'])
        elif isinstance(optimizer, torch.optim.RMSprop):
            if g.get('momentum', 0) != 0 or g.get('centered') or g.get('weight_decay', 0) != 0:
                raise NotImplementedError("RMSprop with momentum / centered / weight_decay is not built")
            self._opt_hparams = dict(kind=_lib.OPT_RMSPROP, lr=g['lr'], betas=(0.0, g['alpha']), eps=g['eps'])
        else:
            raise NotImplementedError("optimizer must be 'sgd' | 'adam' | 'adagrad' | 'rmsprop' or the matching "
                                      "torch.optim instance (reference meta_basemodel.py:612-625)")
        return optimizer

    def _hparams(self):
        """Hyper-parameters of the next fused update.  With a torch.optim instance the learning rate (and betas / eps) are
        read from its live param_groups[0], so lr schedulers and manual edits take effect like in the reference's loop."""
        hp = dict(self._opt_hparams)
        g = self.optim.param_groups[0] if hasattr(self.optim, 'param_groups') else None
        if g is not None:
            hp['lr'] = g['lr']
            if 'eps' in g and hp['kind'] != 1:
                hp['eps'] = g['eps']
            if 'betas' in g:
                hp['betas'] = tuple(g['betas'])
            elif 'alpha' in g:
                hp['betas'] = (0.0, g['alpha'])
        return hp

    def _get_loss_func(self, loss):
        """'binary_crossentropy' | 'mse' | 'mae' (reference meta_basemodel.py:627-640) are fused into the forward kernel's
        logit head.  Any other callable `loss(y_pred, y, reduction='sum')` (the reference passes it through unchanged, :638-639)
        is evaluated by torch on the (B,) prediction vector between the fused forward and the fused backward
        (Engine.train_step(loss_fn=...))."""
        self._loss_fn = None
        if isinstance(loss, str):
            if loss in ("binary_crossentropy", "mse", "mae"):
                self._loss_kind = loss
                if getattr(self, '_engine', None) is not None:
                    self._engine._bound_ptrs = None       # loss kind travels in the model descriptor
                return loss
            raise NotImplementedError
        if callable(loss):
            self._loss_fn = loss
            return loss
        raise NotImplementedError("loss must be one of the reference's strings or a callable loss(y_pred, y, reduction='sum')")

    def _get_metrics(self, metrics, set_eps=False):
        metrics_ = {}
        if metrics:
            for metric in metrics:
                if metric in ("binary_crossentropy", "logloss"):
                    metrics_[metric] = _logloss
                if metric == "auc":
                    metrics_[metric] = _auc
                if metric == "mse":
                    metrics_[metric] = _mse
                if metric in ("accuracy", "acc"):
                    metrics_[metric] = _acc
                self.metrics_names.append(metric)
        return metrics_

    # ---- data staging ------------------------------------------------------------------------------------
    def _n_cols(self):
        return max(end for (_start, end) in self.feature_index.values())

    @staticmethod
    def _normalise_matrix(X):
        """(N, n_cols) array in any numeric dtype -> float32 (the reference's `.float()`), int32 or int64.  Integer matrices
        stay integer (exact ids); anything else becomes float32, refusing ids the float32 route cannot represent."""
        X = np.asarray(X)
        if X.dtype in (np.float32, np.int32, np.int64):
            return X
        if np.issubdtype(X.dtype, np.integer) or X.dtype == np.bool_:
            return X.astype(np.int64)
        if X.size and np.abs(X).max() >= 2 ** 24 and np.all(X == np.round(X)):
            raise ValueError("ids >= 2**24 cannot travel through the reference's float32 input matrix; pass an integer matrix")
        return X.astype(np.float32)

    def _stack_inputs(self, x):
        """dict / list of 1-D columns (or an already stacked matrix / tensor) -> (N, n_cols) matrix in feature_index order
        (reference meta_basemodel.py:221-222,257-264).  float32 exactly like the reference when any column is non-integer;
        int32 otherwise (exact for ids >= 2**24, which the fp32 route cannot represent)."""
        if torch.is_tensor(x):
            x = x.detach().cpu().numpy()
        if isinstance(x, np.ndarray) and x.ndim == 2:
            if x.shape[1] != self._n_cols():
                raise ValueError('expected an (N, %d) input matrix, got %s' % (self._n_cols(), x.shape))
            return self._normalise_matrix(x)
        if isinstance(x, dict):
            x = [x[feature] for feature in self.feature_index]
        cols = [np.asarray(c) for c in x]
        cols = [c.reshape(len(c), -1) for c in cols]
        all_int = all(np.issubdtype(c.dtype, np.integer) for c in cols)
        if all_int and all(c.size == 0 or (c.min() >= -2 ** 31 and c.max() < 2 ** 31) for c in cols):
            return np.concatenate([c.astype(np.int32, copy=False) for c in cols], axis=-1)
        for c in cols:
            if np.issubdtype(c.dtype, np.integer) and c.size and c.max() >= 2 ** 24:
                raise ValueError("ids >= 2**24 cannot travel through the reference's float32 input matrix; "
                                 "pass integer columns only (no dense features) to use the int32 path")
        return np.concatenate([c.astype(np.float32, copy=False) for c in cols], axis=-1)

    # ---- fit ---------------------------------------------------------------------------------------------
    def fit(self, x=None, y=None, batch_size=None, epochs=1, valid_cnt_per_epoch=1, verbose=1, initial_epoch=0,
            validation_split=0., validation_data=None, shuffle=True, callbacks=None):
        """Same signature / History contract as the reference (meta_basemodel.py:200-385)."""
        if isinstance(x, dict) and getattr(self, 'domain_column_list', None):
            self.domain_id_offset = x[self.domain_column_list[0]].min()        # reference :218
        do_validation = False
        val_x, val_y = [], []
        if validation_data:
            do_validation = True
            if len(validation_data) == 2:
                val_x, val_y = validation_data
            elif len(validation_data) == 3:
                val_x, val_y, _ = validation_data
            else:
                raise ValueError('When passing a `validation_data` argument, it must contain either 2 items '
                                 '(x_val, y_val), or 3 items (x_val, y_val, val_sample_weights)')
        X = self._stack_inputs(x)
        y = np.asarray(y).reshape(len(X), -1)[:, 0]
        if not validation_data and validation_split and 0. < validation_split < 1.:
            do_validation = True
            split_at = int(len(X) * (1. - validation_split))
            X, val_x = X[:split_at], X[split_at:]
            y, val_y = y[:split_at], y[split_at:]
        if batch_size is None:
            batch_size = 256
        self.train()
        eng = self.engine
        dev = self._tables.device
        sample_num = len(X)
        steps_per_epoch = (sample_num - 1) // batch_size + 1
        steps_to_valid = steps_per_epoch // valid_cnt_per_epoch + 1

        callbacks = CallbackList((callbacks or []) + [self.history])
        callbacks.set_model(self)
        callbacks.on_train_begin()
        self.stop_training = False

        want_metrics = verbose > 0 and len(self.metrics) > 0
        # Small training sets live in HBM for the whole fit().  Large ones (or stream_host_batches = True) stay in pinned host
        # memory, like the reference's DataLoader over CPU tensors with its per-batch .to(device) (meta_basemodel.py:311-313): every
        # step's batch crosses PCIe on a copy stream ONE STEP AHEAD of the step that consumes it (two device buffers), so the copy
        # rides under the previous step's kernels and the host never waits for the device inside an epoch.
        stream_batches = self.stream_host_batches
        if stream_batches is None:
            stream_batches = X.nbytes > (256 << 20)
        feeder = None
        if stream_batches:
            feeder = _BatchFeeder(X, y, batch_size, dev)
            Xd = None
            yd = feeder.yh.to(dev) if want_metrics else None
        else:
            Xd = torch.as_tensor(np.ascontiguousarray(X)).to(dev)
            yd = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
        if verbose > 0:
            print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(
                sample_num, len(val_y), steps_per_epoch))
        if getattr(self, '_shuffle_gen', None) is None:        # one stream per model: repeated fit() calls keep drawing from it
            self._shuffle_gen = torch.Generator(device='cpu')
            self._shuffle_gen.manual_seed(torch.initial_seed() & (2 ** 63 - 1))
        gen = self._shuffle_gen
        # the every-row table update runs once per `table_sweep_interval` steps (bit-identical parameters, engine.train_step);
        # the regulariser term of a step is then known when its sweep epoch closes: the epoch's loss takes it from there
        defer = int(self.table_sweep_interval) if (self.dense_table_updates and eng.dist is None) else 1
        for epoch in range(initial_epoch, epochs):
            callbacks.on_epoch_begin(epoch)
            epoch_logs = {}
            start_time = time.time()
            reg0 = None
            if defer > 1:
                eng.sync_tables()
                reg0 = eng.deferred_reg_total()
            order_host = torch.randperm(sample_num, generator=gen) if shuffle else None
            order = order_host.to(dev) if (shuffle and (feeder is None or want_metrics)) else None
            preds = torch.empty(sample_num, dtype=torch.float32, device=dev) if want_metrics else None
            losses = torch.zeros((steps_per_epoch, 2), dtype=torch.float32, device=dev)
            step_num = 0
            if feeder is not None:
                feeder.start_epoch(order_host)
            for s0 in range(0, sample_num, batch_size):
                s1 = min(s0 + batch_size, sample_num)
                if feeder is not None:
                    xb, yb = feeder.next_batch()
                elif order is not None:
                    idx = order[s0:s1]
                    xb, yb = Xd.index_select(0, idx).contiguous(), yd.index_select(0, idx).contiguous()
                else:
                    xb, yb = Xd[s0:s1].contiguous(), yd[s0:s1].contiguous()
                hp = self._hparams()
                seed0 = int(torch.initial_seed()) & (2 ** 63 - 1)
                loss_fn = getattr(self, '_loss_fn', None)
                if eng.use_cuda_graph and eng.dist is None and eng.step_count >= 2 and (s1 - s0) == batch_size and loss_fn is None:
                    # full-size batches after the first two steps: the whole step is ONE CUDA-graph launch (engine.graph_train_step)
                    w, _ = eng.graph_train_step(xb, yb, lr=hp['lr'], betas=hp['betas'], eps=hp['eps'], opt_kind=hp['kind'],
                                                l2_emb=self.l2_reg_embedding, dense_mode=self.dense_table_updates,
                                                dropout_p=self.dropout_rate, seed0=seed0, defer=defer)
                    y_pred = eng.last_y_pred
                else:
                    y_pred, w = eng.train_step(xb, yb, lr=hp['lr'], betas=hp['betas'], eps=hp['eps'], opt_kind=hp['kind'],
                                               l2_emb=self.l2_reg_embedding, dense_mode=self.dense_table_updates,
                                               training=True, dropout_p=self.dropout_rate,
                                               seed=(seed0 + eng.step_count) & (2 ** 63 - 1), defer=defer, loss_fn=loss_fn)
                losses[step_num, 0:(1 if defer > 1 else 2)].copy_(w['scalars'][0:(1 if defer > 1 else 2)])
                if feeder is not None:
                    feeder.batch_consumed()
                if want_metrics:
                    preds[s0:s1] = y_pred
                step_num += 1
                if valid_cnt_per_epoch > 1 and step_num % steps_to_valid == 0 and do_validation:
                    eval_result = self.evaluate(val_x, val_y, batch_size)
                    self.train()
                    print('Step: %d/%d, %s' % (step_num, steps_per_epoch, ''.join(
                        " - %s: %.4f" % (k, v) for k, v in eval_result.items())))
            eng.check_errors(eng.workspace(1, True))
            lh = losses.double().cpu().numpy()
            reg_sum = 0.0
            if defer > 1:                               # every table row up to date before metrics / validation / callbacks
                eng.sync_tables()
                reg_sum = float(eng.deferred_reg_total()) - float(reg0)
            epoch_logs["loss"] = (float(lh.sum()) + reg_sum) / sample_num     # (bce + reg) summed / N (:353)
            if want_metrics and getattr(self, 'device_metrics', True) and batch_size <= 16384:
                # per-batch sklearn-equivalent metrics on the device (sat_batch_metrics): one CTA per training batch
                from . import _lib
                yo = (yd.index_select(0, order) if order is not None else yd).contiguous()
                offs = torch.tensor(list(range(0, sample_num, batch_size)) + [sample_num], dtype=torch.int32, device=dev)
                res = torch.empty((steps_per_epoch, 4), dtype=torch.float64, device=dev)
                _lib.check(_lib.lib().sat_batch_metrics(preds.data_ptr(), yo.data_ptr(), offs.data_ptr(), steps_per_epoch,
                                                        min(batch_size, sample_num), res.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream), 'sat_batch_metrics')
                res = res.cpu().numpy()
                col = {"binary_crossentropy": 0, "logloss": 0, "auc": 1, "mse": 2, "accuracy": 3, "acc": 3}
                for name in self.metrics:
                    epoch_logs[name] = np.sum(res[:, col[name]]) / steps_per_epoch
            elif want_metrics:
                ph = preds.cpu().numpy().astype("float64")
                yh = (yd.index_select(0, order) if order is not None else yd).cpu().numpy()
                train_result = {name: [] for name in self.metrics}
                for s0 in range(0, sample_num, batch_size):
                    s1 = min(s0 + batch_size, sample_num)
                    for name, fn in self.metrics.items():
                        try:
                            train_result[name].append(fn(yh[s0:s1], ph[s0:s1]))
                        except ValueError:          # e.g. AUC of a single-class batch: the reference would raise
                            train_result[name].append(float('nan'))
                for name, result in train_result.items():
                    epoch_logs[name] = np.sum(result) / steps_per_epoch
            if do_validation:
                eval_result = self.evaluate(val_x, val_y, batch_size)
                self.train()
                for name, result in eval_result.items():
                    epoch_logs["val_" + name] = result
            if verbose > 0:
                epoch_time = int(time.time() - start_time)
                print('Epoch {0}/{1}'.format(epoch + 1, epochs))
                eval_str = "{0}s - loss: {1: .4f}".format(epoch_time, epoch_logs["loss"])
                for name in self.metrics:
                    if name in epoch_logs:
                        eval_str += " - " + name + ": {0: .4f}".format(epoch_logs[name])
                if do_validation:
                    for name in self.metrics:
                        eval_str += " - val_" + name + ": {0: .4f}".format(epoch_logs["val_" + name])
                print(eval_str)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def train_on_batch(self, x, y, defer_reg=False):
        """One optimiser step on one batch given as HOST data (dict / list of columns or an (B, n_cols) array)
        and labels; returns (bce_sum, reg_loss) as Python floats — the per-step body of `fit`
        (reference meta_basemodel.py:311-328, incl. its `.item()` synchronisation).

        defer_reg=True: the every-row table update runs once per `table_sweep_interval` steps, like inside `fit` (parameters and
        optimiser state bit-identical); the step's regulariser term is then not known yet and (bce_sum, None) is returned --
        `sync_tables()` closes the sweep epoch, `engine.deferred_reg()` has the per-step terms afterwards."""
        defer = int(self.table_sweep_interval) if (defer_reg and self.dense_table_updates) else 1
        if torch.is_tensor(x) and x.dtype in (torch.float32, torch.int32, torch.int64):
            X = x                                           # (pinned) host or device tensor in a kernel-ready dtype: no host pass
        else:
            X = self._stack_inputs(x)
        dev = self._tables.device
        hp = self._hparams()
        eng = self.engine
        self.train()
        seed0 = int(torch.initial_seed()) & (2 ** 63 - 1)
        host_batch = not (torch.is_tensor(X) and X.is_cuda)
        gkey = (tuple(X.shape), str(X.dtype), hp['kind'])
        seen = eng._graph_seen.get(gkey, 0)
        eng._graph_seen[gkey] = seen + 1
        loss_fn = getattr(self, '_loss_fn', None)
        if host_batch and eng.use_cuda_graph and eng.dist is None and seen >= 2 and loss_fn is None:
            # third and later batches of this shape: the step (H2D copies included) is ONE CUDA-graph launch
            w, sc_pin = eng.graph_train_step(X, np.asarray(y, dtype=np.float32) if not torch.is_tensor(y) else y,
                                             lr=hp['lr'], betas=hp['betas'], eps=hp['eps'], opt_kind=hp['kind'],
                                             l2_emb=self.l2_reg_embedding, dense_mode=self.dense_table_updates,
                                             dropout_p=self.dropout_rate, seed0=seed0, defer=defer)
            torch.cuda.current_stream().synchronize()     # like the reference's .item(): the step is done, the scalars are on the host
            sc = sc_pin.tolist()
            if sc[3] != 0:
                eng.check_errors(w)
            return sc[0], (sc[1] if defer <= 1 else None)
        else:
            Xd = torch.as_tensor(X).to(dev, non_blocking=True)
            yd = torch.as_tensor(y).to(dev, non_blocking=True).float().reshape(-1)
            _, w = eng.train_step(Xd.contiguous(), yd.contiguous(), lr=hp['lr'], betas=hp['betas'], eps=hp['eps'],
                                  opt_kind=hp['kind'], l2_emb=self.l2_reg_embedding, dense_mode=self.dense_table_updates,
                                  training=True, dropout_p=self.dropout_rate, seed=(seed0 + eng.step_count) & (2 ** 63 - 1),
                                  defer=defer, loss_fn=loss_fn)
        sc = w['scalars'][0:4].tolist()                 # device -> host: synchronises, like the reference's .item()
        if sc[3] != 0:
            eng.check_errors(w)
        return sc[0], (sc[1] if defer <= 1 else None)

    # ---- evaluate / predict ------------------------------------------------------------------------------
    def evaluate(self, x, y, batch_size=256):
        pred_ans = self.predict(x, batch_size, y)
        return {name: fn(y, pred_ans) for name, fn in self.metrics.items()}

    def predict(self, x, batch_size=256, y=None, domain_ids=None):
        """(N, 1) float64 predictions, eval mode (reference meta_basemodel.py:401-517; the 'showattn' /
        'instattn' analysis dumps are not part of the scoring path and are not implemented)."""
        self.eval()
        X = self._stack_inputs(x)
        dev = self._tables.device
        eng = self.engine
        n = len(X)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        Xd = torch.as_tensor(np.ascontiguousarray(X)).to(dev)
        for s0 in range(0, n, batch_size):
            s1 = min(s0 + batch_size, n)
            out[s0:s1] = eng.forward(Xd[s0:s1].contiguous(), training=False)
        if n:
            eng.check_errors(eng.workspace(1, False))
        return out.cpu().numpy().astype("float64").reshape(-1, 1)

    def _in_multi_worker_mode(self):
        return None
