"""The attention baselines of BASELINE config 5 ("next" rows 2 and 3) through this framework's kernels:
multi-field embedding gather -> field-attention layers -> logit head.  Same constructors, state-dict keys and
compile / fit / evaluate / predict contract as the reference classes.

Scoring (predict / forward under no_grad) launches only this package's kernels.  Training (fit / train_on_batch, reference
basemodel.py:201-384: BCE(sum) + L2(embeddings), every parameter stepped by the torch.optim instance) is the reference's loop with
the hot ops replaced: the gather and the attention layers run their forward AND backward in the CUDA kernels (autograd
Functions around sat_gather_fwd and sat_field_attention_fwd / _bwd; the embedding backward is a sorted-index segmented sum, no
float atomics), while the 1-row logit head, AutoInt's plain DNN tower, the loss and the optimiser update are torch library ops.
SATrans (satrans.py) is the model whose whole step is fused; these two are trained for parity and completeness, not speed.

    Star_Trans   reference models/star_trans.py:674-775 (SelfAttention_Layer x att_layer_num, MDR_InteractingLayer x
                 domain_att_layer_num, dnn_linear), pinned by tests/golden/star_trans.npz
    AutoInt      reference models/autoint.py:39-123 on deepctr-torch's InteractingLayer / DNN -- PARITY UNPINNED (the layers live in
                 an un-vendored third-party package); its plain DNN tower runs as torch.nn.Linear (library GEMM)
"""
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as Fnn

from . import _lib
from .basemodel import BaseModel, CallbackList
from .inputs import DenseFeat, SparseFeat
from .selfatt import InteractingLayer, MDR_InteractingLayer, SelfAttention_Layer


class _EmbeddingRows(torch.autograd.Function):
    """E (B, F, 32) = the batch's rows of every field's table (sat_gather_fwd).  Backward = the reference's dense embedding
    gradient (nn.Embedding(sparse=False)): rows sorted by global index, equal rows summed in batch order (torch.segment_reduce:
    sequential per segment, so two runs agree bit for bit), one dense gradient per table returned as views of one buffer."""

    @staticmethod
    def forward(ctx, model, Xd, *weights):
        ctx.model = model
        ctx.save_for_backward(Xd)
        return model.engine.gather(Xd)

    @staticmethod
    def backward(ctx, dE):
        model = ctx.model
        (Xd,) = ctx.saved_tensors
        lay = model._layout
        dev = dE.device
        ids = Xd[:, lay['sparse_cols']].long()                                          # the reference's .long(): truncation
        rows = (ids + torch.as_tensor(lay['row_off'][:-1], device=dev)[None, :]).reshape(-1)
        srt, order = torch.sort(rows, stable=True)
        uniq, counts = torch.unique_consecutive(srt, return_counts=True)
        seg = torch.segment_reduce(dE.reshape(-1, dE.shape[-1]).float().index_select(0, order), 'sum', lengths=counts)
        g = torch.zeros_like(model._tables)
        g[uniq] = seg
        off = lay['row_off']
        return (None, None) + tuple(g[off[i]:off[i + 1]] for i in range(len(off) - 1))


class _AttentionBaseline(BaseModel):
    """Shared plumbing of the two baselines: gather through the engine's kernel, scoring through the package's kernels only,
    training as the reference's loop around the differentiable kernels."""

    def _field_embeddings(self, Xd):
        """(B, all sparse fields, 32); differentiable w.r.t. the tables when autograd is recording."""
        if torch.is_grad_enabled():
            return _EmbeddingRows.apply(self, Xd, *[self.embedding_dict[f.embedding_name].weight for f in self._sparse_feats])
        return self.engine.gather(Xd)

    # ---- compile: real torch.optim instances (reference basemodel.py:540-582) ----------------------------------------------
    def compile(self, optimizer, loss=None, metrics=None):
        self.metrics_names = ["loss"]
        if isinstance(optimizer, str):
            table = {"sgd": lambda p: torch.optim.SGD(p, lr=0.01), "adam": torch.optim.Adam, "adagrad": torch.optim.Adagrad,
                     "rmsprop": torch.optim.RMSprop}
            if optimizer not in table:
                raise NotImplementedError
            self.optim = table[optimizer](self.parameters())
        else:
            self.optim = optimizer
        if isinstance(loss, str):
            table = {"binary_crossentropy": Fnn.binary_cross_entropy, "mse": Fnn.mse_loss, "mae": Fnn.l1_loss}
            if loss not in table:
                raise NotImplementedError
            self.loss_func = table[loss]
        else:
            self.loss_func = loss
        self.metrics = self._get_metrics(metrics)

    def _train_step(self, Xd, yd):
        """One step of the reference's loop (basemodel.py:309-328) on device tensors -> (y_pred, loss, reg) detached."""
        self.optim.zero_grad()
        y_pred = self.forward(Xd).squeeze()
        loss = self.loss_func(y_pred, yd.squeeze(), reduction='sum')
        reg = self.get_regularization_loss()
        (loss + reg + self.aux_loss.to(loss.device)).sum().backward()
        self.optim.step()
        return y_pred.detach(), loss.detach(), reg.detach()

    def train_on_batch(self, x, y):
        """One optimiser step on one batch given as host data -> (loss, reg) floats, like SATrans.train_on_batch."""
        self.train()
        Xd = self._as_device_matrix(self._stack_inputs(x))
        yd = torch.as_tensor(np.asarray(y, dtype=np.float32).reshape(-1)).to(Xd.device)
        with torch.enable_grad():
            _, loss, reg = self._train_step(Xd, yd)
        self.engine.check_errors(self.engine.workspace(1, False))
        return float(loss), float(reg)

    def fit(self, x=None, y=None, batch_size=None, epochs=1, valid_cnt_per_epoch=1, verbose=1, initial_epoch=0,
            validation_split=0., validation_data=None, shuffle=True, callbacks=None):
        """Same signature / History contract as the reference (basemodel.py:201-384)."""
        do_validation, val_x, val_y = False, [], []
        if validation_data:
            do_validation = True
            if len(validation_data) not in (2, 3):
                raise ValueError('When passing a `validation_data` argument, it must contain either 2 items '
                                 '(x_val, y_val), or 3 items (x_val, y_val, val_sample_weights)')
            val_x, val_y = validation_data[0], validation_data[1]
        X = self._stack_inputs(x)
        y = np.asarray(y).reshape(len(X), -1)[:, 0]
        if not validation_data and validation_split and 0. < validation_split < 1.:
            do_validation = True
            split_at = int(len(X) * (1. - validation_split))
            X, val_x, y, val_y = X[:split_at], X[split_at:], y[:split_at], y[split_at:]
        batch_size = 256 if batch_size is None else batch_size
        dev = self._tables.device
        Xd = torch.as_tensor(np.ascontiguousarray(X)).to(dev)
        yd = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
        n = len(X)
        steps = (n - 1) // batch_size + 1
        callbacks = CallbackList((callbacks or []) + [self.history])
        callbacks.set_model(self)
        callbacks.on_train_begin()
        self.stop_training = False
        if verbose > 0:
            print("Train on {0} samples, validate on {1} samples, {2} steps per epoch".format(n, len(val_y), steps))
        if getattr(self, '_shuffle_gen', None) is None:
            self._shuffle_gen = torch.Generator(device='cpu')
            self._shuffle_gen.manual_seed(torch.initial_seed() & (2 ** 63 - 1))
        for epoch in range(initial_epoch, epochs):
            callbacks.on_epoch_begin(epoch)
            t0 = time.time()
            self.train()
            order = torch.randperm(n, generator=self._shuffle_gen).to(dev) if shuffle else None
            total = torch.zeros((), dtype=torch.float64, device=dev)
            preds = torch.empty(n, dtype=torch.float32, device=dev)
            for s0 in range(0, n, batch_size):
                s1 = min(s0 + batch_size, n)
                if order is not None:
                    xb, yb = Xd.index_select(0, order[s0:s1]).contiguous(), yd.index_select(0, order[s0:s1])
                else:
                    xb, yb = Xd[s0:s1].contiguous(), yd[s0:s1]
                with torch.enable_grad():
                    yp, loss, reg = self._train_step(xb, yb)
                total += loss.double() + reg.double().sum()
                preds[s0:s1] = yp.reshape(-1)
            self.engine.check_errors(self.engine.workspace(1, False))
            epoch_logs = {"loss": float(total) / n}
            if verbose > 0 and self.metrics:
                ph = preds.cpu().numpy().astype("float64")
                yh = (yd.index_select(0, order) if order is not None else yd).cpu().numpy()
                for name, fn in self.metrics.items():
                    vals = []
                    for s0 in range(0, n, batch_size):
                        try:
                            vals.append(fn(yh[s0:s0 + batch_size], ph[s0:s0 + batch_size]))
                        except ValueError:
                            vals.append(float('nan'))
                    epoch_logs[name] = np.sum(vals) / steps
            if do_validation:
                for name, result in self.evaluate(val_x, val_y, batch_size).items():
                    epoch_logs["val_" + name] = result
            if verbose > 0:
                print('Epoch {0}/{1}'.format(epoch + 1, epochs))
                msg = "{0}s - loss: {1: .4f}".format(int(time.time() - t0), epoch_logs["loss"])
                for name in self.metrics:
                    msg += " - " + name + ": {0: .4f}".format(epoch_logs.get(name, float('nan')))
                    if do_validation:
                        msg += " - val_" + name + ": {0: .4f}".format(epoch_logs["val_" + name])
                print(msg)
            callbacks.on_epoch_end(epoch, epoch_logs)
            if self.stop_training:
                break
        callbacks.on_train_end()
        return self.history

    def _init_layout(self, device):
        sparse = self._sparse_feats
        self._layout = dict(F=len(sparse), L=0, H=1, n_dom=1, n_dense=0, n_cols=self._n_cols(), enc_in=32, flags=0,
                            off_dom=0, off_lid=0, off_qkv=0, off_encw=0, off_encb=0, off_layers=0, off_headw=0, off_headb=0,
                            n_dense_params=4, total_rows=int(self._row_off[-1]), row_off=[int(v) for v in self._row_off],
                            sparse_cols=[self.feature_index[f.name][0] for f in sparse], dense_cols=[], dom_col=0)
        self._dense = torch.zeros(4, dtype=torch.float32, device=device)

    def _ensure_flat(self):
        dev = self.out.bias.device
        if self._tables.device != dev:
            tables = self._tables.to(dev)
            for f, a, b in zip(self._sparse_feats, self._row_off[:-1], self._row_off[1:]):
                self.embedding_dict[f.embedding_name].weight.data = tables[int(a):int(b)]
            self._tables = tables
        if self._dense.device != dev:
            self._dense = self._dense.to(dev)

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        if hasattr(self, '_layout'):
            self._ensure_flat()
            self.device = str(self.out.bias.device)
        return out

    def _gather_fields(self, X, keep):
        Xd = self._as_device_matrix(X)
        E = self._field_embeddings(Xd)                               # (B, all sparse fields, 32): one launch
        if len(keep) != E.shape[1]:
            E = E[:, keep, :].contiguous()
        return Xd, E

    def _head(self, feat, dense, w, bias, sigmoid=True):
        if torch.is_grad_enabled():                                  # training: the 1-row head through torch autograd
            z = feat if dense is None else torch.cat([feat, dense], dim=1)
            z = Fnn.linear(z, w, None)
            if bias is not None:
                z = z + bias
            return torch.sigmoid(z) if sigmoid else z
        B, n = feat.shape
        y = torch.empty(B, dtype=torch.float32, device=feat.device)
        nd = 0 if dense is None else dense.shape[1]
        _lib.check(_lib.lib().sat_head_fwd(feat.data_ptr(), n, dense.data_ptr() if nd else None, nd, nd,
                                           w.detach().contiguous().data_ptr(), bias.detach().data_ptr() if bias is not None else None,
                                           B, y.data_ptr(), int(sigmoid), torch.cuda.current_stream().cuda_stream), 'sat_head_fwd')
        return y.unsqueeze(1)

    def predict(self, x, batch_size=256, y=None, domain_ids=None):
        self.eval()
        X = self._stack_inputs(x)
        n = len(X)
        out = torch.empty(n, dtype=torch.float32, device=self._tables.device)
        Xd = torch.as_tensor(np.ascontiguousarray(X)).to(self._tables.device)
        with torch.no_grad():
            for s0 in range(0, n, batch_size):
                s1 = min(s0 + batch_size, n)
                out[s0:s1] = self.forward(Xd[s0:s1].contiguous())[:, 0]
        if n:
            self.engine.check_errors(self.engine.workspace(1, False))
        return out.cpu().numpy().astype("float64").reshape(-1, 1)


class Star_Trans(_AttentionBaseline):
    def __init__(self, linear_feature_columns, dnn_feature_columns, domain_column, num_domains, domain_id_as_feature=False,
                 att_layer_num=2, domain_att_layer_num=1, att_head_num=2, merge='no', att_layer_type='deepctr',
                 use_domain_dnn_linear=False, att_res=True, use_linear=True, use_dnn=False, dnn_hidden_units=(256, 128),
                 dnn_activation='relu', l2_reg_dnn=0, l2_reg_embedding=1e-5, dnn_use_bn=False, dnn_dropout=0, init_std=0.0001,
                 seed=1024, task='binary', device='cpu', gpus=None):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=0, l2_reg_embedding=l2_reg_embedding,
                         init_std=init_std, seed=seed, task=task, device=device, gpus=gpus)
        if att_layer_type != 'deepctr':
            raise NotImplementedError("att_layer_type='%s' (Transformer_Layer, reference star_trans.py:449-509) is not built" % att_layer_type)
        self.use_linear, self.use_dnn, self.num_domains = use_linear, use_dnn, num_domains
        self.use_domain_dnn_linear, self.domain_column = use_domain_dnn_linear, domain_column
        d = self.embedding_size
        names = [f.name for f in self._sparse_feats]
        if domain_id_as_feature:
            self._keep = list(range(len(names)))
        else:                                                            # reference star_trans.py:702-704
            self._keep = [i for i, n in enumerate(names) if n != domain_column]
        field_num = len(self._keep)
        dense = [f for f in linear_feature_columns if isinstance(f, DenseFeat)]
        self._dense_cols = [c for f in dense for c in range(*self.feature_index[f.name])]
        self.att_layer_num, self.dnn_hidden_units = att_layer_num, dnn_hidden_units
        self.int_layers = nn.ModuleList([SelfAttention_Layer(d, att_head_num, att_res, device='cpu', key_name='W_key')
                                         for _ in range(att_layer_num)])
        self.domain_int_layers = nn.ModuleList([MDR_InteractingLayer(d, num_domains, att_head_num, merge, att_res, device='cpu')
                                                for _ in range(domain_att_layer_num)])
        if use_domain_dnn_linear:                                        # reference star_trans.py:728-731: one head per scenario
            self.domain_dnn_linears = nn.ModuleList([nn.Linear(field_num * d + len(self._dense_cols), 1) for _ in range(num_domains)])
        else:
            self.dnn_linear = nn.Linear(field_num * d + len(self._dense_cols), 1)
        self._init_layout(device)
        self.to(device)

    def forward(self, X):
        Xd, x = self._gather_fields(X, self._keep)
        dom = Xd[:, self.feature_index[self.domain_column][0]].to(torch.int32)      # .long() of the reference: truncation
        for layer in self.int_layers:
            x = layer(x)
        for layer in self.domain_int_layers:
            x = layer(x, dom)
        dense = Xd[:, self._dense_cols].float().contiguous() if self._dense_cols else None
        feat = x.reshape(x.shape[0], -1)
        if not self.use_linear and not self.use_domain_dnn_linear:       # main.py's configuration: one fused head launch
            return self._head(feat, dense, self.dnn_linear.weight, self.dnn_linear.bias)
        # the optional logit terms are small torch ops (a dim-1 lookup per field / a per-scenario weight row)
        if self.use_domain_dnn_linear:                                   # reference star_trans.py:762-767
            stack = feat if dense is None else torch.cat([feat, dense], dim=1)
            W = torch.cat([lin.weight for lin in self.domain_dnn_linears], dim=0)          # (n, in)
            bvec = torch.cat([lin.bias for lin in self.domain_dnn_linears], dim=0)         # (n,)
            d64 = dom.long()
            inside = (d64 >= 1) & (d64 <= self.num_domains)
            sel = (d64 - 1).clamp(0, self.num_domains - 1)
            logit = ((stack * W.index_select(0, sel)).sum(dim=1) + bvec.index_select(0, sel)) * inside.to(stack.dtype)
            logit = logit.unsqueeze(1)
        else:
            logit = self._head(feat, dense, self.dnn_linear.weight, self.dnn_linear.bias, sigmoid=False)
        if self.use_linear:                                              # reference star_trans.py:749-750
            logit = logit + self.linear_model(Xd)
        return torch.sigmoid(logit)


class _DNN(nn.Module):
    """deepctr-torch DNN with its state-dict keys (`linears.i.*`): Linear -> activation (-> dropout) for EVERY layer."""

    def __init__(self, inputs_dim, hidden_units, init_std=0.0001):
        super().__init__()
        units = [inputs_dim] + list(hidden_units)
        self.linears = nn.ModuleList([nn.Linear(units[i], units[i + 1]) for i in range(len(units) - 1)])
        for name, t in self.linears.named_parameters():
            if 'weight' in name:
                nn.init.normal_(t, mean=0, std=init_std)

    def forward(self, x):
        for lin in self.linears:
            x = torch.relu(lin(x))
        return x


class AutoInt(_AttentionBaseline):
    def __init__(self, linear_feature_columns, dnn_feature_columns, att_layer_num=3, att_head_num=2, att_res=True,
                 dnn_hidden_units=(256, 128), dnn_activation='relu', l2_reg_dnn=0, l2_reg_embedding=1e-5, dnn_use_bn=False,
                 dnn_dropout=0, init_std=0.0001, seed=1024, task='binary', device='cpu', gpus=None, flag=None,
                 domain_column=None, num_domains=None, meta_dnn_hidden_units=(32, 64, 32)):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=0, l2_reg_embedding=l2_reg_embedding,
                         init_std=init_std, seed=seed, task=task, device=device, gpus=gpus, flag=flag)
        self.flag = flag if flag is not None else ''
        self.domain_column = domain_column
        if domain_column:
            # the "meta module" every reference baseline carries when main.py hands it a scenario column (basemodel.py:132-148):
            # part of the checkpoint; evaluated only with 'usemetatrans' in flag (meta_transformation below)
            from .satrans import DNN_v2, MetaNet
            units = [int(u) for u in meta_dnn_hidden_units]
            self.domain_embeddings = nn.Embedding(num_domains + 1, self.embedding_size)
            self.meta_dnn_hidden_units = units
            self.domain_map_dnn = DNN_v2(self.embedding_size, [sum(units[i] * units[i + 1] for i in range(len(units) - 1))])
            self.meta_net = MetaNet(self.embedding_size, 0.1, 'metanorm' in self.flag, units)
        elif 'usemetatrans' in self.flag:
            raise ValueError("'usemetatrans' needs domain_column / num_domains")
        if len(dnn_hidden_units) <= 0 and att_layer_num <= 0:
            raise ValueError("Either hidden_layer or att_layer_num must > 0")
        if dnn_activation != 'relu' or dnn_use_bn:
            raise NotImplementedError("AutoInt tower: relu without batch norm (reference main.py:235-241 defaults)")
        d = self.embedding_size
        field_num = len(self._sparse_feats)
        dense = [f for f in dnn_feature_columns if isinstance(f, DenseFeat)]
        self._dense_cols = [c for f in dense for c in range(*self.feature_index[f.name])]
        self.use_dnn = len(dnn_feature_columns) > 0 and len(dnn_hidden_units) > 0
        self.dnn_hidden_units, self.att_layer_num = dnn_hidden_units, att_layer_num
        n_in = (dnn_hidden_units[-1] if len(dnn_hidden_units) else 0) + (field_num * d if att_layer_num > 0 else 0)
        self.dnn_linear = nn.Linear(n_in, 1, bias=False)
        if self.use_dnn:
            self.dnn = _DNN(field_num * d + len(self._dense_cols), dnn_hidden_units, init_std)
        self.int_layers = nn.ModuleList([InteractingLayer(d, att_head_num, att_res, device='cpu') for _ in range(att_layer_num)])
        self._init_layout(device)
        self.to(device)

    def meta_transformation(self, Xd, x):
        """BaseModel.meta_transformation (reference basemodel.py:191-199) + MetaNet.forward (submodules.py:77-103) on the gathered
        embeddings: x <- [LN](dropout(MLP_theta(x)) + x) with the MLP weights theta generated from the sample's scenario embedding.
        The scenario encoder runs once per scenario, not per sample; the per-sample products are torch.bmm (library GEMMs:
        this is an optional pre-transformation of a baseline, not the SATrans hot path)."""
        dom = Xd[:, self.feature_index[self.domain_column][0]].long()
        theta = self.domain_map_dnn.linears[0](torch.relu(self.domain_embeddings.weight))       # (num_domains + 1, P)
        units, off, h = self.meta_dnn_hidden_units, 0, x
        for i in range(len(units) - 1):
            W = theta[:, off:off + units[i] * units[i + 1]].reshape(-1, units[i], units[i + 1])
            off += units[i] * units[i + 1]
            h = torch.bmm(h, W.index_select(0, dom))
            if i < len(units) - 2:
                h = torch.relu(h)
        h = Fnn.dropout(h, 0.1, self.training) + x
        if self.meta_net.use_norm:
            ln = self.meta_net.ffn_layer_norm
            h = Fnn.layer_norm(h, (h.shape[-1],), ln.weight, ln.bias, 1e-6)
        return h

    def forward(self, X):
        Xd, x = self._gather_fields(X, list(range(len(self._sparse_feats))))
        if 'usemetatrans' in self.flag:                               # reference autoint.py:86-89
            x = self.meta_transformation(Xd, x).contiguous()
        flat = x.reshape(x.shape[0], -1)
        att = x
        for layer in self.int_layers:
            att = layer(att)
        parts = []
        if self.att_layer_num > 0:
            parts.append(att.reshape(att.shape[0], -1))
        if len(self.dnn_hidden_units) > 0:
            dnn_in = flat if not self._dense_cols else torch.cat([flat, Xd[:, self._dense_cols].float()], dim=1)
            parts.append(self.dnn(dnn_in))
        stack = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()
        return self._head(stack, None, self.dnn_linear.weight, self.out.bias)       # y = sigmoid(dnn_linear(stack) + out.bias)
