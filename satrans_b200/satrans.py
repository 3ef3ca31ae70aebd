"""SATrans — drop-in for `models/satrans.py` of the reference (Scenario-Adaptive Feature Interaction, KDD'23).

Same constructor signature (reference satrans.py:106-118), same `state_dict()` key layout (SURVEY.md §8b:
a reference checkpoint loads 1:1 and vice versa), same `forward(X) -> (B, 1)` contract.  The modules below are
parameter containers only: every dense parameter is a view into ONE flat fp32 buffer (layout in
include/satrans_b200.h) and the arithmetic runs in the fused sm_100a kernels through `Engine`.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from .basemodel import BaseModel
from .inputs import DenseFeat, SparseFeat


class MetaNet(nn.Module):
    """Container for the MetaNet's LayerNorm (reference submodules.py:64-76); the generated weights come from
    the scenario encoder, the math is in the kernels."""

    def __init__(self, hidden_dim, dropout=0.1, use_norm=True, meta_dnn_hidden_units=(32, 64, 32), flag=None):
        super().__init__()
        self.use_norm = use_norm
        self.meta_dnn_hidden_units = meta_dnn_hidden_units
        self.flag = flag
        if use_norm:
            self.ffn_layer_norm = nn.LayerNorm(hidden_dim, eps=1e-6)


class DNN_v2(nn.Module):
    """Scenario encoder container: `linears.0` = Linear(in, 4096) (reference submodules.py:18-45, with the
    single hidden layer SATrans uses; weight N(0, init_std), default-uniform bias)."""

    def __init__(self, inputs_dim, hidden_units, init_std=0.0001):
        super().__init__()
        if len(hidden_units) != 1:
            raise NotImplementedError("SATrans uses a single-layer scenario encoder (reference satrans.py:164,178)")
        self.linears = nn.ModuleList([nn.Linear(inputs_dim, hidden_units[0])])
        nn.init.normal_(self.linears[0].weight, mean=0, std=init_std)


class Meta_Transformer_Layer(nn.Module):
    """Parameter container of one scenario-adaptive interacting layer (reference satrans.py:13-48)."""

    def __init__(self, embedding_size, meta_dnn_hidden_units, flag, mode='Q', head_num=2, use_res=True, scaling=True,
                 seed=1024, device='cpu'):
        super().__init__()
        if head_num <= 0:
            raise ValueError('head_num must be a int > 0')
        if embedding_size % head_num != 0:
            raise ValueError('embedding_size is not an integer multiple of head_num!')
        self.att_embedding_size = embedding_size // head_num
        self.head_num, self.use_res, self.scaling, self.seed, self.flag, self.mode = head_num, use_res, scaling, seed, flag, mode
        d = embedding_size
        self.W_Query = nn.Parameter(torch.empty(d, d))
        self.W_Key = nn.Parameter(torch.empty(d, d))
        self.W_Value = nn.Parameter(torch.empty(d, d))
        self.Out_linear = nn.Linear(d, d, bias=False)
        self.layer_norm = nn.LayerNorm(d, eps=1e-6)
        # reference satrans.py:42-43: N(0, 0.05) on everything created so far — including the layer's own
        # LayerNorm weight/bias — and *before* the MetaNets exist (their LayerNorms stay at (1, 0)).
        for t in self.parameters():
            nn.init.normal_(t, mean=0.0, std=0.05)
        self.Q_meta_mlp = MetaNet(d, 0.1, True, meta_dnn_hidden_units)
        self.K_meta_mlp = self.Q_meta_mlp if 'pos' not in flag else MetaNet(d, 0.1, True, meta_dnn_hidden_units)
        self.V_meta_mlp = self.Q_meta_mlp if 'pos' not in flag else MetaNet(d, 0.1, True, meta_dnn_hidden_units)
        self.normalized_att_scores = None      # (H,B,F,F) side channel of the reference; never materialised here
        self.to(device)


class SATrans(BaseModel):
    def __init__(self, linear_feature_columns, dnn_feature_columns, domain_column_list,
                 num_domains_list, att_layer_num=2, domain_att_layer_num=1,
                 att_head_num=2,
                 share_domain_dnn_across_layers=False,
                 use_domain_dnn_linear=False,
                 att_res=True,
                 use_linear=True,
                 use_dnn=False,
                 dnn_hidden_units=(256, 128), dnn_activation='relu',
                 meta_dnn_hidden_units=(64, 32),
                 meta_mode='Q',
                 l2_reg_dnn=0, l2_reg_embedding=1e-5, dnn_use_bn=False, dnn_dropout=0, init_std=0.0001, seed=1024,
                 task='binary', device='cpu', gpus=None, flag=None, table_shard=None):
        super().__init__(linear_feature_columns, dnn_feature_columns, l2_reg_linear=0,
                         l2_reg_embedding=l2_reg_embedding, init_std=init_std, seed=seed, task=task,
                         device=device, gpus=gpus, table_shard=table_shard)
        if not isinstance(flag, str):
            raise TypeError("flag must be a str (the reference substring-tests it: satrans.py:46,158-173)")
        # ablation variants of the scenario-specific projection (reference satrans.py:59-81,158-176): 'gate' / 'bilinear' replace
        # the MetaNet by a generated gate / per-head bilinear map, 'onlyemb' replaces the scenario encoder by a wide embedding
        gate, bilinear, onlyemb = 'gate' in flag, 'bilinear' in flag, 'onlyemb' in flag
        if gate and bilinear:
            raise ValueError("'gate' and 'bilinear' cannot be combined (the reference's own forward fails on the shapes: "
                             "meta_param_size is the bilinear one, satrans.py:158-162, and the gate needs embedding_size)")
        if gate and not ('Q' in meta_mode or 'K' in meta_mode):
            raise ValueError("'gate' needs 'Q' and / or 'K' in meta_mode")
        if onlyemb and (gate or bilinear):
            raise NotImplementedError("'onlyemb' combined with 'gate' / 'bilinear' is not built")
        if len(domain_column_list) != 1:
            raise NotImplementedError("multi-column scenarios (reference satrans.py:186-188,205-207) are not built")
        d = self.embedding_size
        if d != _lib.SAT_D:
            raise NotImplementedError("kernels are built for embedding_dim=32 (reference readme.md:17-24), got %d" % d)
        units = [d] + [int(u) for u in meta_dnn_hidden_units]
        if units != [32, 64, 32]:
            raise NotImplementedError("kernels are built for meta_dnn_hidden_units=(64, 32), got %s"
                                      % (tuple(meta_dnn_hidden_units),))
        if att_head_num not in (1, 2, 4, 8):
            if att_head_num <= 0:
                raise ValueError('head_num must be a int > 0')
            if d % att_head_num != 0:
                raise ValueError('embedding_size is not an integer multiple of head_num!')
            raise NotImplementedError("att_head_num must be one of 1, 2, 4, 8")
        self.use_linear, self.use_dnn = use_linear, use_dnn
        self.num_domains_list = num_domains_list
        self.use_domain_dnn_linear = use_domain_dnn_linear
        self.share_domain_dnn_across_layers = share_domain_dnn_across_layers
        self.linear_feature_columns, self.dnn_feature_columns = linear_feature_columns, dnn_feature_columns
        self.dnn_hidden_units = dnn_hidden_units
        self.domain_att_layer_num, self.att_layer_num = domain_att_layer_num, att_layer_num
        self.domain_column_list = domain_column_list
        self.flag, self.meta_mode, self.att_res, self.att_head_num = flag, meta_mode, att_res, att_head_num
        self.meta_dnn_hidden_units = units
        self.dropout_rate = 0.1             # attndrop_rate = drop_rate = 0.1 (reference satrans.py:27-28)
        self.dense_table_updates = True     # reference semantics: L2 + Adam on every table row every step
        self.device_metrics = True          # fit(verbose>0): per-batch logloss/auc/mse/accuracy on the GPU (False: sklearn on the host)
        # Debug switch: the reference stores softmax(QK^T/sqrt(dh)) of every layer in
        # `domain_int_layers[i].normalized_att_scores` (H, B, F, F) as a side effect of forward (satrans.py:87), read only by its
        # 'showattn' / 'instattn' analysis dumps.  The fused kernels never materialise it; with this switch on, forward() saves
        # the layer activations and rebuilds the scores from them (plain torch ops, analysis path, not the hot path).
        self.debug_att_scores = False

        field_num = len(self.embedding_dict)
        if field_num > _lib.SAT_MAXF:
            raise NotImplementedError("at most %d sparse fields" % _lib.SAT_MAXF)
        dense_cols = [f for f in linear_feature_columns if isinstance(f, DenseFeat)]
        n_dense = self.compute_input_dim(dense_cols)
        self.domain_embedding_dim = d
        self.domain_embeddings = nn.Embedding(num_domains_list[0] + 1, d)
        self.domain_int_layers = nn.ModuleList(
            [Meta_Transformer_Layer(d, units, flag, meta_mode, att_head_num, att_res, device='cpu')
             for _ in range(domain_att_layer_num)])
        meta_param_size = sum(units[i] * units[i + 1] for i in range(len(units) - 1))
        if bilinear:                                   # reference satrans.py:158-162
            meta_param_size = (d ** 2) // att_head_num
        elif gate:
            meta_param_size = d
        if 'pos' in flag:
            self.domain_embedding_dim *= 2
            self.layerid_embeddings = nn.Embedding(domain_att_layer_num, d)
            self.qkvid_embeddings = nn.Embedding(3, d)
        if onlyemb:                                    # reference satrans.py:173-176: a wide embedding, identity "encoder"
            self.domain_embeddings = nn.Embedding(num_domains_list[0] + 1, meta_param_size)
            self.domain_map_dnn_Q = lambda x: x
        else:
            self.domain_map_dnn_Q = DNN_v2(self.domain_embedding_dim, [meta_param_size])
            self.domain_map_dnn_K = self.domain_map_dnn_Q
            self.domain_map_dnn_V = self.domain_map_dnn_Q
        self.dnn_linear = nn.Linear(field_num * d + n_dense, 1)

        # ---- flat dense-parameter buffer ------------------------------------------------------------------
        L, pos = domain_att_layer_num, 'pos' in flag
        n_dom = num_domains_list[0] + 1
        lay = dict(F=field_num, L=L, H=att_head_num, n_dom=n_dom, n_dense=n_dense, n_cols=self._n_cols(),
                   enc_in=self.domain_embedding_dim)
        meta = not (gate or bilinear)                  # the MetaNet itself (with its LayerNorm and dropout) on that side
        lay['flags'] = ((_lib.FLAG_POS if pos else 0) | (_lib.FLAG_META_Q if meta and 'Q' in meta_mode else 0) |
                        (_lib.FLAG_META_K if meta and 'K' in meta_mode else 0) | (_lib.FLAG_RELU if 'relu' in flag else 0) |
                        (_lib.FLAG_RES if att_res else 0) | (_lib.FLAG_GATE_Q if gate and 'Q' in meta_mode else 0) |
                        (_lib.FLAG_GATE_K if gate and 'K' in meta_mode else 0) | (_lib.FLAG_BILINEAR if bilinear else 0) |
                        (_lib.FLAG_ONLYEMB if onlyemb else 0))
        off = 0
        lay['off_dom'] = off; off += n_dom * (meta_param_size if onlyemb else d)
        lay['off_lid'] = off; off += (L * d if pos else 0)
        lay['off_qkv'] = off; off += (3 * d if pos else 0)
        lay['off_encw'] = off; off += (0 if onlyemb else meta_param_size * lay['enc_in'])
        lay['off_encb'] = off; off += (0 if onlyemb else meta_param_size)
        lay['off_layers'] = off; off += L * _lib.SAT_LAYER_FLOATS
        lay['off_headw'] = off; off += field_num * d + n_dense
        lay['off_headb'] = off; off += 1
        lay['n_dense_params'] = (off + 3) // 4 * 4
        lay['total_rows'] = int(self._row_off[-1])
        lay['row_off'] = [int(v) for v in self._row_off]
        lay['sparse_cols'] = [self.feature_index[f.name][0] for f in self._sparse_feats]
        lay['dense_cols'] = []
        for f in dense_cols:
            a, b = self.feature_index[f.name]
            lay['dense_cols'] += list(range(a, b))
        lay['dom_col'] = self.feature_index[domain_column_list[0]][0]
        self._layout = lay
        self._dense = torch.zeros(lay['n_dense_params'], dtype=torch.float32, device=device)
        self._flat_ptr = None
        self.to(device)
        self._ensure_flat()

    # ---- flat-buffer binding --------------------------------------------------------------------------------
    def _flat_slots(self):
        """[(parameter, float offset)] for every parameter the kernels read."""
        lay, d = self._layout, _lib.SAT_D
        slots = [(self.domain_embeddings.weight, lay['off_dom'])]
        if 'pos' in self.flag:
            slots += [(self.layerid_embeddings.weight, lay['off_lid']), (self.qkvid_embeddings.weight, lay['off_qkv'])]
        if 'onlyemb' not in self.flag:
            enc = self.domain_map_dnn_Q.linears[0]
            slots += [(enc.weight, lay['off_encw']), (enc.bias, lay['off_encb'])]
        for i, layer in enumerate(self.domain_int_layers):
            o = lay['off_layers'] + i * _lib.SAT_LAYER_FLOATS
            slots += [(layer.W_Query, o), (layer.W_Key, o + 1024), (layer.W_Value, o + 2048),
                      (layer.Out_linear.weight, o + 3072), (layer.layer_norm.weight, o + 4096),
                      (layer.layer_norm.bias, o + 4128), (layer.Q_meta_mlp.ffn_layer_norm.weight, o + 4160),
                      (layer.Q_meta_mlp.ffn_layer_norm.bias, o + 4192)]
            if 'pos' in self.flag:
                slots += [(layer.K_meta_mlp.ffn_layer_norm.weight, o + 4224),
                          (layer.K_meta_mlp.ffn_layer_norm.bias, o + 4256)]
        slots += [(self.dnn_linear.weight, lay['off_headw']), (self.dnn_linear.bias, lay['off_headb'])]
        return slots

    def _ensure_flat(self):
        """Make every kernel-visible parameter a view of `_dense` / `_tables` on the parameters' current device
        (after construction, `.to()`, `.cuda()` ...).  Values are preserved."""
        dev = self.dnn_linear.weight.device
        slots = self._flat_slots()
        sharded = self._shard[1] > 1
        ok = (self._dense.device == dev and self._tables.device == dev and
              all(p.data_ptr() == self._dense.data_ptr() + 4 * off for p, off in slots) and
              (sharded or
               all(self.embedding_dict[f.embedding_name].weight.data_ptr() == self._tables.data_ptr() + 4 * _lib.SAT_D * int(r)
                   for f, r in zip(self._sparse_feats, self._row_off[:-1]))))
        if ok:
            return
        flat = torch.zeros(self._layout['n_dense_params'], dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, off in slots:
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data.to(dev, torch.float32))
                p.data = view
            first = self.embedding_dict[self._sparse_feats[0].embedding_name].weight
            if sharded:
                self._tables = self._tables.to(dev)
            elif first.data_ptr() != self._tables.data_ptr() or self._tables.device != dev:
                tables = torch.empty((int(self._row_off[-1]), _lib.SAT_D), dtype=torch.float32, device=dev)
                for f, a, b in zip(self._sparse_feats, self._row_off[:-1], self._row_off[1:]):
                    w = self.embedding_dict[f.embedding_name].weight
                    tables[int(a):int(b)].copy_(w.data.to(dev, torch.float32))
                    w.data = tables[int(a):int(b)]
                self._tables = tables
        self._dense = flat

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        if hasattr(self, '_layout'):
            self._ensure_flat()
            self.device = str(self.dnn_linear.weight.device)
        return out

    def get_feature_columns(self, feature_columns, col_names):
        return [feat for feat in feature_columns if feat.name in col_names]

    # ---- forward ----------------------------------------------------------------------------------------
    def forward(self, X):
        """X: (B, n_cols) matrix with columns in `feature_index` order (float32 as in the reference, or
        int32/int64 ids when there are no dense features) -> y_pred (B, 1) float32.  Inference-only entry
        point (no autograd graph): training goes through `fit` / `Engine.train_step`."""
        Xd = self._as_device_matrix(X)
        if self.debug_att_scores:
            return self._forward_with_att_scores(Xd)
        y = self.engine.forward(Xd, training=self.training, dropout_p=self.dropout_rate,
                                seed=int(torch.initial_seed()) + self.engine.launches)
        return y.unsqueeze(1)

    def _forward_with_att_scores(self, Xd):
        """forward() that also fills `domain_int_layers[i].normalized_att_scores` (H, B, F, F), eval-mode semantics (dropout off):
        the kernels save the pre-LayerNorm q / k of every layer (training workspace), the scores are rebuilt from them."""
        import torch.nn.functional as Fn
        eng = self.engine
        y = eng.forward(Xd, training=True, dropout_p=0.0, seed=0)
        w = eng.workspace(Xd.shape[0], True)
        lay = self._layout
        B, F, L, H, d = Xd.shape[0], lay['F'], lay['L'], lay['H'], _lib.SAT_D
        cap_tok = w.cap_B * F
        A = _lib.SAT_ACT_PER_LAYER
        acts = w['acts'][:(L * A + 1) * cap_tok * d].view(L * A + 1, cap_tok, d)
        perm = w['perm'][:B].long()
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(B, device=perm.device)
        for l, layer in enumerate(self.domain_int_layers):
            zq = acts[l * A + 1, :B * F].view(B, F, d)[inv]              # ACT_ZQ / ACT_ZK, sorted position -> sample order
            zk = acts[l * A + 2, :B * F].view(B, F, d)[inv]
            qn, kn = layer.Q_meta_mlp.ffn_layer_norm, layer.K_meta_mlp.ffn_layer_norm
            q = Fn.layer_norm(zq, (d,), qn.weight, qn.bias, 1e-6) if 'Q' in self.meta_mode else zq
            k = Fn.layer_norm(zk, (d,), kn.weight, kn.bias, 1e-6) if 'K' in self.meta_mode else zk
            qh = q.view(B, F, H, d // H).permute(2, 0, 1, 3)
            kh = k.view(B, F, H, d // H).permute(2, 0, 1, 3)
            layer.normalized_att_scores = torch.softmax(torch.matmul(qh, kh.transpose(-1, -2)) / math.sqrt(d // H), dim=-1).detach()
        return y.unsqueeze(1)
