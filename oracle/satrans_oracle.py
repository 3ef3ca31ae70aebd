"""CPU oracle for the SATrans hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-on-CPU restatement (fp32 or fp64) of the reference algorithm for the path
`models/satrans.py` on `models/meta_basemodel.py` + `models/submodules.py`
(qwerfdsaplking/SATrans).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module; nothing under
`satrans_b200/` does.

Parity status: PINNED.  `oracle/make_golden.py` imports the unmodified reference modules
(through `oracle/ref_shims`) in the build container and dumps `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this restatement against those vectors
(y_pred, loss, reg-loss, every parameter gradient, parameters + Adam moments after 1 and 3
dense Adam steps).  The reference has no tests / golden vectors of its own.

The restatement keeps the reference's *op structure* (per-sample generated weights
materialised as `(B, 4096)`, per-sample `bmm`, dense L2 + dense Adam over whole tables), so that
timing it is a fair stand-in for the reference's own CPU path.

State layout: a dict of tensors keyed by the reference's `state_dict()` names
(SURVEY.md §8b).  Aliased keys (`K_meta_mlp`/`V_meta_mlp` without 'pos',
`domain_map_dnn_K/_V`) are resolved by `canonical_key`.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    """Shape/flag description of one SATrans instance (mirrors `SATrans.__init__` args,
    reference satrans.py:106-118 and main.py:292-306)."""
    sparse_names: List[str]                 # dnn_feature_columns order (SparseFeat)
    vocab_sizes: List[int]                  # vocabulary_size per sparse feature
    dense_names: List[str] = field(default_factory=list)   # DenseFeat(name, 1)
    domain_col: str = '301'
    num_domains: int = 3                    # num_domains_list[0]
    embedding_dim: int = 32
    layers: int = 3                         # domain_att_layer_num
    heads: int = 4                          # att_head_num
    meta_dnn_hidden_units: Tuple[int, ...] = (64, 32)
    meta_mode: str = 'QK'
    flag: str = 'sota'
    att_res: bool = True
    l2_reg_embedding: float = 1e-5
    init_std: float = 1e-4

    @property
    def units(self) -> List[int]:           # satrans.py:149  [d] + meta_dnn_hidden_units
        return [self.embedding_dim] + [int(u) for u in self.meta_dnn_hidden_units]

    @property
    def meta_param_size(self) -> int:       # satrans.py:156-162
        if 'bilinear' in self.flag:
            return self.embedding_dim ** 2 // self.heads
        if 'gate' in self.flag:
            return self.embedding_dim
        u = self.units
        return sum(u[i] * u[i + 1] for i in range(len(u) - 1))

    @property
    def metanet(self) -> bool:              # the MetaNet itself (satrans.py:65,72) and not its 'gate' / 'bilinear' ablations
        return 'gate' not in self.flag and 'bilinear' not in self.flag

    @property
    def onlyemb(self) -> bool:              # satrans.py:173-176: wide scenario embedding, no encoder
        return 'onlyemb' in self.flag

    @property
    def pos(self) -> bool:
        return 'pos' in self.flag

    @property
    def n_cols(self) -> int:
        return len(self.sparse_names) + len(self.dense_names)

    def col_of(self, name: str) -> int:     # build_input_features: sparse first (main.py:182-184), then dense
        names = self.sparse_names + self.dense_names
        return names.index(name)


def canonical_key(cfg: OracleConfig, key: str) -> str:
    """Map an aliased reference state-dict key to the tensor that owns it
    (satrans.py:44-47 K/V MetaNet aliases; satrans.py:179-180 encoder aliases)."""
    key = key.replace('domain_map_dnn_K.', 'domain_map_dnn_Q.').replace('domain_map_dnn_V.', 'domain_map_dnn_Q.')
    if not cfg.pos:
        key = key.replace('.K_meta_mlp.', '.Q_meta_mlp.').replace('.V_meta_mlp.', '.Q_meta_mlp.')
    return key


def trainable_keys(cfg: OracleConfig) -> List[str]:
    """Canonical keys of the parameters that receive a gradient in `SATrans.forward`
    (linear_model.* and out.bias exist in the state dict but are unused: satrans.py:197-256)."""
    keys = ['embedding_dict.%s.weight' % n for n in cfg.sparse_names]
    keys.append('domain_embeddings.weight')
    for i in range(cfg.layers):
        p = 'domain_int_layers.%d.' % i
        keys += [p + 'W_Query', p + 'W_Key', p + 'W_Value', p + 'Out_linear.weight',
                 p + 'layer_norm.weight', p + 'layer_norm.bias']
        if cfg.metanet and ('Q' in cfg.meta_mode or ('K' in cfg.meta_mode and not cfg.pos)):
            keys += [p + 'Q_meta_mlp.ffn_layer_norm.weight', p + 'Q_meta_mlp.ffn_layer_norm.bias']
        if cfg.metanet and cfg.pos and 'K' in cfg.meta_mode:
            keys += [p + 'K_meta_mlp.ffn_layer_norm.weight', p + 'K_meta_mlp.ffn_layer_norm.bias']
    if not cfg.onlyemb:
        keys += ['domain_map_dnn_Q.linears.0.weight', 'domain_map_dnn_Q.linears.0.bias']
    if cfg.pos and not cfg.onlyemb:         # ('onlyemb' slices the position part of the concatenation away: zero gradient)
        keys += ['layerid_embeddings.weight', 'qkvid_embeddings.weight']
    keys += ['dnn_linear.weight', 'dnn_linear.bias']
    return keys


def init_state(cfg: OracleConfig, seed: int = 1021, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Parameters drawn from the same *distributions* as the reference constructors (not the same
    RNG stream — SURVEY.md §7.3 #5): tables N(0,init_std) (meta_basemodel.py:118-119), layer weights
    and the layer's own LayerNorm N(0,0.05) (satrans.py:42-43), MetaNet LayerNorms (1,0)
    (satrans.py:44), domain/layer/qkv embeddings N(0,1) (satrans.py:145,169-170), encoder weight
    N(0,1e-4) + default-uniform bias (submodules.py:31-43), dnn_linear default nn.Linear init."""
    g = torch.Generator().manual_seed(int(seed))
    d = cfg.embedding_dim
    st: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def normal(shape, std):
        return (torch.randn(shape, generator=g, dtype=torch.float64) * std).to(dtype)

    def uniform(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    for n, v in zip(cfg.sparse_names, cfg.vocab_sizes):
        st['embedding_dict.%s.weight' % n] = normal((v, d), cfg.init_std)
    st['domain_embeddings.weight'] = normal((cfg.num_domains + 1, cfg.meta_param_size if cfg.onlyemb else d), 1.0)
    for i in range(cfg.layers):
        p = 'domain_int_layers.%d.' % i
        for w in ('W_Query', 'W_Key', 'W_Value', 'Out_linear.weight'):
            st[p + w] = normal((d, d), 0.05)
        st[p + 'layer_norm.weight'] = normal((d,), 0.05)
        st[p + 'layer_norm.bias'] = normal((d,), 0.05)
        for m in (('Q_meta_mlp', 'K_meta_mlp') if cfg.pos else ('Q_meta_mlp',)):
            st[p + m + '.ffn_layer_norm.weight'] = torch.ones(d, dtype=dtype)
            st[p + m + '.ffn_layer_norm.bias'] = torch.zeros(d, dtype=dtype)
    enc_in = 2 * d if cfg.pos else d
    if not cfg.onlyemb:
        st['domain_map_dnn_Q.linears.0.weight'] = normal((cfg.meta_param_size, enc_in), 1e-4)
        st['domain_map_dnn_Q.linears.0.bias'] = uniform((cfg.meta_param_size,), 1.0 / math.sqrt(enc_in))
    if cfg.pos:
        st['layerid_embeddings.weight'] = normal((cfg.layers, d), 1.0)
        st['qkvid_embeddings.weight'] = normal((3, d), 1.0)
    n_in = len(cfg.sparse_names) * d + len(cfg.dense_names)
    st['dnn_linear.weight'] = uniform((1, n_in), 1.0 / math.sqrt(n_in))
    st['dnn_linear.bias'] = uniform((1,), 1.0 / math.sqrt(n_in))
    return st


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def _layer_norm(x, w, b):                   # nn.LayerNorm(d, eps=1e-6): satrans.py:40, submodules.py:75
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-6)


def _drop(x, p, masks, name):
    """Inverted dropout: identity when masks is None (eval mode); `masks == 'rng'` draws from torch's RNG
    exactly like the reference's nn.Dropout in train mode; a dict injects keep-masks (1 = keep)."""
    if masks is None or p == 0.0:
        return x
    if isinstance(masks, str):
        return F.dropout(x, p, training=True)
    return x * masks[name].to(x.dtype) / (1.0 - p)


def meta_net(x, mlp_params, units, ln_w, ln_b, masks=None, mname=None, p_drop=0.1, relu_mask=None):
    """MetaNet.forward, submodules.py:77-103: per-sample 2-layer MLP whose weights are slices of
    `mlp_params (B, P)`, no biases, ReLU between, dropout, residual, LayerNorm.
    `relu_mask` (B, F, hidden) bool: test hook -- use these ReLU decisions instead of the sign of the pre-activation (a
    hidden unit within round-off of zero may be decided either way by two correct implementations; injecting the decisions
    of the implementation under test makes both differentiate the same piecewise-linear function)."""
    off = 0
    ws = []
    for i in range(len(units) - 1):         # submodules.py:81-86 (row-major reshape)
        n = units[i] * units[i + 1]
        ws.append(mlp_params[:, off:off + n].reshape(-1, units[i], units[i + 1]))
        off += n
    res = x
    for i, w in enumerate(ws):              # submodules.py:94-97
        x = x @ w
        if i < len(ws) - 1:
            x = F.relu(x) if relu_mask is None else x * relu_mask.to(x.dtype)
    x = _drop(x, p_drop, masks, mname)      # submodules.py:99
    x = x + res                             # submodules.py:100
    if ln_w is None:                        # use_norm=False (the baselines' meta module without 'metanorm', basemodel.py:146-147)
        return x
    return _layer_norm(x, ln_w, ln_b)       # submodules.py:101-102


def gather_fields(st, X, cfg: OracleConfig):
    """input_from_feature_columns, meta_basemodel.py:519-545 + cat at satrans.py:211.
    ids are fp32 values truncated toward zero to int64.  Returns (ids (B,F) int64, E0 (B,F,d))."""
    cols = [cfg.col_of(n) for n in cfg.sparse_names]
    ids = torch.stack([X[:, c].long() for c in cols], dim=1)
    embs = [st['embedding_dict.%s.weight' % n][ids[:, f]] for f, n in enumerate(cfg.sparse_names)]
    return ids, torch.stack(embs, dim=1)


def forward(st: Dict[str, torch.Tensor], X: torch.Tensor, cfg: OracleConfig,
            masks: Optional[Dict[str, torch.Tensor]] = None, return_aux: bool = False,
            relu_masks: Optional[Dict[str, torch.Tensor]] = None):
    """SATrans.forward (satrans.py:197-256) + Meta_Transformer_Layer.forward (satrans.py:50-100).

    `masks`: optional dict of dropout keep-masks keyed 'l{i}.q', 'l{i}.k' (B,F,d), 'l{i}.att'
    (H,B,F,F), 'l{i}.out' (B,F,d) — training-mode semantics with p=0.1; None = eval mode.
    """
    d, H, L = cfg.embedding_dim, cfg.heads, cfg.layers
    dh = d // H
    P = cfg.meta_param_size
    aux = {}
    ids, x = gather_fields(st, X, cfg)                                     # [K1]
    dom = X[:, cfg.col_of(cfg.domain_col)].long()                          # satrans.py:203
    e = F.relu(st['domain_embeddings.weight'][dom])                        # satrans.py:204,213
    if cfg.onlyemb:                                                        # satrans.py:173-176: identity "encoder"
        encode = lambda u: u
    else:
        We, be = st['domain_map_dnn_Q.linears.0.weight'], st['domain_map_dnn_Q.linears.0.bias']
        encode = lambda u: u @ We.t() + be
    if not cfg.pos:
        theta = encode(e)                                                  # satrans.py:218  [K2]
        thetas = [(theta, theta)] * L
    else:
        thetas = []
        for l in range(L):                                                 # satrans.py:225-234 [K2']
            per = []
            for q in range(2):                                             # q=2 (V) is generated but unused
                pos_emb = st['layerid_embeddings.weight'][l] + st['qkvid_embeddings.weight'][q]
                u = F.relu(torch.cat([e, pos_emb.expand(e.shape[0], -1)], dim=1))
                per.append(encode(u))
            thetas.append(tuple(per))
    aux['ids'], aux['E0'], aux['layer_in'], aux['att'] = ids, x, [], []
    for l in range(L):
        p = 'domain_int_layers.%d.' % l
        aux['layer_in'].append(x)
        inp = x
        q = inp @ st[p + 'W_Query']                                        # satrans.py:55-57 [K3]
        k = inp @ st[p + 'W_Key']
        v = inp @ st[p + 'W_Value']
        qn = p + 'Q_meta_mlp.ffn_layer_norm.'
        kn = p + ('K_meta_mlp' if cfg.pos else 'Q_meta_mlp') + '.ffn_layer_norm.'
        if 'Q' in cfg.meta_mode and 'gate' in cfg.flag:                    # satrans.py:61-62
            q = q * thetas[l][0].unsqueeze(1).expand(q.shape) * 2
        elif 'Q' in cfg.meta_mode and cfg.metanet:                         # satrans.py:60-66 [K4]
            q = meta_net(q, thetas[l][0][:, :P], cfg.units, st[qn + 'weight'], st[qn + 'bias'], masks, 'l%d.q' % l,
                         relu_mask=None if relu_masks is None else relu_masks['l%d.q' % l])
        if 'K' in cfg.meta_mode and 'gate' in cfg.flag:                    # satrans.py:68-69
            k = k * thetas[l][1].unsqueeze(1).expand(k.shape) * 2
        elif 'K' in cfg.meta_mode and cfg.metanet:                         # satrans.py:67-73 [K4]
            k = meta_net(k, thetas[l][1][:, :P], cfg.units, st[kn + 'weight'], st[kn + 'bias'], masks, 'l%d.k' % l,
                         relu_mask=None if relu_masks is None else relu_masks['l%d.k' % l])
        B_, F_ = q.shape[0], q.shape[1]
        qh = q.reshape(B_, F_, H, dh).permute(2, 0, 1, 3)                  # satrans.py:75-77 [K5]
        kh = k.reshape(B_, F_, H, dh).permute(2, 0, 1, 3)
        vh = v.reshape(B_, F_, H, dh).permute(2, 0, 1, 3)
        if 'bilinear' in cfg.flag:                                         # satrans.py:79-81 (bilinear_params = the Q side's theta)
            qh = qh @ thetas[l][0].reshape(-1, H, dh, dh).permute(1, 0, 2, 3)
        s = torch.matmul(qh, kh.transpose(-1, -2)) / (dh ** 0.5)           # satrans.py:84-86 [K6]
        a = F.softmax(s, dim=-1)
        a = _drop(a, 0.1, masks, 'l%d.att' % l)                            # satrans.py:87
        aux['att'].append(a)                                               # = layer.normalized_att_scores (H, B, F, F)
        r = torch.matmul(a, vh)                                            # satrans.py:88 [K7]
        r = r.permute(1, 2, 0, 3).reshape(B_, F_, d)                       # satrans.py:89-90
        o = r @ st[p + 'Out_linear.weight'].t()                            # satrans.py:91-94 [K8]
        if 'relu' in cfg.flag:
            o = F.relu(o)
        o = _drop(o, 0.1, masks, 'l%d.out' % l)
        if cfg.att_res:
            o = o + inp                                                    # satrans.py:96-97
        x = _layer_norm(o, st[p + 'layer_norm.weight'], st[p + 'layer_norm.bias'])   # satrans.py:99
    aux['EL'] = x
    out = torch.flatten(x, start_dim=1)                                    # satrans.py:244
    if cfg.dense_names:
        dense = torch.stack([X[:, cfg.col_of(n)] for n in cfg.dense_names], dim=1).to(out.dtype)
        out = torch.cat([out, dense], dim=1)                               # satrans.py:247-251
    logit = out @ st['dnn_linear.weight'].t() + st['dnn_linear.bias']      # satrans.py:254
    y = torch.sigmoid(logit)                                               # satrans.py:255
    aux['logit'] = logit
    return (y, aux) if return_aux else y


def self_attention_layer(x, p, heads: int, use_res: bool = True, masks=None):
    """SelfAttention_Layer.forward, submodules.py:208-238 ("next" row: AutoInt / STAR-Trans style front-end).
    p: dict with W_Query, W_Key, W_Value, W_Res, layer_norm.weight, layer_norm.bias (W_Out is never used)."""
    B_, F_, d = x.shape
    dh = d // heads
    q, k, v = x @ p['W_Query'], x @ p['W_Key'], x @ p['W_Value']
    qh = q.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    kh = k.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    vh = v.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    a = F.softmax(torch.matmul(qh, kh.transpose(-1, -2)) / (dh ** 0.5), dim=-1)     # :222-225
    a = _drop(a, 0.1, masks, 'att')
    r = torch.matmul(a, vh).permute(1, 2, 0, 3).reshape(B_, F_, d)                   # :227-230
    r = _drop(r, 0.1, masks, 'out')                                                 # :231
    if use_res:
        r = r + x @ p['W_Res']                                                      # :232-233
    return _layer_norm(F.relu(r), p['layer_norm.weight'], p['layer_norm.bias'])     # :234-235


def interacting_layer(x, p, heads: int, use_res: bool = True, scaling: bool = False):
    """deepctr-torch 0.2.9 `InteractingLayer.forward` (the AutoInt layer; call sites autoint.py:75-76,104-105).  PARITY UNPINNED:
    the source lives in the un-vendored third-party package (pin: readme.md:37), so this restates its published algorithm --
    Q,K,V = x W_{Query,key,Value}; per-head softmax(Q K^T [/ sqrt(dh) only if scaling]) V; + x W_Res; ReLU -- no dropout, no
    LayerNorm.  p: dict with W_Query, W_key, W_Value (and W_Res)."""
    B_, F_, d = x.shape
    dh = d // heads
    q, k, v = x @ p['W_Query'], x @ p['W_key'], x @ p['W_Value']
    qh = q.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    kh = k.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    vh = v.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    s = torch.matmul(qh, kh.transpose(-1, -2))
    if scaling:
        s = s / (dh ** 0.5)
    r = torch.matmul(F.softmax(s, dim=-1), vh).permute(1, 2, 0, 3).reshape(B_, F_, d)
    if use_res:
        r = r + x @ p['W_Res']
    return F.relu(r)


def mdr_interacting_layer(x, domain_ids, p, heads: int, num_domains: int, merge: str = 'no', use_res: bool = True, masks=None):
    """MDR_InteractingLayer.forward, star_trans.py:197-242: shared W_key / W_Value / W_Res, one W_Query and one LayerNorm per
    scenario (merge 'no' | 'sum' | 'dot' with the shared W_Query, :209-217); samples whose scenario id is outside 1..num_domains
    keep a zero result (:208,239).  p: dict keyed like the module's state dict."""
    B_, F_, d = x.shape
    dh = d // heads
    k, v = x @ p['W_key'], x @ p['W_Value']                                          # :204-205
    kh = k.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    vh = v.reshape(B_, F_, heads, dh).permute(2, 0, 1, 3)
    out = torch.zeros_like(x)
    for i in range(num_domains):
        wq = p['domain_W_Queries.%d' % i]
        if merge == 'sum':
            wq = wq + p['W_Query']
        elif merge == 'dot':
            wq = wq * p['W_Query']
        elif merge != 'no':
            raise NotImplementedError('not implemented parameters merge')
        m = domain_ids == (i + 1)                                                    # :218
        if not bool(m.any()):
            continue
        sx = x[m]
        qh = (sx @ wq).reshape(-1, F_, heads, dh).permute(2, 0, 1, 3)
        a = F.softmax(torch.matmul(qh, kh[:, m].transpose(-1, -2)) / (dh ** 0.5), dim=-1)      # :225-228
        a = _drop(a, 0.1, None if masks is None else {'att': masks['att'][:, m]}, 'att')
        r = torch.matmul(a, vh[:, m]).permute(1, 2, 0, 3).reshape(-1, F_, d)
        r = _drop(r, 0.1, None if masks is None else {'out': masks['out'][m]}, 'out')          # :233
        if use_res:
            r = r + sx @ p['W_Res']                                                  # :234-235
        r = _layer_norm(F.relu(r), p['domain_layer_norms.%d.weight' % i], p['domain_layer_norms.%d.bias' % i])   # :236-237
        out[m] = r
    return out


def star_trans_forward(st, X, sparse_names, domain_col, num_domains, heads, att_layer_num, domain_att_layer_num, merge='no',
                       use_res=True, dense_names=(), use_linear=False, use_domain_dnn_linear=False, domain_id_as_feature=False):
    """Star_Trans.forward (star_trans.py:744-775), att_layer_type='deepctr': gather of every field (without the scenario column
    unless domain_id_as_feature, :698-704), SelfAttention_Layer x att_layer_num, MDR_InteractingLayer x domain_att_layer_num,
    flatten (+ dense), dnn_linear -- or one head per scenario with use_domain_dnn_linear (:762-767; samples of no scenario keep
    a zero logit) -- plus the linear model's logit with use_linear (:749-750, basemodel.py:63-92; it sees EVERY column), sigmoid."""
    names = list(sparse_names) + list(dense_names)
    fields = [n for n in sparse_names if domain_id_as_feature or n != domain_col]    # filter_feature_columns, :702-704
    x = torch.stack([st['embedding_dict.%s.weight' % n][X[:, names.index(n)].long()] for n in fields], dim=1)
    dom = X[:, names.index(domain_col)].long()
    for i in range(att_layer_num):
        pre = 'int_layers.%d.' % i
        p = {k[len(pre):]: v for k, v in st.items() if k.startswith(pre)}
        p['W_Key'] = p['W_key']                                                      # star_trans.py's own class spells it W_key
        x = self_attention_layer(x, p, heads, use_res)
    for i in range(domain_att_layer_num):
        pre = 'domain_int_layers.%d.' % i
        p = {k[len(pre):]: v for k, v in st.items() if k.startswith(pre)}
        x = mdr_interacting_layer(x, dom, p, heads, num_domains, merge, use_res)
    out = torch.flatten(x, start_dim=1)
    if dense_names:
        out = torch.cat([out, torch.stack([X[:, names.index(n)] for n in dense_names], dim=1).to(out.dtype)], dim=1)
    logit = torch.zeros((X.shape[0], 1), dtype=out.dtype)
    if use_linear:
        for n in sparse_names:
            logit = logit + st['linear_model.embedding_dict.%s.weight' % n][X[:, names.index(n)].long()]
        if dense_names:
            logit = logit + torch.stack([X[:, names.index(n)] for n in dense_names], dim=1).to(out.dtype) @ st['linear_model.weight']
    if use_domain_dnn_linear:
        att = torch.zeros((X.shape[0], 1), dtype=out.dtype)
        for i in range(num_domains):
            m = dom == (i + 1)
            att = torch.where(m[:, None], out @ st['domain_dnn_linears.%d.weight' % i].t() + st['domain_dnn_linears.%d.bias' % i], att)
        logit = logit + att
    else:
        logit = logit + out @ st['dnn_linear.weight'].t() + st['dnn_linear.bias']
    return torch.sigmoid(logit)


def autoint_forward(st, X, sparse_names, heads, att_layer_num, use_res=True, dense_names=(), domain_col=None, flag='',
                    meta_units=(32, 64, 32)):
    """AutoInt.forward (autoint.py:82-123, flag without 'usemetatrans'): gather, InteractingLayer x att_layer_num, the deepctr DNN
    tower on the flattened embeddings (+ dense), dnn_linear (no bias) on [att_out | deep_out], PredictionLayer (bias + sigmoid).
    PARITY UNPINNED (InteractingLayer / DNN / PredictionLayer are deepctr-torch 0.2.9 code, not vendored)."""
    names = list(sparse_names) + list(dense_names)
    x = torch.stack([st['embedding_dict.%s.weight' % n][X[:, names.index(n)].long()] for n in sparse_names], dim=1)
    if 'usemetatrans' in flag:                                                  # autoint.py:86-89, basemodel.py:191-199
        e = F.relu(st['domain_embeddings.weight'][X[:, names.index(domain_col)].long()])
        theta = e @ st['domain_map_dnn.linears.0.weight'].t() + st['domain_map_dnn.linears.0.bias']
        norm = 'metanorm' in flag
        x = meta_net(x, theta, list(meta_units), st['meta_net.ffn_layer_norm.weight'] if norm else None,
                     st['meta_net.ffn_layer_norm.bias'] if norm else None)
    flat = torch.flatten(x, start_dim=1)
    if dense_names:
        flat = torch.cat([flat, torch.stack([X[:, names.index(n)] for n in dense_names], dim=1).to(flat.dtype)], dim=1)
    att = x
    for i in range(att_layer_num):
        pre = 'int_layers.%d.' % i
        att = interacting_layer(att, {k[len(pre):]: v for k, v in st.items() if k.startswith(pre)}, heads, use_res)
    deep = flat
    i = 0
    while 'dnn.linears.%d.weight' % i in st:                                   # DNN: Linear -> relu for every layer
        deep = F.relu(deep @ st['dnn.linears.%d.weight' % i].t() + st['dnn.linears.%d.bias' % i])
        i += 1
    stack = torch.cat([torch.flatten(att, start_dim=1), deep], dim=1)          # autoint.py:107-110
    return torch.sigmoid(stack @ st['dnn_linear.weight'].t() + st['out.bias'])


def reg_loss(st, cfg: OracleConfig):
    """get_regularization_loss, meta_basemodel.py:577-593: l2 * sum(square(w)) over the embedding
    tables only (linear_model is registered with l2=0, satrans.py:120)."""
    tot = None
    for n in cfg.sparse_names:
        t = torch.sum(cfg.l2_reg_embedding * torch.square(st['embedding_dict.%s.weight' % n]))
        tot = t if tot is None else tot + t
    return tot


def _loss(yp, y, kind='binary_crossentropy'):
    """loss strings of BaseModel._get_loss_func (meta_basemodel.py:627-640), reduction='sum' as fit uses them (:317)."""
    fn = {'binary_crossentropy': F.binary_cross_entropy, 'mse': F.mse_loss, 'mae': F.l1_loss}[kind]
    return fn(yp, y, reduction='sum')


def loss_and_grads(st, X, y, cfg: OracleConfig, masks=None, with_reg: bool = True, loss: str = 'binary_crossentropy',
                   relu_masks=None):
    """One forward + BCE(sum) + reg + backward (meta_basemodel.py:311-327) through torch autograd.
    Returns (y_pred (B,1), bce_sum, reg, {canonical key: grad})."""
    keys = trainable_keys(cfg)
    leaves = {k: st[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(st)
    work.update(leaves)
    yp = forward(work, X, cfg, masks, relu_masks=relu_masks)
    bce = _loss(yp.squeeze(-1), y.to(yp.dtype).reshape(-1), loss)                               # :317
    reg = reg_loss(work, cfg) if with_reg else torch.zeros((), dtype=yp.dtype)
    total = bce + reg                                                                           # :320
    total.backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in keys}
    return yp.detach(), bce.detach(), reg.detach(), grads


# --------------------------------------------------------------------------------------
# optimiser: torch.optim.Adam defaults (main.py:343) restated
# --------------------------------------------------------------------------------------
def adam_init(st, cfg: OracleConfig):
    return {'step': 0,
            'm': {k: torch.zeros_like(st[k]) for k in trainable_keys(cfg)},
            'v': {k: torch.zeros_like(st[k]) for k in trainable_keys(cfg)}}


def adam_apply(st, grads, opt, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False) single-tensor math:
    m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps;
    p.addcdiv_(m, denom, value=-(lr/bc1)).  In place on `st`/`opt`."""
    opt['step'] += 1
    t = opt['step']
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, g in grads.items():
        m, v, p = opt['m'][k], opt['v'][k], st[k]
        m.lerp_(g, 1.0 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(m, denom, value=-step_size)


def train_step(st, X, y, cfg: OracleConfig, opt, lr: float, masks=None):
    """forward + BCE(sum) + dense L2 + backward + dense Adam (meta_basemodel.py:311-328)."""
    yp, bce, reg, grads = loss_and_grads(st, X, y, cfg, masks)
    adam_apply(st, grads, opt, lr)
    return yp, bce, reg


class Trainer:
    """The reference's training step on CPU, shaped like meta_basemodel.py:311-328 + main.py:343: leaf parameters,
    train-mode forward (dropout from torch's RNG), BCE(sum) + dense L2, autograd backward (dense table
    gradients), torch.optim.Adam over every parameter.  Used as the timed CPU baseline."""

    def __init__(self, st, cfg: OracleConfig, lr: float, optimizer: str = 'adam', loss: str = 'binary_crossentropy',
                 **opt_kwargs):
        self.cfg = cfg
        self.loss = loss
        self.state = dict(st)
        self.params = {k: st[k].detach().clone().requires_grad_(True) for k in trainable_keys(cfg)}
        self.state.update(self.params)
        cls = {'adam': torch.optim.Adam, 'sgd': torch.optim.SGD, 'adagrad': torch.optim.Adagrad,
               'rmsprop': torch.optim.RMSprop}[optimizer]                     # meta_basemodel.py:612-625
        self.optim = cls(list(self.params.values()), lr=lr, **opt_kwargs)

    def step(self, X, y, train_mode=True, relu_masks=None):
        yp = forward(self.state, X, self.cfg, masks='rng' if train_mode else None, relu_masks=relu_masks).squeeze(-1)
        self.optim.zero_grad()
        loss = _loss(yp, y.reshape(-1), self.loss)
        reg = reg_loss(self.state, self.cfg)
        total = loss + reg
        loss_v, total_v = loss.item(), total.item()          # the reference's two .item() calls (:324-325)
        total.backward()
        self.optim.step()
        return loss_v, total_v - loss_v

    @torch.no_grad()
    def predict(self, X):
        return forward(self.state, X, self.cfg)


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------
def to_dtype(st, dtype):
    return OrderedDict((k, v.to(dtype)) for k, v in st.items())


def make_masks(cfg: OracleConfig, B: int, seed: int, p: float = 0.1):
    """Bernoulli keep-masks for the 4 dropout sites per layer (for mask-injected training parity)."""
    g = torch.Generator().manual_seed(seed)
    Fn, d, H = len(cfg.sparse_names), cfg.embedding_dim, cfg.heads
    m = {}
    for l in range(cfg.layers):
        m['l%d.q' % l] = (torch.rand((B, Fn, d), generator=g) >= p)
        m['l%d.k' % l] = (torch.rand((B, Fn, d), generator=g) >= p)
        m['l%d.att' % l] = (torch.rand((H, B, Fn, Fn), generator=g) >= p)
        m['l%d.out' % l] = (torch.rand((B, Fn, d), generator=g) >= p)
    return m
