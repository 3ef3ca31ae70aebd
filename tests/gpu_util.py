"""Helpers for the GPU parity tests: build the drop-in model from a golden fixture, map gradients back
to reference state-dict keys."""
import numpy as np
import torch

from golden_util import Golden
from oracle import satrans_oracle as O
from satrans_b200 import DenseFeat, SATrans, SparseFeat, _lib


def model_from_golden(g: Golden, device='cuda:0', precision=None):
    m = g.meta
    cols = ([SparseFeat(n, v, 32) for n, v in zip(m['sparse_names'], m['vocab_sizes'])] +
            [DenseFeat(n, 1) for n in m['dense_names']])
    model = SATrans(cols, cols, [m['domain_col']], [m['num_domains']], att_layer_num=0,
                    domain_att_layer_num=m['domain_att_layer_num'], att_head_num=m['att_head_num'],
                    meta_mode=m['meta_mode'], seed='1021', device=device, flag=m['flag'])
    missing = model.load_state_dict(g.full_state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    if precision is not None:          # None: the package default (tcgen05 3xTF32, what bench.py times)
        model.precision = precision
    return model


def model_from_cfg(cfg: O.OracleConfig, state, device='cuda:0', precision=None):
    cols = ([SparseFeat(n, v, cfg.embedding_dim) for n, v in zip(cfg.sparse_names, cfg.vocab_sizes)] +
            [DenseFeat(n, 1) for n in cfg.dense_names])
    model = SATrans(cols, cols, [cfg.domain_col], [cfg.num_domains], att_layer_num=0,
                    domain_att_layer_num=cfg.layers, att_head_num=cfg.heads, meta_mode=cfg.meta_mode,
                    seed=1021, device=device, flag=cfg.flag)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in state.items():
            sd[k].copy_(v)
    if precision is not None:
        model.precision = precision
    return model


def relu_masks(model, w, B):
    """The MetaNet ReLU decisions the forward kernel took, as {'l{i}.q' | 'l{i}.k': bool (B, F, 64)} in sample order
    (saved behind the activations: uint32 [L][2 sides][cap_B*F][2], indexed by sorted position * F + field)."""
    lay = model._layout
    L, F = lay['L'], lay['F']
    cap_tok = w.cap_B * F
    A = _lib.SAT_ACT_PER_LAYER
    raw = w['acts'].view(torch.int32)[(L * A + 1) * cap_tok * 32:(L * A + 1) * cap_tok * 32 + L * 2 * cap_tok * 2]
    raw = raw.view(L, 2, cap_tok, 2)[:, :, :B * F].reshape(L, 2, B, F, 2).cpu()
    perm = w['perm'][:B].long().cpu()
    bits = torch.arange(32, dtype=torch.int64)
    out = {}
    for l in range(L):
        for side, name in ((0, 'q'), (1, 'k')):
            words = raw[l, side].long() & 0xFFFFFFFF                       # (B sorted, F, 2)
            m = ((words.unsqueeze(-1) >> bits) & 1).bool().reshape(B, F, 64)
            full = torch.zeros_like(m)
            full[perm] = m                                                 # sorted position p holds sample perm[p]
            out['l%d.%s' % (l, name)] = full
    return out


def compact_vocab(cfg: O.OracleConfig, state, X):
    """(cfg', state', X') with every table cut down to the rows this batch touches (ids renumbered): the same function of the
    same rows, small enough for an fp64 autograd pass at full-vocabulary configurations.  Also returns per field the original
    ids of the kept rows."""
    import copy
    X = X.clone()
    st = dict(state)
    kept = {}
    sizes = []
    for f, n in enumerate(cfg.sparse_names):
        c = cfg.col_of(n)
        ids = X[:, c].long()
        uniq, inv = torch.unique(ids, return_inverse=True)
        if n == cfg.domain_col:                                            # the scenario id also indexes domain_embeddings: keep as is
            uniq = torch.arange(cfg.vocab_sizes[f])
            inv = ids
        st['embedding_dict.%s.weight' % n] = state['embedding_dict.%s.weight' % n][uniq]
        X[:, c] = inv.to(X.dtype)
        kept[n] = uniq
        sizes.append(int(uniq.numel()))
    cfg2 = copy.copy(cfg)
    cfg2.vocab_sizes = sizes
    return cfg2, st, X, kept


def dense_grads(model, w):
    """{canonical reference key: gradient tensor (cpu)} for the dense parameters, from the flat dgrad buffer."""
    names = {id(p): n for n, p in model.named_parameters()}
    flat = w['dgrad'].detach().cpu()
    out = {}
    for p, off in model._flat_slots():
        out[names[id(p)]] = flat[off:off + p.numel()].view(p.shape).clone()
    return out


def table_grads(model, w, l2):
    """{embedding_dict.<f>.weight: dense gradient incl. the 2*l2*w regulariser term (cpu)}."""
    n = int(w['scalars'][2].item())
    rows = w['uniq_rows'][:n].long().cpu()
    rg = w['row_grad'][:n * 32].view(n, 32).cpu()
    tables = model._tables.detach().cpu()
    dense = 2.0 * l2 * tables
    dense[rows] += rg
    out = {}
    for f, a, b in zip(model._sparse_feats, model._row_off[:-1], model._row_off[1:]):
        out['embedding_dict.%s.weight' % f.embedding_name] = dense[int(a):int(b)]
    return out, rows


def field_attention_dropout_masks(B, seed, p_drop=0.1, F=8, H=4):
    """The keep-masks the field-attention kernels draw for (seed, B, F, H) in train mode, read out of the FORWARD kernel alone
    with inputs that make every output element a product of mask bits (the masks are counter-based: they depend on the seed,
    the token index, the dropout site and the element index, never on the data):
        pass 0: x = ones, W_Q = W_K = 0, W_V = I, no residual / LayerNorm -> out[b,i,c] = keep_out[b,i,c] * (mean of keep_att) / .81
        pass s: x[b,j] = one-hot at channel (j + s) % dh of every head       -> out[b,i,h*dh+c] = keep_att[h,b,i,j] * keep_out / (F * .81)
    A few shifts s resolve every attention bit (an element hidden by a dropped output channel shows up under another shift).
    Returns {'att': (H, B, F, F) , 'out': (B, F, 32)} float 0/1 masks in the layout of oracle._drop."""
    dh = 32 // H
    assert F <= dh
    lib = _lib.lib()
    dev = 'cuda:0'
    eye = torch.eye(32, device=dev).contiguous()
    zero = torch.zeros((32, 32), device=dev)

    def run(x):
        out = torch.empty_like(x)
        _lib.check(lib.sat_field_attention_fwd(x.data_ptr(), out.data_ptr(), zero.data_ptr(), 1, zero.data_ptr(), eye.data_ptr(), None,
                                               None, None, None, B, F, H, 0, 1, 1, float(p_drop), int(seed),
                                               torch.cuda.current_stream().cuda_stream), 'sat_field_attention_fwd')
        return out.cpu()
    keep_out = (run(torch.ones((B, F, 32), device=dev)) != 0).float()
    att = torch.full((H, B, F, F), -1.0)
    for shift in range(dh):
        x = torch.zeros((B, F, 32), device=dev)
        for j in range(F):
            for h in range(H):
                x[:, j, h * dh + (j + shift) % dh] = 1.0
        o = run(x)
        for j in range(F):
            for h in range(H):
                c = h * dh + (j + shift) % dh
                seen = keep_out[:, :, c] > 0                                   # (B, F): query tokens whose channel c survived
                bit = (o[:, :, c] != 0).float()
                att[h][:, :, j] = torch.where(seen & (att[h][:, :, j] < 0), bit, att[h][:, :, j])
        if (att >= 0).all():
            break
    assert (att >= 0).all()
    frac = 1.0 - att.mean().item()
    assert abs(frac - p_drop) < 0.03 and abs(1.0 - keep_out.mean().item() - p_drop) < 0.03, (frac, keep_out.mean().item())
    return {'att': att, 'out': keep_out}
