"""Helpers for the GPU parity tests: build the drop-in model from a golden fixture, map gradients back
to reference state-dict keys."""
import numpy as np
import torch

from golden_util import Golden
from oracle import satrans_oracle as O
from satrans_b200 import DenseFeat, SATrans, SparseFeat


def model_from_golden(g: Golden, device='cuda:0', precision='fp32'):
    m = g.meta
    cols = ([SparseFeat(n, v, 32) for n, v in zip(m['sparse_names'], m['vocab_sizes'])] +
            [DenseFeat(n, 1) for n in m['dense_names']])
    model = SATrans(cols, cols, [m['domain_col']], [m['num_domains']], att_layer_num=0,
                    domain_att_layer_num=m['domain_att_layer_num'], att_head_num=m['att_head_num'],
                    meta_mode=m['meta_mode'], seed='1021', device=device, flag=m['flag'])
    missing = model.load_state_dict(g.full_state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.precision = precision        # CUDA-core kernels unless a test selects the tensor-core path (the package default)
    return model


def model_from_cfg(cfg: O.OracleConfig, state, device='cuda:0', precision='fp32'):
    cols = ([SparseFeat(n, v, cfg.embedding_dim) for n, v in zip(cfg.sparse_names, cfg.vocab_sizes)] +
            [DenseFeat(n, 1) for n in cfg.dense_names])
    model = SATrans(cols, cols, [cfg.domain_col], [cfg.num_domains], att_layer_num=0,
                    domain_att_layer_num=cfg.layers, att_head_num=cfg.heads, meta_mode=cfg.meta_mode,
                    seed=1021, device=device, flag=cfg.flag)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in state.items():
            sd[k].copy_(v)
    model.precision = precision
    return model


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
