"""Loading of tests/golden/*.npz (made by oracle/make_golden.py from the unmodified reference)."""
import json
import os

import numpy as np
import torch

from oracle import satrans_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['en_small', 'enp_small', 'q_small']
FORWARD_ONLY_CASES = ['relu_small_fwd']
# the 'gate' / 'bilinear' / 'onlyemb' ablation flags of the reference (satrans.py:59-81,158-176), with and without 'pos'
VARIANT_CASES = ['gate_small', 'gatepos_small', 'bilinear_small', 'bilinearpos_small', 'onlyemb_small', 'onlyembpos_small']
ENC_W = 'domain_map_dnn_Q.linears.0.weight'


def subsample(key):
    return slice(None, None, 8) if key == ENC_W else slice(None)


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.meta = json.loads(bytes(self.z['config']).decode())
        m = self.meta
        self.cfg = O.OracleConfig(sparse_names=m['sparse_names'], vocab_sizes=m['vocab_sizes'],
                                  dense_names=m['dense_names'], domain_col=m['domain_col'],
                                  num_domains=m['num_domains'], embedding_dim=32,
                                  layers=m['domain_att_layer_num'], heads=m['att_head_num'],
                                  meta_mode=m['meta_mode'], flag=m['flag'])
        self.B = m['B']

    def state(self, dtype=torch.float32):
        return {k[len('state/'):]: torch.from_numpy(self.z[k]).to(dtype) for k in self.z.files
                if k.startswith('state/')}

    def full_state_dict(self):
        """Reference-style state dict including the aliased keys."""
        st = self.state()
        L = self.cfg.layers
        full = dict(st)
        for a in ('K', 'V'):
            for s in ('weight', 'bias'):
                if 'domain_map_dnn_Q.linears.0.%s' % s in st:         # ('onlyemb' has no encoder)
                    full['domain_map_dnn_%s.linears.0.%s' % (a, s)] = st['domain_map_dnn_Q.linears.0.%s' % s]
        for i in range(L):
            for a in ('K', 'V'):
                for s in ('weight', 'bias'):
                    k = 'domain_int_layers.%d.%s_meta_mlp.ffn_layer_norm.%s' % (i, a, s)
                    if k not in full:
                        full[k] = st['domain_int_layers.%d.Q_meta_mlp.ffn_layer_norm.%s' % (i, s)]
        return full

    def X(self, k=0):
        return torch.from_numpy(self.z['X'][k * self.B:(k + 1) * self.B])

    def y(self, k=0):
        return torch.from_numpy(self.z['y'][k * self.B:(k + 1) * self.B])

    def arr(self, key):
        return torch.from_numpy(self.z[key])

    def x_dict(self, k=0):
        names = self.cfg.sparse_names + self.cfg.dense_names
        X = self.z['X'][k * self.B:(k + 1) * self.B]
        return {n: (X[:, i].astype(np.int64) if n in self.cfg.sparse_names else X[:, i]) for i, n in enumerate(names)}
