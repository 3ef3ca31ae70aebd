"""Feature-column namedtuples + column-offset bookkeeping (API of deepctr-torch 0.2.9 `inputs`).

Only what /root/reference/models/{satrans,meta_basemodel,submodules}.py touch.
"""
from collections import OrderedDict, namedtuple

import torch
import torch.nn as nn

DEFAULT_GROUP_NAME = "default_group"


class SparseFeat(namedtuple('SparseFeat', ['name', 'vocabulary_size', 'embedding_dim', 'use_hash', 'dtype',
                                           'embedding_name', 'group_name'])):
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                group_name=DEFAULT_GROUP_NAME):
        if embedding_name is None:
            embedding_name = name
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        return super(SparseFeat, cls).__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype,
                                              embedding_name, group_name)

    def __hash__(self):
        return self.name.__hash__()


class VarLenSparseFeat(namedtuple('VarLenSparseFeat', ['sparsefeat', 'maxlen', 'combiner', 'length_name'])):
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super(VarLenSparseFeat, cls).__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda s: s.sparsefeat.name)
    vocabulary_size = property(lambda s: s.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda s: s.sparsefeat.embedding_dim)
    embedding_name = property(lambda s: s.sparsefeat.embedding_name)
    group_name = property(lambda s: s.sparsefeat.group_name)

    def __hash__(self):
        return self.name.__hash__()


class DenseFeat(namedtuple('DenseFeat', ['name', 'dimension', 'dtype'])):
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super(DenseFeat, cls).__new__(cls, name, dimension, dtype)

    def __hash__(self):
        return self.name.__hash__()


def build_input_features(feature_columns):
    features = OrderedDict()
    start = 0
    for feat in feature_columns:
        if feat.name in features:
            continue
        if isinstance(feat, SparseFeat):
            features[feat.name] = (start, start + 1)
            start += 1
        elif isinstance(feat, DenseFeat):
            features[feat.name] = (start, start + feat.dimension)
            start += feat.dimension
        elif isinstance(feat, VarLenSparseFeat):
            features[feat.name] = (start, start + feat.maxlen)
            start += feat.maxlen
            if feat.length_name is not None and feat.length_name not in features:
                features[feat.length_name] = (start, start + 1)
                start += 1
        else:
            raise TypeError("Invalid feature column type,got", type(feat))
    return features


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns).keys())


def varlen_embedding_lookup(X, embedding_dict, sequence_input_dict, varlen_sparse_feature_columns):
    if len(varlen_sparse_feature_columns) > 0:
        raise NotImplementedError("shim: no VarLenSparseFeat on the SATrans hot path")
    return {}


def get_varlen_pooling_list(embedding_dict, features, feature_index, varlen_sparse_feature_columns, device):
    if len(varlen_sparse_feature_columns) > 0:
        raise NotImplementedError("shim: no VarLenSparseFeat on the SATrans hot path")
    return []


def create_embedding_matrix(feature_columns, init_std=0.0001, linear=False, sparse=False, device='cpu'):
    cols = [f for f in feature_columns if isinstance(f, (SparseFeat, VarLenSparseFeat))]
    d = nn.ModuleDict({f.embedding_name: nn.Embedding(f.vocabulary_size, f.embedding_dim if not linear else 1,
                                                      sparse=sparse) for f in cols})
    for t in d.values():
        nn.init.normal_(t.weight, mean=0, std=init_std)
    return d.to(device)


def combined_dnn_input(sparse_embedding_list, dense_value_list):
    parts = []
    if len(sparse_embedding_list) > 0:
        parts.append(torch.flatten(torch.cat(sparse_embedding_list, dim=-1), start_dim=1))
    if len(dense_value_list) > 0:
        parts.append(torch.flatten(torch.cat(dense_value_list, dim=-1), start_dim=1))
    if not parts:
        raise NotImplementedError
    return torch.cat(parts, dim=-1) if len(parts) > 1 else parts[0]
