"""Feature columns and input-column bookkeeping for the SATrans drop-in API.

The reference takes these names from deepctr-torch 0.2.9 (`from deepctr_torch.inputs import
SparseFeat, DenseFeat, get_feature_names`, reference main.py:17,182-191;
`build_input_features`, reference meta_basemodel.py:26,164).  deepctr-torch is a third-party
dependency that is not part of the reference tree, so the drop-in ships its own
implementation with the same constructor signatures, field order and column-offset rule.
Only fixed-length sparse and dense features exist on the SATrans path (main.py:103,146:
`var_features = []`); VarLenSparseFeat is accepted as a type so that user code that merely
imports it keeps working, but a model built with one raises.
"""
from collections import OrderedDict, namedtuple

DEFAULT_GROUP_NAME = "default_group"


class SparseFeat(namedtuple('SparseFeat', ['name', 'vocabulary_size', 'embedding_dim', 'use_hash', 'dtype',
                                           'embedding_name', 'group_name'])):
    """One categorical input column → one embedding table of `vocabulary_size` rows."""
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32",
                embedding_name=None, group_name=DEFAULT_GROUP_NAME):
        if embedding_name is None:
            embedding_name = name
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        if use_hash:
            raise NotImplementedError("feature hashing is not supported (it is a no-op in deepctr-torch too)")
        return super().__new__(cls, name, int(vocabulary_size), int(embedding_dim), use_hash, dtype,
                               embedding_name, group_name)

    def __hash__(self):
        return hash(self.name)


class DenseFeat(namedtuple('DenseFeat', ['name', 'dimension', 'dtype'])):
    """`dimension` raw float columns passed through to the logit head."""
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super().__new__(cls, name, int(dimension), dtype)

    def __hash__(self):
        return hash(self.name)


class VarLenSparseFeat(namedtuple('VarLenSparseFeat', ['sparsefeat', 'maxlen', 'combiner', 'length_name'])):
    """Accepted for import compatibility only; not on the SATrans path."""
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super().__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda s: s.sparsefeat.name)
    vocabulary_size = property(lambda s: s.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda s: s.sparsefeat.embedding_dim)
    embedding_name = property(lambda s: s.sparsefeat.embedding_name)

    def __hash__(self):
        return hash(self.name)


def build_input_features(feature_columns):
    """name -> (start, end) column slice of the stacked input matrix, in first-seen order;
    a name that appears twice (linear + dnn lists, reference meta_basemodel.py:164-165) keeps
    its first slot."""
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
            raise NotImplementedError("VarLenSparseFeat is not on the SATrans path (reference main.py:103,146)")
        else:
            raise TypeError("Invalid feature column type,got", type(feat))
    return features


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns).keys())
