"""Synthetic Ali-CCP-shaped and Alimama-shaped inputs.

Schemas follow the reference's field lists and id ranges: Ali-CCP field order
reference main.py:99-101 and `data_max` main.py:124-127 (vocabulary_size = data_max + 2,
main.py:182); Alimama field order main.py:143-145 with benchmark-defined maxima
(SURVEY.md §8d config 3).  There is no network for the real datasets; every number
produced from these generators is labelled `"data": "synthetic"`.
"""
from collections import OrderedDict

import numpy as np

from .inputs import DenseFeat, SparseFeat

ALICCP_FIELDS = ['101', '121', '122', '124', '125', '126', '127', '128', '129', '205', '206', '207', '210',
                 '216', '508', '509', '702', '853', '301']
ALICCP_DATA_MAX = {'101': 444861, '121': 97, '122': 13, '124': 2, '125': 7, '126': 3, '127': 3, '128': 2,
                   '129': 4, '205': 4348615, '206': 8993, '207': 695124, '210': 99606, '216': 234880,
                   '508': 8185, '509': 472354, '702': 167813, '853': 91358, '301': 3}

ALIMAMA_FIELDS = ['user_id', 'adgroup_id', 'pid', 'cms_segid', 'cms_group_id', 'final_gender_code', 'age_level',
                  'pvalue_level', 'shopping_level', 'occupation', 'new_user_class_level', 'cate_id',
                  'campaign_id', 'customer', 'brand']
ALIMAMA_DATA_MAX = {'user_id': 1141729, 'adgroup_id': 846811, 'pid': 1, 'cms_segid': 97, 'cms_group_id': 13,
                    'final_gender_code': 2, 'age_level': 7, 'pvalue_level': 3, 'shopping_level': 3,
                    'occupation': 2, 'new_user_class_level': 4, 'cate_id': 12977, 'campaign_id': 423436,
                    'customer': 255875, 'brand': 461497}


class Schema:
    """Field list + id ranges + scenario column of one dataset shape."""

    def __init__(self, name, fields, data_max, dense, domain_col, domain_values, flag, lr):
        self.name, self.fields, self.data_max, self.dense = name, list(fields), dict(data_max), list(dense)
        self.domain_col, self.domain_values, self.flag, self.lr = domain_col, list(domain_values), flag, lr

    def scaled(self, vocab_scale=1.0, min_rows=4):
        """Same schema with every table's row count multiplied by `vocab_scale` (tables that are
        already tiny, and the scenario column, keep their range)."""
        dm = {}
        for f in self.fields:
            m = self.data_max[f]
            if f == self.domain_col or m <= 128:
                dm[f] = m
            else:
                dm[f] = max(min_rows, int(m * vocab_scale))
        return Schema(self.name, self.fields, dm, self.dense, self.domain_col, self.domain_values, self.flag,
                      self.lr)

    @property
    def vocab_sizes(self):
        return [self.data_max[f] + 2 for f in self.fields]          # reference main.py:182

    @property
    def num_domains(self):
        return max(len(self.domain_values), self.data_max[self.domain_col])   # reference main.py:129-133

    def feature_columns(self, embedding_dim=32):
        return ([SparseFeat(f, vocabulary_size=self.data_max[f] + 2, embedding_dim=embedding_dim)
                 for f in self.fields] + [DenseFeat(f, 1) for f in self.dense])

    @property
    def total_rows(self):
        return int(sum(self.vocab_sizes))


def aliccp_schema(vocab_scale=1.0):
    """Ali-CCP shape: 19 sparse fields, scenario column '301' in {1,2,3} (reference main.py:112-114)."""
    s = Schema('aliccp', ALICCP_FIELDS, ALICCP_DATA_MAX, [], '301', [1, 2, 3], 'sota', 0.005)
    return s if vocab_scale == 1.0 else s.scaled(vocab_scale)


def alimama_schema(vocab_scale=1.0):
    """Alimama shape: 15 sparse fields + dense `price`, scenario column shopping_level in {0..3}."""
    s = Schema('alimama', ALIMAMA_FIELDS, ALIMAMA_DATA_MAX, ['price'], 'shopping_level', [0, 1, 2, 3],
               'sota-pos', 0.001)
    return s if vocab_scale == 1.0 else s.scaled(vocab_scale)


def make_inputs(schema, n, seed=1021, dist='uniform', zipf_a=1.05, ctr=0.04):
    """dict name -> 1-D numpy column (int64 ids / float32 dense) and labels (n,) int64.

    dist='uniform': ids uniform on [0, data_max] (worst case for locality);
    dist='zipf': Zipf(a) ranks clipped to the range (realistic skew).
    """
    rng = np.random.default_rng(seed)
    x = OrderedDict()
    for f in schema.fields:
        m = schema.data_max[f]
        if f == schema.domain_col:
            x[f] = rng.choice(np.asarray(schema.domain_values, dtype=np.int64), size=n)
        elif dist == 'uniform':
            x[f] = rng.integers(0, m + 1, size=n, dtype=np.int64)
        elif dist == 'zipf':
            x[f] = np.minimum(rng.zipf(zipf_a, size=n) - 1, m).astype(np.int64)
        else:
            raise ValueError(dist)
    for f in schema.dense:
        x[f] = rng.random(n, dtype=np.float32)
    y = (rng.random(n) < ctr).astype(np.int64)
    return x, y


def stack_columns(x, names):
    """Column-stack a dict of 1-D arrays in `names` order into the (N, n_cols) float32 matrix that the
    reference feeds its model (meta_basemodel.py:221-222,257-264 then `.float()` at :311)."""
    return np.stack([np.asarray(x[k]).astype(np.float32) for k in names], axis=1)
