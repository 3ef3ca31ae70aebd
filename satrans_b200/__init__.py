"""satrans_b200 — B200-native SATrans training/scoring hot path (drop-in for the reference's
`models/satrans.py` on `models/meta_basemodel.py`).

    from satrans_b200 import SATrans, SparseFeat, DenseFeat, get_feature_names
"""
from .inputs import DenseFeat, SparseFeat, VarLenSparseFeat, build_input_features, get_feature_names  # noqa: F401
from .basemodel import BaseModel, History  # noqa: F401
from .satrans import SATrans, Meta_Transformer_Layer, MetaNet, DNN_v2  # noqa: F401
from .selfatt import InteractingLayer, MDR_InteractingLayer, SelfAttention_Layer  # noqa: F401
from .baselines import AutoInt, Star_Trans  # noqa: F401

__all__ = ['SATrans', 'BaseModel', 'Meta_Transformer_Layer', 'MetaNet', 'DNN_v2', 'SparseFeat', 'DenseFeat',
           'VarLenSparseFeat', 'build_input_features', 'get_feature_names', 'History', 'SelfAttention_Layer', 'MDR_InteractingLayer', 'InteractingLayer', 'Star_Trans', 'AutoInt']
