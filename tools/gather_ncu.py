#!/usr/bin/env python
"""Five launches of the stand-alone multi-field gather at B = 262,144 (uniform ids, full Ali-CCP vocabulary) and five at
B = 8,192 on rotating input slices -- the target of `ncu --set full -k regex:k_gather -s 3 -c 1` (262,144) / `-s 8 -c 1` (8,192)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satrans_b200 import SATrans, synth  # noqa: E402

schema = synth.aliccp_schema()
cols = schema.feature_columns(32)
model = SATrans(cols, cols, ['301'], [3], att_layer_num=0, domain_att_layer_num=3, att_head_num=4, meta_mode='QK', seed=1021,
                device='cuda:0', flag='sota')
xg, _ = synth.make_inputs(schema, 5 * 262144, seed=7, dist='uniform')
Xg = torch.from_numpy(np.stack([np.asarray(xg[k]) for k in schema.fields], axis=1).astype(np.int32)).cuda()
for i in range(5):
    model.engine.gather(Xg[i * 262144:(i + 1) * 262144])
for i in range(5):
    model.engine.gather(Xg[i * 8192:(i + 1) * 8192])
torch.cuda.synchronize()
