"""Diagnostic (not collected by pytest): do the tcgen05 3xTF32 and the CUDA-core fp32 forward take different MetaNet ReLU
decisions on the same model / batch, and how far apart are the two paths' gradients when they do?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from satrans_b200 import synth
from test_gpu_api import _small, _model
from gpu_util import dense_grads

schema, cfg, st = _small()
xs, ys = synth.make_inputs(schema, 1000, seed=11, ctr=0.3)
X = torch.from_numpy(synth.stack_columns(xs, schema.fields + schema.dense)).cuda(); y = torch.from_numpy(ys.astype(np.float32)).cuda()
model = _model(cfg, st)
eng = model.engine
L, F = cfg.layers, len(schema.fields)
for s0 in range(0, 1000, 256):
    xb, yb = X[s0:s0 + 256].contiguous(), y[s0:s0 + 256].contiguous()
    res = {}
    for prec in ('fp32', 'tf32x3'):
        eng.precision = prec
        yp, w = eng.loss_and_grads(xb, yb, training=False)
        torch.cuda.synchronize()
        stride = w.cap_B * F * 32
        masks = w['acts'][(L * 5 + 1) * stride:].view(torch.int32).clone()
        res[prec] = (dense_grads(model, w), masks, yp.clone())
    (g0, m0, y0), (g1, m1, y1) = res['fp32'], res['tf32x3']
    nb = xb.shape[0] * F
    flips = 0
    for l in range(L):
        for side in range(2):
            a = m0[((l * 2 + side) * (stride // 32)) * 2:][:nb * 2]; b = m1[((l * 2 + side) * (stride // 32)) * 2:][:nb * 2]
            x = (a ^ b)
            flips += sum(bin(int(v) & 0xffffffff).count('1') for v in x[x != 0].tolist())
    worst = sorted(((g1[k] - g0[k]).abs().max().item() / max(g0[k].abs().max().item(), 1e-30), k) for k in g0)[::-1][:3]
    print('batch', s0 // 256, 'relu decisions that differ:', flips, 'of', nb * 64 * 2 * L, '| y_pred diff %.1e |' % (y1 - y0).abs().max().item(),
          ' '.join('%s %.1e' % (k.replace('domain_int_layers.', 'L'), e) for e, k in worst), flush=True)
