"""Diagnostic (not collected by pytest): after ONE optimiser step, do the tcgen05 and the CUDA-core paths still agree on the
next batch's gradients?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from oracle import satrans_oracle as O
from satrans_b200 import synth
from test_gpu_api import _small, _model
from gpu_util import dense_grads

schema, cfg, st = _small()
xs, ys = synth.make_inputs(schema, 1000, seed=11, ctr=0.3)
Xh = torch.from_numpy(synth.stack_columns(xs, schema.fields + schema.dense)); yh = torch.from_numpy(ys.astype(np.float32))
X, y = Xh.cuda(), yh.cuda()
for step_prec in ('tf32x3', 'fp32'):
    model = _model(cfg, st)
    eng = model.engine
    eng.side_streams = False
    model.dropout_rate = 0.0
    model.compile(torch.optim.Adam(model.parameters(), lr=0.005, eps=1e-4), "binary_crossentropy")
    hp = model._opt_hparams
    tr = O.Trainer({k: v.clone() for k, v in st.items()}, cfg, lr=0.005, eps=1e-4)
    eng.precision = step_prec
    model.train()
    eng.train_step(X[:256].contiguous(), y[:256].contiguous(), lr=hp['lr'], betas=hp['betas'], eps=hp['eps'], opt_kind=hp['kind'],
                   l2_emb=model.l2_reg_embedding, dense_mode=True, training=True, dropout_p=0.0, seed=1)
    tr.step(Xh[:256], yh[:256], train_mode=False)
    torch.cuda.synchronize()
    _, _, _, gref = O.loss_and_grads(tr.params, Xh[256:512], yh[256:512], cfg, with_reg=False)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    pr = sorted(((sd[k] - tr.params[k].detach()).abs().max().item(), k) for k in O.trainable_keys(cfg))[::-1][:2]
    print('stepped with', step_prec, '| params vs oracle:', pr, flush=True)
    fresh = _model(cfg, {k: sd[k] for k in st})
    fe = fresh.engine; fe.side_streams = False; fe.precision = 'tf32x3'
    ypf, wf = fe.loss_and_grads(X[256:512].contiguous(), y[256:512].contiguous(), training=False)
    torch.cuda.synchronize()
    gf = dense_grads(fresh, wf)
    worst = sorted(((gf[k] - gref[k].detach()).abs().max().item() / max(gref[k].abs().max().item(), 1e-30), k) for k in gf if k in gref)[::-1][:3]
    print('stepped with', step_prec, '| FRESH model with the stepped params, tf32x3 vs oracle:', ' '.join('%s %.1e' % (k.replace('domain_int_layers.', 'L'), e) for e, k in worst), flush=True)
    res = {}
    for prec in ('fp32', 'tf32x3'):
        eng.precision = prec
        yp, w = eng.loss_and_grads(X[256:512].contiguous(), y[256:512].contiguous(), training=False)
        torch.cuda.synchronize()
        res[prec] = dense_grads(model, w)
        F_ = len(schema.fields); stride = w.cap_B * F_ * 32
        res[prec + '_mask'] = w['acts'][(cfg.layers * 5 + 1) * stride:].view(torch.int32).clone()
        res[prec + '_z'] = w['acts'][:(cfg.layers * 5 + 1) * stride].clone()
    x = res['fp32_mask'] ^ res['tf32x3_mask']
    nfl = sum(bin(int(v) & 0xffffffff).count('1') for v in x[x != 0].tolist())
    print('stepped with', step_prec, '| relu decisions differing between the fp32 and tf32x3 forward at the SAME params:', nfl, '| max |activation diff| %.2e' % (res['fp32_z'] - res['tf32x3_z']).abs().max().item(), flush=True)
    for prec in ('fp32', 'tf32x3'):
        worst = sorted(((res[prec][k] - gref[k].detach()).abs().max().item() / max(gref[k].abs().max().item(), 1e-30), k) for k in res[prec] if k in gref)[::-1][:3]
        print('stepped with', step_prec, '| grads of', prec, 'vs oracle:', ' '.join('%s %.1e' % (k.replace('domain_int_layers.', 'L'), e) for e, k in worst), flush=True)
