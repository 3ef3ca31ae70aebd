"""Diagnostic (not collected by pytest): repeat the fit-vs-oracle scenario of test_fit_matches_oracle_training_loop in one
process and print, per repetition, the worst keys of assert_adam_close's 'tight' fraction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from oracle import satrans_oracle as O
from satrans_b200 import synth
from test_gpu_api import _small, _model

schema, cfg, st = _small()
n, bs, lr = 1000, 256, 0.005
xs, ys = synth.make_inputs(schema, n, seed=11, ctr=0.3)
EPS = float(os.environ.get("ADAM_EPS", "1e-8"))
tr = O.Trainer({k: v.clone() for k, v in st.items()}, cfg, lr=lr, eps=EPS)
X = torch.from_numpy(synth.stack_columns(xs, schema.fields + schema.dense)); y = torch.from_numpy(ys.astype(np.float32))
for s0 in range(0, n, bs):
    tr.step(X[s0:s0 + bs], y[s0:s0 + bs], train_mode=False)
side = (sys.argv[2] == '1') if len(sys.argv) > 2 else True
prec = sys.argv[3] if len(sys.argv) > 3 else 'tf32x3'
print('side_streams', side, 'precision', prec)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    model = _model(cfg, st)
    model.engine.side_streams = side
    model.engine.precision = prec
    model.dropout_rate = 0.0
    model.compile(torch.optim.Adam(model.parameters(), lr=lr, eps=EPS), "binary_crossentropy")
    model.fit(x=xs, y=ys, batch_size=bs, epochs=1, verbose=0, shuffle=False)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    rows = []
    for k in O.trainable_keys(cfg):
        ref = tr.params[k].detach(); got = sd[k]
        err = (got - ref).abs().flatten()
        tight = 2e-5 * max(ref.abs().max().item(), 1e-3) + 1e-7
        rows.append(((err <= tight).float().mean().item(), err.max().item(), k))
    rows.sort()
    print('rep', rep, ' '.join('%s frac=%.4f max=%.2e' % (k.replace('domain_int_layers.', 'L'), f, m) for f, m, k in rows[:5]), flush=True)
