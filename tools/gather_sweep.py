#!/usr/bin/env python
"""Time the stand-alone multi-field gather (sat_gather_fwd) for every kernel variant (SAT_GATHER_VARIANT) at B = 8,192 and
262,144 on the full Ali-CCP vocabulary, checking bit-exactness against torch indexing.  One subprocess per variant.

    python tools/gather_sweep.py            # prints one JSON line per (variant, batch)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch
    from satrans_b200 import SATrans, synth
    v = int(os.environ.get('SAT_GATHER_VARIANT', '0'))
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
    schema = synth.aliccp_schema()
    cols = schema.feature_columns(32)
    model = SATrans(cols, cols, ['301'], [3], att_layer_num=0, domain_att_layer_num=3, att_head_num=4, meta_mode='QK',
                    seed=1021, device='cuda:0', flag='sota')
    eng = model.engine
    for dist in ('uniform', 'zipf'):
        xg, _ = synth.make_inputs(schema, 262144, seed=7, dist=dist)
        Xg = torch.from_numpy(np.stack([np.asarray(xg[k]) for k in schema.fields], axis=1).astype(np.int32)).cuda()
        off = torch.as_tensor(model._row_off[:-1], device='cuda')
        for Bg, reps in ((8192, 64), (262144, 20)):
            n_slices = 262144 // Bg
            sl = [Xg[i * Bg:(i + 1) * Bg] for i in range(n_slices)]
            E = eng.gather(sl[0])
            ref = model._tables[(sl[0].long() + off[None, :]).reshape(-1)].view(Bg, 19, 32)
            exact = bool(torch.equal(E, ref))
            for i in range(3):
                eng.gather(sl[i % n_slices])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(reps):
                eng.gather(sl[i % n_slices])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gb = Bg * (2 * 19 * 128 + 19 * 4) / (ms * 1e-3) / 1e9
            print(json.dumps({"variant": v, "ids": dist, "batch": Bg, "ms": round(ms, 5), "GBps": round(gb, 1),
                              "frac": round(gb / peak, 4), "bit_exact": exact}), flush=True)


if __name__ == '__main__':
    if os.environ.get('SAT_GATHER_CHILD'):
        child()
    else:
        for v in (sys.argv[1:] or ["2", "5", "6"]):
            env = dict(os.environ, SAT_GATHER_VARIANT=v, SAT_GATHER_CHILD='1')
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=300)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                print(json.dumps({"variant": int(v), "error": r.stderr[-500:]}))
