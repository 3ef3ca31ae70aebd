#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations that bench.py's single headline line does not cover
(SURVEY.md §8d).  One JSON line per measurement on stdout; run on one B200:

    python bench_configs.py [--only infer_sweep,selfatt,baselines,variants,alimama,zipf,lazy,config4,fit]

  infer_sweep : config 5 — eval-mode scoring, batch 1 Ki .. 1 Mi, meta_mode QK and QKV (identical math in the
                reference: satrans.py has no 'V' branch), fp32 CUDA-core vs tcgen05 3xTF32 vs tcgen05 TF32
  alimama     : config 3 — ENP+MetaNet (flag 'sota-pos', scenario column shopping_level, dense price), train step B=8192
  zipf        : config 2 with Zipf(1.05) ids instead of uniform (realistic skew: hot rows)
  variants    : config 2 with the reference's ablation flags: 'gate' / 'bilinear' (CUDA-core token kernels) and 'onlyemb' (tcgen05 path)
  lazy        : config 2 with touched-rows-only table updates (NOT the reference's semantics; reported separately)
  selfatt     : config 5's "SelfAttention_Layer x 3" (reference submodules.py:178-238) on (B, 19, 32) inputs, eval mode
  config4     : config 4 on one GPU — vocab x16 (105 M rows, 13.5 GB of tables + 27 GB of Adam state), B=65,536, int32 ids
  fit         : config 1(b) — fit() / predict() end to end from host numpy columns (verbose=0 and verbose=1 with sklearn metrics)
  baselines   : config 5's "AutoInt / STAR-Trans baselines through the field-attention kernel": satrans_b200.AutoInt / Star_Trans
                on the Ali-CCP-shaped full vocabulary -- scoring sweep (forward under no_grad: gather + attention layers + head,
                this package's kernels only) and one training step (B = 8192; gather / attention forward + backward in the kernels,
                head / DNN tower / torch.optim.Adam in torch)

All timings: CUDA events, >=3 warm-up iterations, inputs resident in HBM.
"""
import argparse
import json
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def timeit(fn, iters, warm=3, finish=None):
    import torch
    for _ in range(warm):
        fn()
    if finish:
        finish()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    if finish:
        finish()                  # (e.g. close the last table-sweep epoch inside the timed region)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def build(schema, flag, mode, lr, precision):
    import torch
    from satrans_b200 import SATrans
    cols = schema.feature_columns(32)
    model = SATrans(cols, cols, [schema.domain_col], [schema.num_domains], att_layer_num=0, domain_att_layer_num=3,
                    att_head_num=4, meta_mode=mode, seed=1021, device='cuda:0', flag=flag)
    model.compile(torch.optim.Adam(model.parameters(), lr=lr), "binary_crossentropy")
    model.engine.precision = precision
    return model


def train_bench(name, schema, flag, lr, ids, lazy, precision, B=8192, steps=20, int_ids=False, nb=8):
    import torch
    from satrans_b200 import synth
    model = build(schema, flag, 'QK', lr, precision)
    model.dense_table_updates = not lazy
    xs, ys = synth.make_inputs(schema, nb * B, seed=1021, dist=ids, ctr=0.04)
    if int_ids:      # ids >= 2^24 are not exact in the reference's fp32 input matrix: int32 matrix instead
        assert not schema.dense
        X = torch.from_numpy(np.stack([xs[f].astype(np.int32) for f in schema.fields], axis=1)).cuda()
    else:
        X = torch.from_numpy(synth.stack_columns(xs, schema.fields + schema.dense)).cuda()
    y = torch.from_numpy(ys.astype(np.float32)).cuda()
    eng, hp = model.engine, model._opt_hparams
    it = [0]

    def step():
        k = it[0] % nb
        it[0] += 1
        eng.train_step(X[k * B:(k + 1) * B], y[k * B:(k + 1) * B], lr=hp['lr'], l2_emb=1e-5, dense_mode=not lazy,
                       training=True, dropout_p=0.1, seed=it[0], defer=1 if lazy else 8)
    for _ in range(5):
        step()
    eng.sync_tables()
    eng.timing = {}
    ms = timeit(step, steps, warm=0, finish=eng.sync_tables)
    phases = eng.timing_ms()
    print(json.dumps({"bench": name, "metric": "train_samples_per_sec", "value": B / (ms * 1e-3), "unit": "samples/s",
                      "ms_per_step": ms, "batch": B, "rows": schema.total_rows, "fields": len(schema.fields),
                      "flag": flag, "ids": ids, "table_updates": "lazy" if lazy else "dense (reference semantics), sweep every 8 steps",
                      "precision": precision, "phases_ms": phases}), flush=True)
    del model
    torch.cuda.empty_cache()


def infer_sweep(precisions):
    import torch
    from satrans_b200 import synth
    schema = synth.aliccp_schema()
    for mode in ('QK', 'QKV'):
        model = build(schema, 'sota', mode, 0.005, 'fp32').eval()
        for B in (1024, 4096, 16384, 65536, 262144, 1048576):
            xs, _ = synth.make_inputs(schema, B, seed=7)
            X = torch.from_numpy(np.stack([xs[f].astype(np.int32) for f in schema.fields], axis=1)).cuda()
            for prec in precisions:
                model.engine.precision = prec
                ms = timeit(lambda: model.engine.forward(X, training=False), 10 if B <= 65536 else 4)
                print(json.dumps({"bench": "infer_sweep", "metric": "infer_samples_per_sec", "value": B / (ms * 1e-3),
                                  "unit": "samples/s", "ms": ms, "batch": B, "meta_mode": mode, "precision": prec,
                                  "gflops": B * 1540672 / (ms * 1e-3) / 1e9,
                                  "hbm_gbs_algorithmic": B * 2508 / (ms * 1e-3) / 1e9}), flush=True)
            del X
        del model
        torch.cuda.empty_cache()


def selfatt_sweep():
    import torch
    from satrans_b200 import SelfAttention_Layer
    layers = [SelfAttention_Layer(32, head_num=2, use_res=True, device='cuda:0').eval() for _ in range(3)]
    for B in (1024, 16384, 65536, 262144, 1048576):
        x = torch.randn(B, 19, 32, device='cuda:0') * 0.05

        def run():
            with torch.no_grad():
                h = x
                for lay in layers:
                    h = lay(h)
            return h
        ms = timeit(run, 10 if B <= 65536 else 4)
        print(json.dumps({"bench": "selfatt_x3", "metric": "infer_samples_per_sec", "value": B / (ms * 1e-3), "unit": "samples/s",
                          "ms": ms, "batch": B, "fields": 19, "heads": 2,
                          "hbm_gbs_algorithmic": 3 * 2 * B * 19 * 128 / (ms * 1e-3) / 1e9}), flush=True)
        del x
    torch.cuda.empty_cache()


def baselines_bench():
    import torch
    from satrans_b200 import AutoInt, Star_Trans, synth
    schema = synth.aliccp_schema()
    cols = schema.feature_columns(32)
    for name in ('AutoInt', 'Star_Trans'):
        if name == 'AutoInt':
            model = AutoInt(cols, cols, att_layer_num=3, att_head_num=2, device='cuda:0', seed=1021)
        else:
            model = Star_Trans(cols, cols, schema.domain_col, schema.num_domains, att_layer_num=2, domain_att_layer_num=1,
                               att_head_num=4, use_linear=False, device='cuda:0', seed=1021)
        model.eval()
        for B in (16384, 262144):
            xs, _ = synth.make_inputs(schema, B, seed=7)
            X = torch.from_numpy(np.stack([xs[f].astype(np.int32) for f in schema.fields], axis=1)).cuda()

            def score():
                with torch.no_grad():
                    return model(X)
            ms = timeit(score, 10 if B <= 65536 else 4)
            print(json.dumps({"bench": "baseline_scoring", "model": name, "metric": "infer_samples_per_sec", "value": B / (ms * 1e-3),
                              "unit": "samples/s", "ms": ms, "batch": B}), flush=True)
            del X
        B = 8192
        xs, ys = synth.make_inputs(schema, B, seed=9, ctr=0.04)
        X = torch.from_numpy(np.stack([xs[f].astype(np.int32) for f in schema.fields], axis=1)).cuda()
        y = torch.from_numpy(ys.astype(np.float32)).cuda()
        model.train()
        model.compile(torch.optim.Adam(model.parameters(), lr=0.005), 'binary_crossentropy')
        ms = timeit(lambda: model._train_step(X, y), 5)
        print(json.dumps({"bench": "baseline_train_step", "model": name, "metric": "train_samples_per_sec", "value": B / (ms * 1e-3),
                          "unit": "samples/s", "ms": ms, "batch": B,
                          "note": "reference semantics (dense L2 + torch.optim.Adam on all 6.57 M rows every step); gather and "
                                  "attention layers forward + backward in this package's kernels, head / tower / optimiser in torch"}),
              flush=True)
        del model, X
        torch.cuda.empty_cache()


def fit_bench():
    """fit() / predict() through the reference-facing API with host numpy columns (reference meta_basemodel.py:200-385,401-517)."""
    import time
    import torch
    from satrans_b200 import synth
    schema = synth.aliccp_schema()
    model = build(schema, 'sota', 'QK', 0.005, 'tf32x3')
    N = 32 * 8192
    xs, ys = synth.make_inputs(schema, N, seed=1021, ctr=0.04)
    for verbose in (0, 1):
        # (compile() starts from fresh optimiser state, like a new torch.optim instance: warm up after it)
        model.compile(torch.optim.Adam(model.parameters(), lr=0.005), "binary_crossentropy",
                      metrics=["binary_crossentropy", "auc"] if verbose else None)
        model.fit(xs, ys, batch_size=8192, epochs=1, verbose=0, shuffle=False)        # warm-up epoch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(xs, ys, batch_size=8192, epochs=1, verbose=verbose, shuffle=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"bench": "fit_e2e", "metric": "train_samples_per_sec", "value": N / dt, "unit": "samples/s",
                          "rows": N, "batch": 8192, "verbose": verbose, "seconds": dt,
                          "note": "host numpy dict -> fit(); includes column stacking, shuffling, H2D, per-epoch metrics"}), flush=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = model.predict(xs, batch_size=32768)
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "predict_e2e", "metric": "infer_samples_per_sec", "value": N / dt, "unit": "samples/s", "rows": N,
                      "batch": 32768, "seconds": dt, "out": [list(p.shape), str(p.dtype)]}), flush=True)
    del model
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='infer_sweep,selfatt,baselines,variants,alimama,zipf,lazy,config4,fit')
    args = ap.parse_args()
    from satrans_b200 import synth
    only = set(args.only.split(','))
    if 'alimama' in only:
        for prec in ('tf32x3', 'fp32'):
            train_bench('alimama_enp_train', synth.alimama_schema(), 'sota-pos', 0.001, 'uniform', False, prec, steps=24)
    if 'zipf' in only:
        train_bench('aliccp_en_train_zipf', synth.aliccp_schema(), 'sota', 0.005, 'zipf', False, 'tf32x3', steps=24)
    if 'lazy' in only:
        train_bench('aliccp_en_train_lazy', synth.aliccp_schema(), 'sota', 0.005, 'uniform', True, 'tf32x3')
    if 'config4' in only:
        train_bench('aliccp_x16_config4_train', synth.aliccp_schema(16.0), 'sota', 0.005, 'uniform', False, 'tf32x3', B=65536,
                    steps=8, int_ids=True, nb=2)
    if 'variants' in only:      # the reference's ablation flags (satrans.py:59-81,158-176) on the config-2 step
        train_bench('aliccp_en_train_gate', synth.aliccp_schema(), 'sota-gate', 0.005, 'uniform', False, 'fp32', steps=16)
        train_bench('aliccp_en_train_bilinear', synth.aliccp_schema(), 'sota-bilinear', 0.005, 'uniform', False, 'fp32', steps=16)
        train_bench('aliccp_en_train_onlyemb', synth.aliccp_schema(), 'sota-onlyemb', 0.005, 'uniform', False, 'tf32x3', steps=16)
    if 'fit' in only:
        fit_bench()
    if 'selfatt' in only:
        selfatt_sweep()
    if 'baselines' in only:
        baselines_bench()
    if 'infer_sweep' in only:
        infer_sweep(('fp32', 'tf32x3', 'tf32'))


if __name__ == '__main__':
    main()
