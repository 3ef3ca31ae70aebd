#!/usr/bin/env python
"""bench.py — SATrans training-step throughput on B200 (BASELINE.json: "SATrans train+infer samples/sec at
1/2/4/8 B200; embedding-gather HBM GB/s"), config 2: EN+MetaNet (meta_mode=QK, d=32, 3 layers, 4 heads,
domain_col=301) on synthetic Ali-CCP-shaped batches, B=8192, full Ali-CCP vocabulary (6,571,961 rows).

One "step" = one optimiser step with the REFERENCE's semantics (meta_basemodel.py:311-328): train-mode
forward (dropout on), BCE(sum), dense L2 over all table rows, backward, dense Adam over every parameter.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this framework (prints ONE JSON line)
    python bench.py --impl reference [--steps K] [--warmup W]        # the reference's algorithm on the host CPUs

Keys beyond the base contract: `infer` (eval-mode scoring samples/s at B=32768), `gather` (stand-alone
multi-field gather GB/s), `phases_ms` (CUDA-event breakdown of the step), `roofline`, `cpu_baseline`.
Timing: CUDA events on the launching streams, W>=3 warm-up steps; every step uses a different batch and
sweeps 3 x 841 MB of table state (> the 126 MB L2), so no explicit L2 flush is needed ("l2": "inputs>L2").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_samples_per_sec"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--vocab-scale', type=float, default=1.0)
    ap.add_argument('--ids', default='uniform', choices=['uniform', 'zipf'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--overlap', action='store_true', help='run the dense optimiser sweep on a second stream, concurrent with forward/backward (default: on the main stream)')
    ap.add_argument('--lazy', action='store_true', help='touched-rows-only table updates (NOT reference semantics)')
    ap.add_argument('--cpu-budget-s', type=float, default=25.0)
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32'],
                    help="forward GEMMs: fp32 CUDA cores | tcgen05 3xTF32 (fp32-level parity, default) | tcgen05 TF32")
    return ap.parse_args()


def workload_config(args, n_gpus):
    return {"workload": "SATrans EN+MetaNet train step (meta_mode=QK, d=32, L=3, H=4, domain_col=301), synthetic "
                        "Ali-CCP-shaped, full vocab 6,571,961 rows" if args.vocab_scale == 1.0 else
                        "SATrans EN+MetaNet train step, Ali-CCP-shaped, vocab x%g" % args.vocab_scale,
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus, "fields": 19, "ids": args.ids,
            "optimizer": "Adam lr=0.005 + L2(emb)=1e-5, " + ("lazy touched-rows-only" if args.lazy else
                                                               "dense every-row (reference semantics)"),
            "dropout": 0.1, "l2": "inputs>L2 (3 x 841 MB swept per step, new batch every step)",
            "forward_gemms": {"fp32": "fp32 FFMA (CUDA cores)", "tf32x3": "tcgen05 3xTF32 error-compensated (fp32-level parity)",
                              "tf32": "tcgen05 TF32"}[args.precision],
            "parallelism": "1 GPU" if n_gpus == 1 else "row-sharded tables + dp%d dense" % n_gpus}


def oracle_setup(args, n_batches):
    from oracle import satrans_oracle as O
    from satrans_b200 import synth
    schema = synth.aliccp_schema(args.vocab_scale)
    cfg = O.OracleConfig(sparse_names=schema.fields, vocab_sizes=schema.vocab_sizes, domain_col='301', num_domains=3,
                         layers=3, heads=4, meta_mode='QK', flag='sota')
    xs, ys = synth.make_inputs(schema, n_batches * args.batch, seed=1021, dist=args.ids, ctr=0.04)
    X = synth.stack_columns(xs, schema.fields)
    return O, schema, cfg, X, ys.astype(np.float32)


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's train step on the host cores
# ------------------------------------------------------------------------------------------------------------
def time_cpu_reference(args, steps, warmup, budget_s):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, schema, cfg, X, y = oracle_setup(args, 4)
    st = O.init_state(cfg, seed=1021)
    tr = O.Trainer(st, cfg, lr=0.005)
    B = args.batch
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y)

    def step(i):
        k = i % 4
        return tr.step(Xt[k * B:(k + 1) * B], yt[k * B:(k + 1) * B], train_mode=True)

    t0 = time.perf_counter()
    step(0)
    first = time.perf_counter() - t0
    n_warm = max(0, min(warmup, int(budget_s * 0.25 / max(first, 1e-3))) - 1)
    for i in range(n_warm):
        step(1 + i)
    n = max(1, min(steps, int(budget_s * 0.75 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        step(1 + n_warm + i)
    dt = (time.perf_counter() - t0) / n
    return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d timed full steps (B=%d, full vocab, train-mode forward + BCE + dense L2 + backward + dense "
                      "torch.optim.Adam) of oracle/satrans_oracle.Trainer after %d warm-up, torch %d threads"
                      % (n, B, n_warm + 1, cores),
            "ms_per_step": dt * 1e3, "steps": n, "warmup": n_warm + 1}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = time_cpu_reference(args, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            parts = [t.strip() for t in ln.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    from satrans_b200 import SATrans, synth
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)') if 'hbm_gbs' in peaks else \
        (6650.0, 'fallback (B200_PROFILING.md)')

    schema = synth.aliccp_schema(args.vocab_scale)
    cols = schema.feature_columns(32)
    model = SATrans(cols, cols, ['301'], [3], att_layer_num=0, domain_att_layer_num=3, att_head_num=4, meta_mode='QK',
                    seed=1021, device=dev, flag='sota', table_shard=(rank, world) if world > 1 else None)
    model.dense_table_updates = not args.lazy
    model.compile(torch.optim.Adam(model.parameters(), lr=0.005), "binary_crossentropy",
                  metrics=["binary_crossentropy", "auc"])
    eng = model.engine
    eng.precision = args.precision
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    if world > 1:
        from satrans_b200.dist import DistContext
        eng.attach_dist(DistContext())
        eng.prepare_training(B)
    n_batches = 16
    xs, ys = synth.make_inputs(schema, n_batches * B, seed=1021 + rank, dist=args.ids, ctr=0.04)
    X_host = torch.from_numpy(synth.stack_columns(xs, schema.fields)).pin_memory()
    y_host = torch.from_numpy(ys.astype(np.float32)).pin_memory()
    Xd, yd = X_host.to(dev), y_host.to(dev)
    hp = model._opt_hparams
    model.train()

    def step(i, overlap=args.overlap):
        k = i % n_batches
        return eng.train_step(Xd[k * B:(k + 1) * B], yd[k * B:(k + 1) * B], lr=hp['lr'], betas=hp['betas'], eps=hp['eps'],
                              opt_kind=hp['kind'], l2_emb=1e-5, dense_mode=not args.lazy, training=True, dropout_p=0.1, seed=1021 + i,
                              overlap=overlap)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    # ---- timed region: K steps, device-resident inputs ------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    eng.timing = {}
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        _, w = step(W + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / K
    launches = eng.launches - l0
    phases = eng.timing_ms()
    eng.timing = None
    n_unique = float(w['scalars'][2].item())
    clk = clocks.stop() if clocks else {}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world / (ms * 1e-3)

    # ---- e2e: the public per-batch API with HOST buffers (pinned), H2D + D2H inside the timed region -------
    for i in range(3):
        model.train_on_batch(X_host[:B], y_host[:B])
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        k = (i + 3) % n_batches
        model.train_on_batch(X_host[k * B:(k + 1) * B], y_host[k * B:(k + 1) * B])
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / K
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e = {"value": B * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(B * X_host.shape[1] * 4 + B * 4), "d2h_bytes_per_step": 16,
           "api": "SATrans.train_on_batch(x_host, y_host) -> (bce, reg)"}
    if rank != 0:
        if world > 1:
            dist.barrier()          # keep this rank's table shard mapped until rank 0 is done with the extras
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the dense L2+Adam sweep (HBM-bound) ---------------------------------
    rows = model._tables.shape[0]                                    # rows of this rank's shard
    sweep_bytes = (rows - n_unique) * 6 * 128 + (rows / 8)           # p,m,v read + written once; bitmap read
    roof = None
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel (profiles/)
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))['k_adam_sweep']
        if world == 1 and args.vocab_scale == 1.0:
            traffic = tr['dram_bytes_read'] + tr['dram_bytes_write']
    except Exception:
        pass
    if 'adam_sweep' in phases:
        ach = sweep_bytes / (phases['adam_sweep'] * 1e-3) / 1e9
        roof = {"kernel": "k_adam_sweep", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": phases['adam_sweep'],
                "note": "6 x 128 B per untouched table row; timed with CUDA events on its own stream while "
                        "forward/backward run concurrently on the main stream" if args.overlap else
                        "6 x 128 B per untouched table row; CUDA events around its launch on the main stream (the ~20 us "
                        "touched-row and dense-parameter updates run beside it on a second stream)"}

    # ---- extra: stand-alone gather GB/s and eval-mode scoring ----------------------------------------------
    extra = {}
    try:
        Bg = 262144
        xg, _ = synth.make_inputs(schema, Bg, seed=7, dist=args.ids)
        Xg = torch.from_numpy(synth.stack_columns(xg, schema.fields)).to(dev).to(torch.int32)
        for _ in range(3):
            eng.gather(Xg)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(10):
            eng.gather(Xg)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 10
        gb = Bg * (2 * 19 * 128 + 19 * 4) / (gms * 1e-3) / 1e9
        extra["gather"] = {"kernel": "k_gather", "batch": Bg, "ms": gms, "achieved": gb, "unit": "GB/s", "peak": hbm_peak,
                           "frac": gb / hbm_peak, "bytes_per_sample": 2 * 19 * 128 + 19 * 4}
        model.eval()
        Bi = 32768
        Xi = Xg[:Bi].contiguous()
        for _ in range(3):
            eng.forward(Xi, training=False)
        torch.cuda.synchronize()
        g0.record()
        for _ in range(10):
            eng.forward(Xi, training=False)
        g1.record()
        torch.cuda.synchronize()
        ims = g0.elapsed_time(g1) / 10
        extra["infer"] = {"batch": Bi, "ms": ims, "value": Bi / (ims * 1e-3), "unit": UNIT,
                          "gflops": Bi * 1540672 / (ims * 1e-3) / 1e9, "precision": args.precision}
        for prec in ('fp32', 'tf32x3', 'tf32'):
            if prec == args.precision:
                continue
            eng.precision = prec
            for _ in range(3):
                eng.forward(Xi, training=False)
            torch.cuda.synchronize()
            g0.record()
            for _ in range(10):
                eng.forward(Xi, training=False)
            g1.record()
            torch.cuda.synchronize()
            pms = g0.elapsed_time(g1) / 10
            extra["infer_" + prec] = {"batch": Bi, "ms": pms, "value": Bi / (pms * 1e-3), "unit": UNIT}
        eng.precision = args.precision
    except Exception as e:  # the headline number must survive a failure of the extras
        extra["extra_error"] = repr(e)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            del Xg
        except Exception:
            pass
        r = time_cpu_reference(args, 3, 1, budget_s=args.cpu_budget_s)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "phases_ms": phases}
    line.update(extra)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, text.encode())
    else:
        sys.stdout.write(text)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse()
    # Everything else that writes to fd 1 (NCCL prints "NCCL version ..." from C at communicator creation, libraries may
    # warn) is sent to stderr, so that stdout carries exactly the JSON line.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
