#!/usr/bin/env python
"""bench.py — SATrans training-step throughput on B200 (BASELINE.json: "SATrans train+infer samples/sec at
1/2/4/8 B200; embedding-gather HBM GB/s"), config 2: EN+MetaNet (meta_mode=QK, d=32, 3 layers, 4 heads,
domain_col=301) on synthetic Ali-CCP-shaped batches, B=8192, full Ali-CCP vocabulary (6,571,961 rows).

One "step" = one optimiser step with the REFERENCE's semantics (meta_basemodel.py:311-328): train-mode
forward (dropout on), BCE(sum), dense L2 over all table rows, backward, dense Adam over every parameter.
The rows no batch read receive their L2 + Adam updates in sweep epochs of `--sweep-interval` (8) steps, each row taken
through its steps one after the other: bit-identical to updating them every step (tests/test_gpu_parity.py::
test_deferred_table_sweep_is_bit_identical); the last epoch is closed INSIDE every timed region.
`value` = the step replayed from a CUDA graph on device-resident batches; `e2e` = SATrans.fit() -- the reference's public
training call -- on pinned host data (every batch copied H2D one step ahead on a copy stream, every step's scalars copied
back D2H, fit() returning after the tables are in sync and the epoch loss is on the host); N > 1: train_on_batch per step.

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this framework (prints ONE JSON line)
    python bench.py --impl reference [--steps K] [--warmup W]        # the reference itself on the host CPUs
    python bench.py --config 4 --gpus N                              # BASELINE config 4: vocab x (4 N), B = 65,536 per GPU

The reference arm runs the UNMODIFIED reference (oracle/_ref, vendored by oracle/build_ref.py; `kind: "reference"`)
through its own public API -- SATrans(...).compile(Adam).fit(x, y, batch_size) -- or, where oracle/_ref is absent, the
oracle port of the same step (`kind: "port"`).

Keys beyond the base contract: `infer` (eval-mode scoring samples/s at B=32768), `phases_ms` (CUDA-event breakdown of
the step), `roofline` (the kernel that moves the most bytes = the every-row L2+Adam sweep, HBM-bound; `roofline.others` = the
stand-alone multi-field gather at B = 8,192 and 262,144 with its ncu DRAM bytes beside the algorithmic bytes;
`roofline.token_kernels` = pipe utilisation of the three latency-bound token kernels from the ncu captures), `cpu_baseline`, `eager_gpu_baseline` (the
reference's eager PyTorch path on the same B200), `bf16` (the kind::f16 variant, never the headline).
Timing: CUDA events on the launching streams, W>=3 warm-up steps; every step uses a different batch whose working set
(20 MB of gathered rows out of 841 MB of tables + ~400 MB of saved activations / gradients) exceeds the 126 MB L2, and
every sweep streams 3 x 841 MB of table state, so no explicit L2 flush is needed ("l2": "inputs>L2").
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
    ap.add_argument('--steps', type=int, default=256)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--vocab-scale', type=float, default=None)
    ap.add_argument('--config', type=int, default=2, choices=[2, 4],
                    help='BASELINE.json config: 2 = full Ali-CCP vocab, B=8192 per GPU (headline); 4 = scaled vocab '
                         '(x 4 per GPU, so the shard per GPU stays constant), B=65536 per GPU, int32 ids')
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of the CUDA-graph replay of the step')
    ap.add_argument('--ids', default='uniform', choices=['uniform', 'zipf'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--overlap', action='store_true', help='run the dense optimiser sweep on a second stream, concurrent with forward/backward (default: on the main stream)')
    ap.add_argument('--lazy', action='store_true', help='touched-rows-only table updates (NOT reference semantics)')
    ap.add_argument('--cpu-budget-s', type=float, default=25.0)
    ap.add_argument('--sweep-interval', type=int, default=8,
                    help='the every-row table update (L2 + Adam over rows no batch read) runs once per this many steps, each row '
                         'receiving its steps one after the other (bit-identical parameters; single GPU); 1 = after every step')
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32'],
                    help="forward GEMMs: fp32 CUDA cores | tcgen05 3xTF32 (fp32-level parity, default) | tcgen05 TF32")
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.batch is None:
        args.batch = 8192 if args.config == 2 else 65536
    if args.vocab_scale is None:
        args.vocab_scale = 1.0 if args.config == 2 else 4.0 * world
    return args


def workload_config(args, n_gpus):
    return {"workload": "SATrans EN+MetaNet train step (meta_mode=QK, d=32, L=3, H=4, domain_col=301), synthetic "
                        "Ali-CCP-shaped, full vocab 6,571,961 rows" if args.vocab_scale == 1.0 else
                        "SATrans EN+MetaNet train step, Ali-CCP-shaped, vocab x%g (BASELINE config %d)" % (args.vocab_scale, args.config),
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus, "fields": 19, "ids": args.ids,
"step_launch": "cuda graph replay" if (n_gpus == 1 and not args.no_graph and not args.overlap) else "eager launches",
            "optimizer": "Adam lr=0.005 + L2(emb)=1e-5, " + ("lazy touched-rows-only" if args.lazy else
                                                               "dense every-row (reference semantics)"),
            "table_sweep_interval": (args.sweep_interval if (not args.lazy and not args.overlap) else 1),
            "dropout": 0.1, "l2": "inputs>L2 (3 x %.0f MB of table state swept per sweep per GPU, new batch every step; the "
                                  "per-step working set -- 20 MB of rows gathered from 841 MB of tables + ~340 MB of saved "
                                  "activations / gradients -- exceeds the 126 MB L2)" % (841.2 * args.vocab_scale / n_gpus),
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
# reference arm / cpu_baseline / eager_gpu_baseline
# ------------------------------------------------------------------------------------------------------------
def reference_available():
    return all(os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'models', f))
               for f in ('satrans.py', 'meta_basemodel.py', 'submodules.py'))


def time_reference(args, device, steps, warmup, budget_s):
    """The UNMODIFIED reference (oracle/_ref, vendored by oracle/build_ref.py) through its own public API and stock code path:
    SATrans(...).compile(torch.optim.Adam(...)).fit(x, y, batch_size=B, verbose=0) -- DataLoader, train-mode forward, BCE(sum),
    dense L2, autograd backward, dense Adam (meta_basemodel.py:200-385).  One fit() call of `warmup` batches, then one timed
    fit() call of `steps` batches; the per-step time is that call's wall time / steps (the whole call is the user-visible cost)."""
    import contextlib
    import io
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import build_ref
    SATransRef, SparseFeat, _ = build_ref.import_reference()
    from satrans_b200 import synth
    cuda = str(device).startswith('cuda')
    cores = os.cpu_count() or 1
    if not cuda:
        torch.set_num_threads(cores)
    schema = synth.aliccp_schema(args.vocab_scale)
    cols = [SparseFeat(f, vocabulary_size=v, embedding_dim=32) for f, v in zip(schema.fields, schema.vocab_sizes)]
    quiet = io.StringIO()
    with contextlib.redirect_stdout(quiet):
        model = SATransRef(linear_feature_columns=cols, dnn_feature_columns=cols, domain_column_list=['301'],
                           num_domains_list=[3], att_layer_num=0, domain_att_layer_num=3, att_head_num=4, meta_mode='QK',
                           use_linear=False, use_dnn=False, seed='1021', device=device, flag='sota')
        model.compile(torch.optim.Adam(model.parameters(), lr=0.005), "binary_crossentropy",
                      metrics=["binary_crossentropy", "auc"])
    B = args.batch

    def run(n_steps, seed):
        xs, ys = synth.make_inputs(schema, n_steps * B, seed=seed, dist=args.ids, ctr=0.04)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(quiet):
            model.fit(xs, ys, batch_size=B, epochs=1, verbose=0)
        if cuda:
            torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n_steps

    first = run(1, 1)
    n_warm = max(0, min(warmup, int(budget_s * 0.25 / max(first, 1e-3))) - 1)
    if n_warm:
        run(n_warm, 2)
    n = max(1, min(steps, int(budget_s * 0.75 / max(first, 1e-3))))
    dt = run(n, 3)
    where = 'cuda:0 (eager PyTorch, ~4.7 k kernel launches per step)' if cuda else 'host CPUs, torch %d threads' % cores
    return {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "one fit() call of %d batches (B=%d, %s vocab: train-mode forward + BCE + dense L2 + backward + dense "
                      "torch.optim.Adam through the unmodified reference's SATrans.fit, verbose=0) after %d warm-up batches, %s"
                      % (n, B, 'full' if args.vocab_scale == 1.0 else 'x%g' % args.vocab_scale, n_warm + 1, where),
            "ms_per_step": dt * 1e3, "steps": n, "warmup": n_warm + 1}


def time_port(args, device, steps, warmup, budget_s):
    """Fallback where oracle/_ref is absent: the oracle's restatement of the same step (oracle/satrans_oracle.Trainer)."""
    import torch
    cores = os.cpu_count() or 1
    cuda = str(device).startswith('cuda')
    if not cuda:
        torch.set_num_threads(cores)
    O, schema, cfg, X, y = oracle_setup(args, 4)
    st = O.init_state(cfg, seed=1021)
    if cuda:
        st = {k: v.to(device) for k, v in st.items()}
    tr = O.Trainer(st, cfg, lr=0.005)
    B = args.batch
    Xt, yt = torch.from_numpy(X).to(device), torch.from_numpy(y).to(device)

    def step(i):
        k = i % 4
        r = tr.step(Xt[k * B:(k + 1) * B], yt[k * B:(k + 1) * B], train_mode=True)
        if cuda:
            torch.cuda.synchronize()
        return r

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
            "sample": "%d timed full steps (B=%d, train-mode forward + BCE + dense L2 + backward + dense torch.optim.Adam) of "
                      "oracle/satrans_oracle.Trainer after %d warm-up on %s" % (n, B, n_warm + 1, device),
            "ms_per_step": dt * 1e3, "steps": n, "warmup": n_warm + 1}


def time_cpu_reference(args, steps, warmup, budget_s, device='cpu'):
    if reference_available():
        try:
            return time_reference(args, device, steps, warmup, budget_s)
        except Exception as e:          # e.g. a missing third-party name on this box: report the port instead, say why
            r = time_port(args, device, steps, warmup, budget_s)
            r["sample"] += " (oracle/_ref failed to run: %r)" % (e,)
            return r
    return time_port(args, device, steps, warmup, budget_s)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = time_cpu_reference(args, args.steps, max(args.warmup, 1), budget_s=150.0)
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
    """SM clock + throttle reasons sampled DURING the timed region through NVML (a polling thread, ~2 ms period: the timed
    region of the default driver invocation is ~50 ms, `nvidia-smi -lms` would see one sample of it).  mark() brackets the
    region(s) whose samples count."""

    def __init__(self, gpu_index, period_s=0.002):
        import threading
        self.samples = []            # (t, sm_mhz, reasons bitmask)
        self.windows = []
        self._stop = False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.period = period_s
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml polling thread"}
        if not self.ok:
            return out
        self._stop = True
        self.th.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 "hw_thermal_slowdown": getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 "sw_thermal_slowdown": getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 "sw_power_cap": getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        sel = [x for x in self.samples if any(a <= x[0] <= b for a, b in self.windows)] if self.windows else self.samples
        if sel:
            mask = 0
            for _, _, r in sel:
                mask |= r
            out.update(sm_mhz=float(np.median([x[1] for x in sel])), samples=len(sel),
                       reasons=sorted(k for k, bit in names.items() if mask & bit))
        return out


# ------------------------------------------------------------------------------------------------------------
# this framework
# ------------------------------------------------------------------------------------------------------------
def _time_loop(fn, n, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _time_graph(fn, n, torch):
    """ms per call of n calls captured in ONE CUDA graph (kernels back to back on the device: for launches shorter than the host
    needs to enqueue them)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(2):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


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
    traffic_tab = {}
    try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of this build (profiles/)
        traffic_tab = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
    except Exception:
        pass

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
    n_batches = 16 if args.config == 2 else 4
    xs, ys = synth.make_inputs(schema, n_batches * B, seed=1021 + rank, dist=args.ids, ctr=0.04)
    if args.config == 4 or max(schema.vocab_sizes) >= 2 ** 24:
        Xnp = np.stack([np.asarray(xs[k]) for k in schema.fields], axis=1).astype(np.int32)   # ids >= 2^24 cannot travel as float32
    else:
        Xnp = synth.stack_columns(xs, schema.fields)                       # the reference's float32 input matrix
    X_host = torch.from_numpy(Xnp).pin_memory()
    y_host = torch.from_numpy(ys.astype(np.float32)).pin_memory()
    Xd, yd = X_host.to(dev), y_host.to(dev)
    hp = model._opt_hparams
    model.train()

    # (row-sharded steps are launched eagerly: capturing them -- peer-barrier kernels + the NCCL all-reduce -- measured 1.4 % faster
    #  at 2 GPUs but the processes did not exit cleanly afterwards)
    use_graph = world == 1 and not args.no_graph and not args.overlap
    defer = args.sweep_interval if (not args.lazy and not args.overlap) else 1
    model.table_sweep_interval = defer

    def step(i, overlap=args.overlap, eager=False):
        """One optimiser step on a device-resident batch: the engine call fit() makes per batch -- the CUDA-graph replay of the
        whole step (single GPU; the first steps of a shape run eagerly), or the eager launch sequence."""
        k = i % n_batches
        xb, yb = Xd[k * B:(k + 1) * B], yd[k * B:(k + 1) * B]
        if use_graph and not eager and eng.step_count >= 2:
            w, _ = eng.graph_train_step(xb, yb, lr=hp['lr'], betas=hp['betas'], eps=hp['eps'], opt_kind=hp['kind'], l2_emb=1e-5,
                                        dense_mode=not args.lazy, dropout_p=0.1, seed0=1021, defer=defer)
            return eng.last_y_pred, w
        return eng.train_step(xb, yb, lr=hp['lr'], betas=hp['betas'], eps=hp['eps'],
                              opt_kind=hp['kind'], l2_emb=1e-5, dense_mode=not args.lazy, training=True, dropout_p=0.1, seed=1021 + i,
                              overlap=overlap, defer=defer)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for i in range(W):
        step(i)
    barrier()
    # ---- phase breakdown (CUDA events around every phase of the EAGER launch sequence; not part of `value`) -----------------
    eng.timing = {}
    for i in range(min(K, 10) if defer <= 1 else 2 * defer):
        _, w = step(W + i, eager=True)
    eng.sync_tables()
    phases = eng.timing_ms()
    eng.timing = None
    barrier()
    # ---- timed region: K steps, device-resident inputs ------------------------------------------------------
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    ev0.record()
    for i in range(K):
        _, w = step(W + i)
    eng.sync_tables()                      # (the last sweep epoch closes inside the timed region: every row is up to date)
    ev1.record()
    barrier()
    tw1 = time.perf_counter()
    if clocks:
        clocks.mark(tw0, tw1)
    ms = ev0.elapsed_time(ev1) / K
    launches = eng.launches - l0
    n_unique = float(w['scalars'][2].item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world / (ms * 1e-3)

    # ---- e2e: the reference's public training call, fit(), on HOST data (pinned): every step's batch crosses PCIe inside the
    #      timed region (copy stream, one step ahead), every step's loss scalars are copied back to pinned host memory by the
    #      last node of the step's graph, and fit() returns after the tables are in sync and the epoch loss is on the host -------
    if world == 1:
        reps = -(-K // n_batches)
        Xe = X_host.repeat(reps, 1)[:K * B].contiguous().pin_memory()
        ye = y_host.repeat(reps)[:K * B].contiguous().pin_memory()
        model.stream_host_batches = True
        model.fit(Xe[:4 * B], ye[:4 * B], batch_size=B, epochs=1, shuffle=False, verbose=0)      # warm-up (graph capture)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        hist = model.fit(Xe, ye, batch_size=B, epochs=1, shuffle=False, verbose=0)
        e1.record()
        barrier()
        tw1 = time.perf_counter()
        e2e_api = ("SATrans.fit(x_host, y_host, batch_size=%d, epochs=1, shuffle=False, verbose=0) on %d samples in pinned host "
                   "memory -> History (epoch loss %.6f)" % (B, K * B, hist.history['loss'][-1]))
        e2e_wall_ms = (tw1 - tw0) * 1e3 / K
    else:
        for i in range(3):
            model.train_on_batch(X_host[:B], y_host[:B], defer_reg=defer > 1)
        model.sync_tables()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        for i in range(K):
            k = (i + 3) % n_batches
            model.train_on_batch(X_host[k * B:(k + 1) * B], y_host[k * B:(k + 1) * B], defer_reg=defer > 1)
        model.sync_tables()
        e1.record()
        barrier()
        tw1 = time.perf_counter()
        e2e_api = ("SATrans.train_on_batch(x_host, y_host, defer_reg=True) -> (bce, None) per step (synchronous: the step's bce is read "
                   "back) + SATrans.sync_tables() at the end, inside the timed region") if defer > 1 else \
            "SATrans.train_on_batch(x_host, y_host) -> (bce, reg)"
        e2e_wall_ms = (tw1 - tw0) * 1e3 / K
    if clocks:
        clocks.mark(tw0, tw1)
    clk = clocks.stop() if clocks else {}
    e2e_ms = e0.elapsed_time(e1) / K
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e = {"value": B * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms,
           "h2d_bytes_per_step": int(B * X_host.shape[1] * 4 + B * 4), "d2h_bytes_per_step": 16,
           "api": e2e_api}
    if rank != 0:
        if world > 1:
            dist.barrier()          # keep this rank's table shard mapped until rank 0 is done with the extras
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the dense L2+Adam sweep (HBM-bound) ---------------------------------
    rows = model._tables.shape[0]                                    # rows of this rank's shard
    sweep_bytes = (rows - n_unique) * 6 * 128 + (rows / 8)           # p,m,v read + written once; bitmap read
    roof = None
    full_cfg2 = world == 1 and args.config == 2 and args.vocab_scale == 1.0

    def traffic_of(key):
        tr = traffic_tab.get(key) if full_cfg2 else None
        return (tr['dram_bytes_read'] + tr['dram_bytes_write']) if tr else None

    if 'table_sweep' in phases:
        sweep_bytes = (rows - n_unique) * 6 * 128 + 2 * rows       # p, m, v read + written once per EPOCH; the per-row step byte
        ach = sweep_bytes / (phases['table_sweep'] * 1e-3) / 1e9
        roof = {"kernel": "k_defer_sweep", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic_of('k_defer_sweep'), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": phases['table_sweep'],
                "launches_per_step": 1.0 / defer,
                "note": "every-row table update (L2 + Adam of the rows no batch read) once per %d steps: 6 x 128 B per row per "
                        "LAUNCH (p, m, v read once + written once, each row taken through the epoch's steps in registers) + 2 B "
                        "per row of step bookkeeping; CUDA events around the launch (sweep + its 1-CTA finalize) on the main "
                        "stream" % defer}
    elif 'adam_sweep' in phases:
        ach = sweep_bytes / (phases['adam_sweep'] * 1e-3) / 1e9
        roof = {"kernel": "k_adam_sweep", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic_of('k_adam_sweep'), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sweep_bytes, "ms_per_launch": phases['adam_sweep'],
                "note": "6 x 128 B per untouched table row; timed with CUDA events on its own stream while "
                        "forward/backward run concurrently on the main stream" if args.overlap else
                        "6 x 128 B per untouched table row; CUDA events around its launch on the main stream (the ~20 us "
                        "touched-row and dense-parameter updates run beside it on a second stream)"}

    # ---- extra: stand-alone gather GB/s (second roofline entry) and eval-mode scoring ---------------------------
    extra = {}
    try:
        others = []
        per_sample = 2 * 19 * 128 + 19 * 4
        Bg_max = 262144
        xg, _ = synth.make_inputs(schema, Bg_max, seed=7, dist=args.ids)
        Xg = torch.from_numpy(np.stack([np.asarray(xg[k]) for k in schema.fields], axis=1).astype(np.int32)).to(dev)
        # rotating output / input slices so that consecutive launches do not find their rows in L2 (B=8192: 20 MB per launch)
        for Bg, reps in ((8192, 32), (Bg_max, 10)):
            n_slices = max(1, Bg_max // Bg)
            slices = [Xg[i * Bg:(i + 1) * Bg] for i in range(n_slices)]
            for i in range(3):
                eng.gather(slices[i % n_slices])
            gms = _time_graph(lambda i: eng.gather(slices[i % n_slices]), reps, torch)
            gb = Bg * per_sample / (gms * 1e-3) / 1e9
            others.append({"kernel": "k_gather", "bound": "hbm", "batch": Bg, "achieved": gb, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gb / hbm_peak, "traffic": traffic_of('k_gather_B%d' % Bg), "ms_per_launch": gms,
                           "algorithmic_bytes_per_launch": Bg * per_sample,
                           "note": "4,940 B per sample: 19 x 128 B rows read + written, 19 int32 ids; a third of the row reads "
                                   "(the tiny tables) are served by L1/L2, so DRAM traffic is below the algorithmic bytes"})
        if roof is not None:
            roof["others"] = others
            try:        # the three token kernels (85 % of the step) are latency-bound: their pipe utilisation from the ncu captures
                roof["token_kernels"] = dict(json.load(open(os.path.join(ROOT, 'profiles', 'r2_pipe_utilisation.json'))),
                                             note="static: ncu --set full captures of this build (profiles/r2_details_k_*.txt), not re-measured by this run")
            except Exception:
                pass
        extra["gather"] = others[-1]
        model.eval()
        Bi = 32768
        Xi = Xg[:Bi].contiguous()

        def infer(i):
            eng.forward(Xi, training=False)
        for _ in range(3):
            infer(0)
        ims = _time_loop(infer, 10, torch)
        extra["infer"] = {"batch": Bi, "ms": ims, "value": Bi / (ims * 1e-3), "unit": UNIT,
                          "gflops": Bi * 1540672 / (ims * 1e-3) / 1e9, "precision": args.precision}
        for prec in ('fp32', 'tf32x3', 'tf32'):
            if prec == args.precision:
                continue
            eng.precision = prec
            for _ in range(3):
                infer(0)
            pms = _time_loop(infer, 10, torch)
            extra["infer_" + prec] = {"batch": Bi, "ms": pms, "value": Bi / (pms * 1e-3), "unit": UNIT}
        eng.precision = args.precision
        del Xg
    except Exception as e:  # the headline number must survive a failure of the extras
        extra["extra_error"] = repr(e)

    # ---- the bf16 variant (kind::f16 MMAs on bf16 operands, fp32 accumulation): reported separately, never the headline ------
    if world == 1:
        try:
            eng.precision = 'bf16'
            model.train()
            for i in range(3):
                step(i)
            bms = _time_loop(lambda i: step(1000 + i), min(K, 20), torch)
            model.eval()
            for _ in range(3):
                eng.forward(Xd[:B], training=False)
            bi = _time_loop(lambda i: eng.forward(Xd[:B], training=False), 10, torch)
            extra["bf16"] = {"train_samples_per_sec": B / (bms * 1e-3), "ms_per_step": bms, "infer_samples_per_sec": B / (bi * 1e-3),
                             "note": "projection GEMMs as tcgen05 kind::f16 on bf16 operands, fp32 accumulate; tolerance vs the reference: "
                                     "y_pred atol 2e-2, loss rtol 2e-2, gradient cosine >= 0.98 (tests/test_gpu_parity.py)"}
        except Exception as e:
            extra["bf16"] = {"error": repr(e)}
        eng.precision = args.precision
        model.train()

    cpu = None
    eager = None
    if world == 1 and args.config == 2:
        if not args.no_eager_baseline:
            try:        # the reference's eager PyTorch path on this same B200
                r = time_cpu_reference(args, 10, 3, budget_s=20.0, device=dev)
                eager = {k: r[k] for k in ("value", "unit", "kind", "sample", "ms_per_step")}
            except Exception as e:
                eager = {"error": repr(e)}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            r = time_cpu_reference(args, 3, 1, budget_s=args.cpu_budget_s)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clk, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "eager_gpu_baseline": eager, "phases_ms": phases}
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
