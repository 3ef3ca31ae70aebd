#!/bin/bash
# Everything one GPU lease should produce, each step under its own timeout (run ON the box from the repo root):
#   tools/gpu_batch.sh <tag> [steps...]     steps: test testall smoke bench benchlong gather profile launches sweepvar configs (default: test smoke bench gather profile)
TAG=${1:-r2}; shift
STEPS=${@:-test smoke bench gather profile}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
for S in $STEPS; do
  case $S in
    test)   timeout 1200 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider -x --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log ;;
    testall) timeout 1500 python -m pytest tests -m gpu -q --no-header -rf -p no:cacheprovider --timeout 600 > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log ;;
    smoke)  timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/${TAG}_smoke.log ;;
    bench)  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" >> $OUT/${TAG}_bench.err ;;
    benchlong) timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err ;;
    gather) timeout 600 python tools/gather_sweep.py > $OUT/${TAG}_gather.jsonl 2>&1 ;;
    profile) timeout 1500 tools/profile_r2.sh $TAG > $OUT/${TAG}_profile.log 2>&1 ;;
    launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > $OUT/${TAG}_launches.log 2>&1 ;;
    sweepvar) for BPS in 8 32 128 256; do SAT_SWEEP_BPS=$BPS timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(json.dumps({'bps': $BPS, 'ms': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'sweep_frac': d['roofline']['frac'], 'phases': d['phases_ms']}))" >> $OUT/${TAG}_sweepvar.jsonl; done ;;
    configs) timeout 900 python bench_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err ;;
  esac
done
echo "== pytest"; tail -15 $OUT/${TAG}_pytest.log 2>/dev/null
echo "== smoke"; tail -3 $OUT/${TAG}_smoke.log 2>/dev/null
echo "== bench"; cut -c1-1500 $OUT/${TAG}_bench.json 2>/dev/null; tail -5 $OUT/${TAG}_bench.err 2>/dev/null
cut -c1-1500 $OUT/${TAG}_bench_default.json 2>/dev/null; tail -3 $OUT/${TAG}_bench_default.err 2>/dev/null
[ -f $OUT/${TAG}_launches.csv ] && (grep -c "k_" $OUT/${TAG}_launches.csv; tail -2 $OUT/${TAG}_launches.log)
echo "== gather"; cat $OUT/${TAG}_gather.jsonl 2>/dev/null | cut -c1-200
echo "== sweepvar"; cat $OUT/${TAG}_sweepvar.jsonl 2>/dev/null | cut -c1-400
