#!/bin/bash
# 2+ GPU evidence (run ON the box from the repo root): tools/gpu_multi.sh <tag> <ngpus> [configs]
TAG=${1:-r2m}; N=${2:-2}; CFGS=${3:-"2 4"}
[ -n "${SKIP_DIST_TEST:-}" ] || timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/${TAG}_dist_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_dist_pytest.log
for CFG in $CFGS; do
NAME=${TAG}_bench_${N}gpu_cfg${CFG}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 48 --warmup 5 --config $CFG --no-cpu-baseline --no-eager-baseline > gpurun_out/$NAME.json 2> gpurun_out/$NAME.err
python -c "
import json;d=json.loads(open('gpurun_out/$NAME.json').readline());print('$NAME', d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['phases_ms'])"; tail -2 gpurun_out/$NAME.err | cut -c1-300
done
