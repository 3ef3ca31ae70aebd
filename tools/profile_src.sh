#!/bin/bash
# Source-level hot spots of the token kernels (run ON the GPU box from the repo root):  tools/profile_src.sh <tag> [kernels...]
# Writes gpurun_out/<tag>_src_<kernel>.csv.gz (ncu --page source --csv, CUDA + SASS) and <tag>_stalls_<kernel>.txt (warp-state metrics).
set -u
TAG=${1:-src}; shift
KERNELS=${@:-k_forward_tc k_bwd_meta_tc k_bwd_attn}
OUT=gpurun_out; mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph"
for K in $KERNELS; do
  timeout 600 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section InstructionStats --import-source on --clock-control none \
      -k regex:$K -s 4 -c 1 -f -o $OUT/${TAG}_$K $BENCH > $OUT/${TAG}_$K.log 2>&1
  if [ -f $OUT/${TAG}_$K.ncu-rep ]; then
    ncu -i $OUT/${TAG}_$K.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $OUT/${TAG}_src_$K.csv.gz
    ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>=3:
    for h,u,v in zip(rows[0],rows[1],rows[2]):
        if any(s in h for s in ('issue_stalled','warps_','inst_executed','gpu__time','pipe_','issue_active','registers','thread_inst')):
            print('%-100s %s %s'%(h,v,u))
" > $OUT/${TAG}_stalls_$K.txt
    rm -f $OUT/${TAG}_$K.ncu-rep
  fi
done
ls -la $OUT | grep ${TAG}_
