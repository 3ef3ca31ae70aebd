#!/bin/bash
# One gpurun call worth of evidence for profiles/ (run ON the GPU box from the repo root):
#   tools/profile_r2.sh <tag>
# Writes gpurun_out/<tag>_*: ncu launch list of a short bench run, `ncu --set full` details + raw metrics of one launch of each
# dominant kernel, SASS opcode histograms per kernel.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/sat_rawparse.py <<'PY'
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next((i for i, r in enumerate(rows) if 'Kernel Name' in r), None)
if start is not None and len(rows) >= start + 3:
    hdr, units, vals = rows[start], rows[start + 1], rows[start + 2]
    keep = ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active', 'launch__grid_size',
            'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_inst0.avg.pct_of_peak_sustained_active')
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals) if h in keep or 'pipe_tensor' in h or 'issue_stalled' in h or h.startswith('dram__bytes')}
    json.dump(d, open(sys.argv[2], 'w'), indent=1)
PY
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph"
# 1. launch list (every launch with its device time; cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
# 2. one full capture per dominant kernel
for K in k_forward_tc k_bwd_meta_tc k_bwd_attn k_defer_sweep k_defer_catchup k_gather; do
  SKIP=4; case $K in k_defer_*) SKIP=1 ;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/${TAG}_$K $BENCH > $OUT/${TAG}_$K.log 2>&1
  if [ -f $OUT/${TAG}_$K.ncu-rep ]; then
    ncu -i $OUT/${TAG}_$K.ncu-rep --page details > $OUT/${TAG}_details_$K.txt 2>/dev/null
    ncu -i $OUT/${TAG}_$K.ncu-rep --page raw --csv 2>/dev/null > $OUT/${TAG}_raw_$K.csv
    python /tmp/sat_rawparse.py $OUT/${TAG}_raw_$K.csv $OUT/${TAG}_raw_$K.json; rm -f $OUT/${TAG}_raw_$K.csv
    rm -f $OUT/${TAG}_$K.ncu-rep      # (tens of MB each; the text / json pages are what gets committed)
  fi
done
# 3. SASS opcode histogram per kernel
python - <<'PY' > $OUT/${TAG}_sass_histogram.txt 2>&1
import collections, re, subprocess
out = subprocess.run(['cuobjdump', '-sass', 'satrans_b200/lib/libsatrans_b200.so'], capture_output=True, text=True).stdout
cur, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', ln)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', ln)
    if m and cur:
        hist[cur][m.group(1).split('.')[0]] += 1
for k, h in hist.items():
    if sum(h.values()) < 200:
        continue
    key = [o for o in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'LDGSTS', 'FFMA2', 'FFMA', 'LDS', 'STS', 'LDG', 'STG', 'MUFU', 'BAR', 'ATOMG', 'REDG', 'ATOMS') if h.get(o)]
    print('%s\n   total %d | %s' % (k, sum(h.values()), '  '.join('%s %d' % (o, h[o]) for o in key)))
PY
ls -la $OUT | tail -30
