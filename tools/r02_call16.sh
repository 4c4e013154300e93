#!/usr/bin/env bash
# r02 call 16: (1) one-role flush via red + 4-at-a-time prologue walk: cfg3 / shard8 / cfg2 lines and phases,
# (2) GWI_TUNE_SLICE_DIV sweep on the CTA-kernel workloads, (3) full capture of the CTA kernel at cfg2 (fixed overhead?)
set -u
OUT=gpurun_out
TAG=r02c16
mkdir -p $OUT
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chunks', d['plan']['n_chunks'])
except Exception as e: print('$1 FAILED', e)"; }
python bench.py --no-cpu-baseline --no-nuts 2>/dev/null | tee $OUT/${TAG}_cfg3.json | line cfg3 | tee -a $OUT/${TAG}_lines.txt
for d in 2 3 4; do
  GWI_TUNE_SLICE_DIV=$d python bench.py --workload cfg3 --emulate-world 8 --steps 100 --no-cpu-baseline 2>/dev/null | line "shard8 div=$d" | tee -a $OUT/${TAG}_lines.txt
  GWI_TUNE_SLICE_DIV=$d python bench.py --workload cfg2 --steps 500 --warmup 20 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg2 div=$d" | tee -a $OUT/${TAG}_lines.txt
  GWI_TUNE_SLICE_DIV=$d python bench.py --workload cfg5 --steps 50 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg5 div=$d" | tee -a $OUT/${TAG}_lines.txt
  GWI_TUNE_SLICE_DIV=$d python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | line "cfg4 div=$d" | tee -a $OUT/${TAG}_lines.txt
done
for w in "cfg3" "cfg2 --steps 300 --warmup 20"; do
  echo "== $w" >> $OUT/${TAG}_phases.txt
  GWI_PHASE_TIMING=1 python bench.py --no-cpu-baseline --no-nuts --workload $w 2>&1 >/dev/null | grep "gwi phases" >> $OUT/${TAG}_phases.txt
done
cat $OUT/${TAG}_phases.txt
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:stream_cta_kernel -s 20 -c 1 -f -o $OUT/${TAG}_cfg2_stream_cta_kernel \
    python bench.py --workload cfg2 --steps 30 --warmup 10 --no-cpu-baseline --no-nuts > /dev/null 2> $OUT/${TAG}_ncu_err.txt
tail -2 $OUT/${TAG}_ncu_err.txt
ls -la $OUT | grep $TAG
