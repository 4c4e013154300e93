#!/usr/bin/env bash
# r02 call 17: (1) one slice per CTA on small catalogs (cfg2) on / off, (2) kernel choice re-measured after the one-role
# flush change: 8-way / 4-way cfg3 shards and cfg5 with either kernel, (3) adaptive prologue walk: phases
set -u
OUT=gpurun_out
TAG=r02c17
mkdir -p $OUT
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chunks', d['plan']['n_chunks'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$1 FAILED', e)"; }
for o in 1 0; do
  GWI_TUNE_ONE_SLICE=$o python bench.py --workload cfg2 --steps 500 --warmup 20 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg2 one_slice=$o" | tee -a $OUT/${TAG}_lines.txt
done
python bench.py --workload cfg2 --steps 200 --warmup 20 --flush-l2 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg2 cold" | tee -a $OUT/${TAG}_lines.txt
for k in 1 0; do
  GWI_CTA_KERNEL=$k python bench.py --workload cfg3 --emulate-world 8 --steps 100 --no-cpu-baseline 2>/dev/null | line "shard8 cta=$k" | tee -a $OUT/${TAG}_lines.txt
  GWI_CTA_KERNEL=$k python bench.py --workload cfg3 --emulate-world 4 --steps 60 --no-cpu-baseline 2>/dev/null | line "shard4 cta=$k" | tee -a $OUT/${TAG}_lines.txt
  GWI_CTA_KERNEL=$k python bench.py --workload cfg5 --steps 50 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg5 cta=$k" | tee -a $OUT/${TAG}_lines.txt
done
for w in "cfg3" "cfg2 --steps 300 --warmup 20"; do
  echo "== $w" >> $OUT/${TAG}_phases.txt
  GWI_PHASE_TIMING=1 python bench.py --no-cpu-baseline --no-nuts --workload $w 2>&1 >/dev/null | grep "gwi phases" >> $OUT/${TAG}_phases.txt
done
cat $OUT/${TAG}_phases.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee $OUT/${TAG}_pytest.txt
