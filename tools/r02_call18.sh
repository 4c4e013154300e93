#!/usr/bin/env bash
# r02 call 18: kernel choice after the one-role flush change (small catalogs, chain batch, 16/32/64-way shard sizes) and the
# one-role slice parameters (LMIN, GUIDED_DIV) on the 8-way shard and the full catalog
set -u
OUT=gpurun_out
TAG=r02c18
mkdir -p $OUT
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chunks', d['plan']['n_chunks'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$1 FAILED', e)"; }
for k in 1 0; do
  GWI_CTA_KERNEL=$k python bench.py --workload cfg2 --steps 500 --warmup 20 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg2 cta=$k" | tee -a $OUT/${TAG}_lines.txt
  GWI_CTA_KERNEL=$k python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | line "cfg4 cta=$k" | tee -a $OUT/${TAG}_lines.txt
  for w in 16 32 64; do
    GWI_CTA_KERNEL=$k python bench.py --workload cfg3 --emulate-world $w --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | line "shard$w cta=$k" | tee -a $OUT/${TAG}_lines.txt
  done
done
for lm in 16 32 64; do for gd in 1 2; do
  GWI_CTA_KERNEL=0 GWI_TUNE_LMIN=$lm GWI_TUNE_GUIDED_DIV=$gd python bench.py --workload cfg3 --emulate-world 8 --steps 100 --no-cpu-baseline 2>/dev/null | line "shard8 one-role lmin=$lm gdiv=$gd" | tee -a $OUT/${TAG}_lines.txt
done; done
for lm in 16 64; do for gd in 1 2; do
  GWI_TUNE_LMIN=$lm GWI_TUNE_GUIDED_DIV=$gd python bench.py --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg3 lmin=$lm gdiv=$gd" | tee -a $OUT/${TAG}_lines.txt
done; done
