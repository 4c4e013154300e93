#!/usr/bin/env bash
# fused second reduction level A/B (cfg3, 8-way shard, cfg5) + the new GPU tests
set -u
OUT=gpurun_out/r02c32_fuse_reduce.txt
: > $OUT
python -m pytest tests -m gpu -q -k "fused_second or full_size or run_to_run" 2>&1 | tail -2 | tee -a $OUT
for rep in 1 2; do
for fuse in 1 0; do
  for w in "cfg3 --steps 60" "cfg3 --emulate-world 8 --steps 200" "cfg5 --steps 80"; do
    GWI_FUSE_REDUCE=$fuse python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse $fuse | $w |', round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'tail', round(d['ms_per_step']-d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches']//d['steps'], 'clk', d['clocks']['sm_mhz'], 'log_l', d['result']['log_l'])" | tee -a $OUT
  done
done
done
