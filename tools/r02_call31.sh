#!/usr/bin/env bash
# reduction fan-in: one level of 128 / 256 inputs per task against two levels of 64 (cfg3, 8-way shard, cfg5)
set -u
OUT=gpurun_out/r02c31_reduce_fan.txt
: > $OUT
for fan in auto 64 128 256; do
  for w in "cfg3 --steps 40" "cfg3 --emulate-world 8 --steps 150" "cfg5 --steps 60"; do
    if [ "$fan" = auto ]; then E=GWI_X=0; else E=GWI_TUNE_REDUCE_FAN=$fan; fi
    env $E GWI_PHASE_TIMING=0 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fan $fan | $w |', round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'tail', round(d['ms_per_step']-d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches']//d['steps'], 'clk', d['clocks']['sm_mhz'], 'log_l', d['result']['log_l'])" | tee -a $OUT
  done
done
python -m pytest tests -m gpu -q -k "full_size or medium or golden_values" 2>&1 | tail -2 | tee -a $OUT
