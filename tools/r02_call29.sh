#!/usr/bin/env bash
# one-role kernel on an 8-way shard (and cfg3): slices per warp / minimum slice / guided divisor
set -u
OUT=gpurun_out/r02c29_slices.txt
: > $OUT
run() {  # label, env...
  local label="$1"; shift
  for w in "--emulate-world 8 --steps 150" "--steps 40"; do
    env "$@" python bench.py --workload cfg3 $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label | $w |', round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'chunks', d['plan']['n_chunks'], 'clk', d['clocks']['sm_mhz'])" | tee -a $OUT
  done
}
run "default (spw 4, lmin 32, div 1)" GWI_X=0
run "spw 2" GWI_TUNE_SLICES_PER_WARP=2
run "spw 3" GWI_TUNE_SLICES_PER_WARP=3
run "spw 6" GWI_TUNE_SLICES_PER_WARP=6
run "spw 3 lmin 64" GWI_TUNE_SLICES_PER_WARP=3 GWI_TUNE_LMIN=64
run "spw 4 lmin 16" GWI_TUNE_LMIN=16
run "spw 4 div 2" GWI_TUNE_GUIDED_DIV=2
run "spw 2 div 2 lmin 16" GWI_TUNE_SLICES_PER_WARP=2 GWI_TUNE_GUIDED_DIV=2 GWI_TUNE_LMIN=16
