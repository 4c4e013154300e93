#!/usr/bin/env bash
# speculative shift (build -DGWI_EXP_TRACK_MAX=1, GWI_SPECULATIVE_SHIFT=1) on cfg1: the host call's e2e with one pass instead of two
set -u
OUT=gpurun_out/r02c37_speculative_shift.txt
: > $OUT
run() {
  local label="$1"; shift
  env "$@" python bench.py --workload cfg1 --steps 400 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label | device loop', round(d['ms_per_step']*1e3,1), 'us/step | e2e host call', round(1e6/d['e2e']['value'],1), 'us/eval | e2e_python', round(1e6/d['e2e_python']['value'],1) if d.get('e2e_python') else None, '| log_l', d['result']['log_l'], 'launches/eval', d['gpu_launches']//d['steps'])" | tee -a $OUT
}
run "product build" GWI_X=0
run "product build (again)" GWI_X=0
run "track-max build, speculation off" GWI_LIBRARY=gwinferno_b200/libgwi_spec.so
run "track-max build, GWI_SPECULATIVE_SHIFT=1" GWI_LIBRARY=gwinferno_b200/libgwi_spec.so GWI_SPECULATIVE_SHIFT=1
run "track-max build, GWI_SPECULATIVE_SHIFT=1 (again)" GWI_LIBRARY=gwinferno_b200/libgwi_spec.so GWI_SPECULATIVE_SHIFT=1
