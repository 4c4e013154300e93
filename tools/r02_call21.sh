#!/usr/bin/env bash
# device plan builder: parity tests, then build times (device vs host builder) on cfg3 / cfg5 / cfg2
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -x -k "device_plan" 2>&1 | tail -15 | tee $OUT/r02c21_pytest_device_plan.txt
python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/r02c21_pytest_gpu.txt
for w in cfg3 cfg5 cfg2; do
  for dev in 1 0; do
    echo "== $w GWI_PLAN_DEVICE=$dev" | tee -a $OUT/r02c21_plan_timing.txt
    GWI_PLAN_DEVICE=$dev GWI_PLAN_TIMING=1 python bench.py --workload $w --steps 30 --no-cpu-baseline 2> $OUT/r02c21_err_${w}_$dev.txt > $OUT/r02c21_bench_${w}_dev$dev.json
    grep "gwi plan" $OUT/r02c21_err_${w}_$dev.txt | tee -a $OUT/r02c21_plan_timing.txt
    python - $OUT/r02c21_bench_${w}_dev$dev.json <<'PY' | tee -a $OUT/r02c21_plan_timing.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "ms/step kernel", round(d["roofline"]["kernel_ms"], 4), "setup", d["setup_s"], "log_l", d["result"])
except Exception as e:
    print("FAILED", e)
PY
  done
done
