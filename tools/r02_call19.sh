#!/usr/bin/env bash
# r02 call 19: multi-chain NUTS (batched evaluations) on cfg2 + the new kernel-choice threshold (shard8 / cfg5 / cfg3 lines)
set -u
OUT=gpurun_out
TAG=r02c19
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -q -x -k "chains_advance or native_nuts" 2>&1 | tail -2 | tee $OUT/${TAG}_pytest.txt
timeout 900 python tools/chains_probe.py --nuts 8,16 2>&1 | tee $OUT/${TAG}_chains_probe.txt
line() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/step kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'chunks', d['plan']['n_chunks'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$1 FAILED', e)"; }
python bench.py --workload cfg3 --emulate-world 8 --steps 100 --no-cpu-baseline 2>/dev/null | line "shard8" | tee -a $OUT/${TAG}_lines.txt
python bench.py --workload cfg5 --steps 50 --no-cpu-baseline --no-nuts 2>/dev/null | line "cfg5" | tee -a $OUT/${TAG}_lines.txt
