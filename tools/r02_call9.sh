#!/usr/bin/env bash
set -u
OUT=gpurun_out
TAG=r02c9
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $OUT/${TAG}_pytest_gpu.txt 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/${TAG}_smoke.txt
for w in "cfg3" "cfg3 --emulate-world 8" "cfg2 --steps 300 --warmup 20" "cfg5"; do
  echo "== $w" >> $OUT/${TAG}_phases.txt
  GWI_PHASE_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --workload $w 2>&1 >/dev/null | grep "gwi phases" >> $OUT/${TAG}_phases.txt
done
cat $OUT/${TAG}_phases.txt
timeout 900 python bench.py > $OUT/${TAG}_bench_cfg3_n1.json 2> $OUT/${TAG}_bench_err.txt
tail -c 600 $OUT/${TAG}_bench_err.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c9_bench_cfg3_n1.json"))
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "e2e_python", "parity_at_size", "cpu_baseline", "nuts", "clocks", "gpu_launches")}, default=str)[:3000])
print("roofline", d["roofline"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench_err.txt
head -c 1500 $OUT/${TAG}_bench_reference.json
