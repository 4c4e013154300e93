#!/usr/bin/env bash
# after the device plan builder, explicit knots, PPD grids: full GPU suite, smoke, headline bench (device-built plan), 8-way shard
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $OUT/r02c22_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/r02c22_smoke.txt
python bench.py --steps 100 --no-nuts > $OUT/r02c22_bench_cfg3.json 2> $OUT/r02c22_bench_err.txt
python bench.py --workload cfg3 --emulate-world 8 --steps 100 --no-cpu-baseline > $OUT/r02c22_bench_shard8.json 2>> $OUT/r02c22_bench_err.txt
python bench.py --workload cfg2 --steps 300 --warmup 20 --no-cpu-baseline > $OUT/r02c22_bench_cfg2.json 2>> $OUT/r02c22_bench_err.txt
for f in cfg3 shard8 cfg2; do
python - $OUT/r02c22_bench_$f.json $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "ms/step kernel", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 4),
          "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity_at_size"), "setup", d["setup_s"], "clocks", d["clocks"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
tail -3 $OUT/r02c22_bench_err.txt
