#!/usr/bin/env bash
set -u
OUT=gpurun_out
TAG=r02c7
mkdir -p $OUT
: > $OUT/${TAG}.jsonl
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/${TAG}_pytest_gpu.txt 2>&1
tail -4 $OUT/${TAG}_pytest_gpu.txt
run_one() {  # label, extra env (KEY=VAL ...), bench args
  local label="$1" envs="$2"; shift 2
  local line
  line=$(env $envs timeout 300 python bench.py --no-cpu-baseline "$@" 2>> $OUT/${TAG}_err.txt | tail -1)
  python - "$label" "$line" <<'PY' | tee -a gpurun_out/r02c7.jsonl
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d.get("roofline", {})
    print(json.dumps({"label": label, "workload": d["config"]["workload"].split(":")[0], "value": round(d["value"], 2), "ms_per_step": round(d["ms_per_step"], 4),
                      "kernel_ms": round(r.get("kernel_ms", float("nan")), 4), "frac": round(r.get("frac", float("nan")), 4), "e2e": round(d["e2e"]["value"], 2),
                      "log_l": d["result"]["log_l"], "n_chunks": d["plan"]["n_chunks"], "block": d["plan"]["block_threads"], "n_deep": d["plan"]["n_deep"], "sm_mhz": d["clocks"]["sm_mhz"]}))
except Exception as e:
    print(json.dumps({"label": label, "error": str(e), "raw": line[:300]}))
PY
}
for k in 0 1; do
  run_one "cta=$k" "GWI_CTA_KERNEL=$k" --workload cfg3 --steps 20
  run_one "cta=$k n_deep=3" "GWI_CTA_KERNEL=$k" --workload cfg3 --n-deep 3 --steps 20
  run_one "cta=$k shard8" "GWI_CTA_KERNEL=$k" --workload cfg3 --emulate-world 8 --steps 50
  run_one "cta=$k shard8 n_deep=3" "GWI_CTA_KERNEL=$k" --workload cfg3 --emulate-world 8 --n-deep 3 --steps 50
  run_one "cta=$k shard2" "GWI_CTA_KERNEL=$k" --workload cfg3 --emulate-world 2 --steps 30
  run_one "cta=$k cfg2" "GWI_CTA_KERNEL=$k" --workload cfg2 --steps 300 --warmup 20
  run_one "cta=$k cfg5" "GWI_CTA_KERNEL=$k" --workload cfg5 --steps 50
  run_one "cta=$k cfg5 n_deep=3" "GWI_CTA_KERNEL=$k" --workload cfg5 --n-deep 3 --steps 50
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu && /tmp/dmma_probe | tee $OUT/${TAG}_dmma_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $OUT/${TAG}_cfg3_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> $OUT/${TAG}_ncu_launches_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $OUT/${TAG}_shard8_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-world 8 > /dev/null 2>> $OUT/${TAG}_ncu_launches_err.txt
tail -c 1200 $OUT/${TAG}_err.txt
