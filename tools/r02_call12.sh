#!/usr/bin/env bash
set -u
OUT=gpurun_out
TAG=r02c12
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $OUT/${TAG}_pytest_gpu.txt 2>&1
tail -3 $OUT/${TAG}_pytest_gpu.txt
: > $OUT/${TAG}.jsonl
run_one() {  # label, extra env (KEY=VAL ...), bench args
  local label="$1" envs="$2"; shift 2
  local line
  line=$(env $envs timeout 300 python bench.py --no-cpu-baseline "$@" 2>> $OUT/${TAG}_err.txt | tail -1)
  python - "$label" "$line" <<'PY' | tee -a gpurun_out/r02c12.jsonl
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d.get("roofline", {})
    print(json.dumps({"label": label, "workload": d["config"]["workload"].split(":")[0], "value": round(d["value"], 2), "ms_per_step": round(d["ms_per_step"], 4),
                      "kernel_ms": round(r.get("kernel_ms", float("nan")), 4), "frac": round(r.get("frac", float("nan")), 4), "e2e": round(d["e2e"]["value"], 2),
                      "log_l": d["result"]["log_l"], "n_chunks": d["plan"]["n_chunks"], "block": d["plan"]["block_threads"], "n_deep": d["plan"]["n_deep"], "sm_mhz": d["clocks"]["sm_mhz"], "l2": d["config"]["l2_policy"][:20]}))
except Exception as e:
    print(json.dumps({"label": label, "error": str(e), "raw": line[:300]}))
PY
}
for u in 1 0; do
  run_one "pdl=$u cfg3" "GWI_PDL=$u" --workload cfg3 --steps 30
  run_one "pdl=$u shard8" "GWI_PDL=$u" --workload cfg3 --emulate-world 8 --steps 100
  run_one "pdl=$u cfg2" "GWI_PDL=$u" --workload cfg2 --steps 500 --warmup 20
  for w in "cfg3 --emulate-world 8" "cfg2 --steps 300 --warmup 20"; do
    echo "== pdl=$u $w" >> $OUT/${TAG}_phases.txt
    GWI_PDL=$u GWI_PHASE_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --workload $w 2>&1 >/dev/null | grep "gwi phases" >> $OUT/${TAG}_phases.txt
  done
done
cat $OUT/${TAG}_phases.txt
run_one "cfg5" "" --workload cfg5 --steps 50
run_one "cfg4" "" --workload cfg4 --steps 5 --warmup 3
tail -c 800 $OUT/${TAG}_err.txt
