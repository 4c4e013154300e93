#!/usr/bin/env bash
# Round-2, first GPU call: (1) full parity suite of the product build incl. the full-size oracle
# comparisons, (2) the prepared default-off switches against the product build on cfg3 / an 8-way
# shard / cfg2 (variants are sanity-checked through the printed log_l; the winner gets the whole
# parity suite in the next call), (3) runtime switches (fused epilogue, CUDA graph), (4) the other
# configs and the native NUTS driver.  One JSON line per measurement -> gpurun_out/r02_call1.jsonl
set -u
OUT=gpurun_out
TAG=r02c1
mkdir -p $OUT
: > $OUT/${TAG}.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
free -g | head -2 >> $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > $OUT/${TAG}_pytest_gpu.txt 2>&1
tail -5 $OUT/${TAG}_pytest_gpu.txt
run_one() {  # label, library ("" = product), extra env (KEY=VAL ...), bench args
  local label="$1" lib="$2" envs="$3"; shift 3
  local line
  line=$(env $envs ${lib:+GWI_LIBRARY=$lib} timeout 600 python bench.py --no-cpu-baseline "$@" 2>> $OUT/${TAG}_err.txt | tail -1)
  python - "$label" "$line" <<'PY' | tee -a $OUT/r02c1.jsonl
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d.get("roofline", {})
    print(json.dumps({"label": label, "workload": d["config"]["workload"].split(":")[0], "value": round(d["value"], 2), "ms_per_step": round(d["ms_per_step"], 4),
                      "kernel_ms": round(r.get("kernel_ms", float("nan")), 4), "frac": round(r.get("frac", float("nan")), 4), "e2e": round(d["e2e"]["value"], 2),
                      "log_l": d["result"]["log_l"], "n_chunks": d["plan"]["n_chunks"], "sm_mhz": d["clocks"]["sm_mhz"], "setup": d["setup_s"]}))
except Exception as e:
    print(json.dumps({"label": label, "error": str(e), "raw": line[:300]}))
PY
}
for v in "" uni2 exp5; do
  lib=""; [ -n "$v" ] && lib=gwinferno_b200/libgwi_$v.so
  [ -n "$v" ] && [ ! -f "$lib" ] && continue
  run_one "${v:-product}" "$lib" "" --workload cfg3 --steps 20
  run_one "${v:-product} shard8" "$lib" "" --workload cfg3 --emulate-world 8 --steps 50
  run_one "${v:-product} cfg2" "$lib" "" --workload cfg2 --steps 300 --warmup 20
done
if [ -f gwinferno_b200/libgwi_split.so ]; then
  run_one "split" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg3 --steps 20
  run_one "split shard8" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg3 --emulate-world 8 --steps 50
fi
# runtime switches on the product build
run_one "fused" "" "GWI_FUSED_EPILOGUE=1" --workload cfg3 --steps 20
run_one "fused shard8" "" "GWI_FUSED_EPILOGUE=1" --workload cfg3 --emulate-world 8 --steps 50
run_one "fused cfg2" "" "GWI_FUSED_EPILOGUE=1" --workload cfg2 --steps 300 --warmup 20
run_one "graph+fused cfg2" "" "GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1" --workload cfg2 --steps 300 --warmup 20
run_one "graph cfg2" "" "GWI_GRAPH=1" --workload cfg2 --steps 300 --warmup 20
[ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 fused+graph cfg2" gwinferno_b200/libgwi_uni2.so "GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1" --workload cfg2 --steps 300 --warmup 20
[ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 fused shard8" gwinferno_b200/libgwi_uni2.so "GWI_FUSED_EPILOGUE=1" --workload cfg3 --emulate-world 8 --steps 50
res=$(GWI_TEST_EXPERIMENTAL=1 GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1)
echo "{\"label\": \"product graph+fused\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}.jsonl
# guided-schedule tunables on a shard
for tune in "GWI_TUNE_GUIDED_DIV=1 GWI_TUNE_LMIN=16" "GWI_TUNE_GUIDED_DIV=1 GWI_TUNE_LMIN=32"; do
  run_one "product $tune shard8" "" "$tune" --workload cfg3 --emulate-world 8 --steps 50
done
# deep-dim split with the unified pair path
for nd in 2 4; do
  [ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 n_deep=$nd" gwinferno_b200/libgwi_uni2.so "" --workload cfg3 --n-deep $nd --steps 20
done
# the other configs (builder lines)
run_one "product cfg5" "" "" --workload cfg5 --steps 50
[ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 cfg5" gwinferno_b200/libgwi_uni2.so "" --workload cfg5 --steps 50
run_one "product cfg1" "" "" --workload cfg1 --steps 300 --warmup 20
run_one "product cfg4" "" "" --workload cfg4 --steps 5 --warmup 3
run_one "product batch-hint cfg4" "" "GWI_TUNE_BATCH_HINT=1024" --workload cfg4 --steps 5 --warmup 3
[ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 batch-hint cfg4" gwinferno_b200/libgwi_uni2.so "GWI_TUNE_BATCH_HINT=1024" --workload cfg4 --steps 5 --warmup 3
# NUTS ESS/s with the native driver
timeout 600 python tools/nuts_ess.py --driver native --warmup 200 --samples 300 2>> $OUT/${TAG}_err.txt | tail -1 | tee -a $OUT/${TAG}.jsonl
GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1 timeout 900 python tools/nuts_ess.py --driver native --flags 7 --warmup 1000 --samples 500 2>> $OUT/${TAG}_err.txt | tail -1 | tee -a $OUT/${TAG}.jsonl
tail -c 2000 $OUT/${TAG}_err.txt
