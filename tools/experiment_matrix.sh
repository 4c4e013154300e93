#!/usr/bin/env bash
# One gpurun call that measures every prepared, default-off switch against the product build
# (all of them pass the full parity suite on the host warp emulator, tests/emu; none has run on a GPU):
#
#   # here (no GPU): build the variants, they travel with the snapshot
#   bash tools/experiment_matrix.sh build
#   # on the box
#   gpurun --timeout 2400 -- 'bash tools/experiment_matrix.sh run r02'          # everything, ~35 GPU-minutes
#   gpurun --timeout 1200 -- 'ONLY="uni exp5" bash tools/experiment_matrix.sh run r02'   # a subset of the builds
#
# `run` prints and stores (gpurun_out/<tag>_matrix.jsonl) one line per (library, environment, workload):
# parity (pytest -m gpu, tail), evals/s, ms/step, stream-kernel ms and roofline fraction, e2e evals/s.
# Workloads: cfg3 (headline), cfg3 rank-0 shard of an 8-way run (--emulate-world 8), cfg2 (launch-bound).
set -u
MODE="${1:-run}"
TAG="${2:-rXX}"
CSRC=gwinferno_b200/csrc
declare -A VARIANTS=(
  [exp1]="-DGWI_EXP_DEEP_GROUPED=1 -DGWI_EXP_RESET_CUR=1"
  [exp2]="-DGWI_EXP_DEEP_GROUPED=1 -DGWI_EXP_RESET_CUR=1 -DGWI_EXP_SINGLE_BUF=1"
  [exp3]="-DGWI_EXP_RESET_CUR=1 -DGWI_EXP_SINGLE_BUF=1"
  [exp4]="-DGWI_EXP_RED_SPILL=1 -DGWI_EXP_DEEP_GROUPED=1 -DGWI_EXP_RESET_CUR=1 -DGWI_EXP_SINGLE_BUF=1"
  [split]="-DGWI_EXP_SPLIT=1"
  [track]="-DGWI_EXP_TRACK_MAX=1"
  [stage]="-DGWI_EXP_STAGE_DESC=1"
  [uni]="-DGWI_EXP_UNIFIED_PAIR=1"
  [uni2]="-DGWI_EXP_UNIFIED_PAIR=1 -DGWI_EXP_RESET_CUR=1 -DGWI_EXP_RED_SPILL=1"
  [exp5]="-DGWI_EXP_UNIFIED_PAIR=1 -DGWI_EXP_RED_SPILL=1 -DGWI_EXP_DEEP_GROUPED=1 -DGWI_EXP_RESET_CUR=1 -DGWI_EXP_SINGLE_BUF=1"
)
if [ "$MODE" = build ]; then
  make -C $CSRC -j8 > /dev/null || exit 1
  for v in "${!VARIANTS[@]}"; do
    make -C $CSRC -j8 VARIANT=$v EXTRA="${VARIANTS[$v]}" > /dev/null || exit 1
    echo "built $CSRC/../libgwi_$v.so  (${VARIANTS[$v]})"
  done
  exit 0
fi
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_matrix.jsonl
export GWI_TEST_EXPERIMENTAL=1  # include the tests of the default-off switches in every pytest -m gpu below
run_one() {  # label, library ("" = product), extra env (KEY=VAL ...), bench args
  local label="$1" lib="$2" envs="$3"; shift 3
  local line
  line=$(env $envs ${lib:+GWI_LIBRARY=$lib} python bench.py --no-cpu-baseline "$@" 2> $OUT/${TAG}_matrix_err.txt | tail -1)
  python - "$label" "$line" <<'PY' | tee -a $OUT/${TAG}_matrix.jsonl
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d.get("roofline", {})
    print(json.dumps({"label": label, "workload": d["config"]["workload"].split(":")[0], "value": round(d["value"], 2), "ms_per_step": round(d["ms_per_step"], 4),
                      "kernel_ms": round(r.get("kernel_ms", float("nan")), 4), "frac": round(r.get("frac", float("nan")), 4), "e2e": round(d["e2e"]["value"], 2),
                      "log_l": d["result"]["log_l"], "n_chunks": d["plan"]["n_chunks"], "sm_mhz": d["clocks"]["sm_mhz"]}))
except Exception as e:
    print(json.dumps({"label": label, "error": str(e), "raw": line[:200]}))
PY
}
# parity first: a variant whose GPU tests fail is not timed.  ONLY="uni exp5" restricts the compile-time
# variants (the product build is always included); budget ~2.5 GPU-minutes per variant.
for v in "" "${!VARIANTS[@]}"; do
  if [ -n "$v" ] && [ -n "${ONLY:-}" ]; then case " $ONLY " in *" $v "*) ;; *) continue ;; esac; fi
  [ "$v" = split ] && continue  # needs GWI_SPLIT=1: measured separately below
  [ "$v" = track ] && continue  # needs GWI_SPECULATIVE_SHIFT=1: measured separately below
  lib=""; [ -n "$v" ] && lib=gwinferno_b200/libgwi_$v.so
  [ -n "$v" ] && [ ! -f "$lib" ] && continue
  res=$(env ${lib:+GWI_LIBRARY=$lib} python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
  echo "{\"label\": \"${v:-product}\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}_matrix.jsonl
  case "$res" in *failed*|*error*) continue ;; esac
  run_one "${v:-product}" "$lib" "" --workload cfg3
  run_one "${v:-product} shard8" "$lib" "" --workload cfg3 --emulate-world 8
  run_one "${v:-product} cfg2" "$lib" "" --workload cfg2 --steps 200 --warmup 20
done
# runtime switches on the product build
run_one "product fused-epilogue" "" "GWI_FUSED_EPILOGUE=1" --workload cfg3
run_one "product fused-epilogue shard8" "" "GWI_FUSED_EPILOGUE=1" --workload cfg3 --emulate-world 8
run_one "product fused-epilogue cfg2" "" "GWI_FUSED_EPILOGUE=1" --workload cfg2 --steps 200 --warmup 20
# CUDA graph replay of gwi_loglike_host: shows in the `e2e` column (host buffers), not in `value`
run_one "product graph cfg2" "" "GWI_GRAPH=1" --workload cfg2 --steps 200 --warmup 20
run_one "product graph+fused cfg2" "" "GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1" --workload cfg2 --steps 200 --warmup 20
run_one "product graph+fused cfg3" "" "GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1" --workload cfg3
res=$(GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
echo "{\"label\": \"product graph+fused\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}_matrix.jsonl
res=$(GWI_FUSED_EPILOGUE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
echo "{\"label\": \"product fused-epilogue\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}_matrix.jsonl
for tune in "GWI_TUNE_GUIDED_DIV=1 GWI_TUNE_LMIN=16" "GWI_TUNE_GUIDED_DIV=1 GWI_TUNE_LMIN=8" "GWI_TUNE_GUIDED_DIV=2 GWI_TUNE_LMIN=16"; do
  run_one "product $tune shard8" "" "$tune" --workload cfg3 --emulate-world 8
  run_one "product $tune" "" "$tune" --workload cfg3
done
# config-2 size: one flush (~1 700 instructions) per 8-step chunk today; longer minimum slices halve / quarter that
for tune in "GWI_TUNE_LMIN=16" "GWI_TUNE_LMIN=32" "GWI_TUNE_GUIDED_DIV=1 GWI_TUNE_LMIN=32"; do
  run_one "product $tune cfg2" "" "$tune" --workload cfg2 --steps 200 --warmup 20
done
# role-split stream kernel (producer / consumer warp pairs, 128 registers): needs its build AND GWI_SPLIT=1
if [ -f gwinferno_b200/libgwi_split.so ]; then
  res=$(GWI_SPLIT=1 GWI_LIBRARY=gwinferno_b200/libgwi_split.so python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
  echo "{\"label\": \"split GWI_SPLIT=1\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}_matrix.jsonl
  run_one "split GWI_SPLIT=1" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg3
  run_one "split GWI_SPLIT=1 n_deep=4" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg3 --n-deep 4
  run_one "split GWI_SPLIT=1 shard8" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg3 --emulate-world 8
  run_one "split GWI_SPLIT=1 cfg2" gwinferno_b200/libgwi_split.so "GWI_SPLIT=1" --workload cfg2 --steps 200 --warmup 20
fi
# parametric models (cfg1): one pass instead of two with the shift learned from the previous evaluation (e2e column)
if [ -f gwinferno_b200/libgwi_track.so ]; then
  res=$(GWI_SPECULATIVE_SHIFT=1 GWI_LIBRARY=gwinferno_b200/libgwi_track.so python -m pytest tests -m gpu -x -q 2>&1 | tail -1)
  echo "{\"label\": \"track GWI_SPECULATIVE_SHIFT=1\", \"pytest_gpu\": \"$res\"}" | tee -a $OUT/${TAG}_matrix.jsonl
  run_one "product cfg1" "" "" --workload cfg1 --steps 200 --warmup 20
  run_one "track speculative cfg1" gwinferno_b200/libgwi_track.so "GWI_SPECULATIVE_SHIFT=1" --workload cfg1 --steps 200 --warmup 20
fi
# deep-dim split with the unified pair path (the plan's cost model was calibrated for the old path)
for nd in 2 4; do
  [ -f gwinferno_b200/libgwi_uni2.so ] && run_one "uni2 n_deep=$nd" gwinferno_b200/libgwi_uni2.so "" --workload cfg3 --n-deep $nd
done
# cfg4 (1024 chains per launch on the config-2 catalog): batch geometry, alone and with the unified pair path
run_one "product cfg4" "" "" --workload cfg4 --steps 5 --warmup 3
run_one "product batch-hint cfg4" "" "GWI_TUNE_BATCH_HINT=1024" --workload cfg4 --steps 5 --warmup 3
[ -f gwinferno_b200/libgwi_uni.so ] && run_one "uni cfg4" gwinferno_b200/libgwi_uni.so "" --workload cfg4 --steps 5 --warmup 3
[ -f gwinferno_b200/libgwi_uni.so ] && run_one "uni batch-hint cfg4" gwinferno_b200/libgwi_uni.so "GWI_TUNE_BATCH_HINT=1024" --workload cfg4 --steps 5 --warmup 3
# NUTS ESS/s (the second half of the metric): NumPy driver vs native driver vs native + windowed dense mass
for args in "--driver numpy" "--driver native" "--driver native --flags 7 --warmup 1000 --samples 500"; do
  python tools/nuts_ess.py $args 2>> $OUT/${TAG}_matrix_err.txt | tail -1 | tee -a $OUT/${TAG}_matrix.jsonl
done
GWI_GRAPH=1 GWI_FUSED_EPILOGUE=1 python tools/nuts_ess.py --driver native --flags 7 --warmup 1000 --samples 500 2>> $OUT/${TAG}_matrix_err.txt | tail -1 | tee -a $OUT/${TAG}_matrix.jsonl
