#!/usr/bin/env bash
# One gpurun call that regenerates every artefact kept under profiles/ for the default workload:
#
#   gpurun --timeout 900 -- 'bash tools/gpu_validation.sh r02'
#
# (1) pytest -m gpu, (2) smoke(), (3) the default bench line, (4) the ncu launch list of the same bench
# command (device time of every launch: compare SHARES with the bench, not absolutes), (5) one
# `ncu --set full` capture of the stream kernel.  Outputs go to gpurun_out/<tag>_*; copy the ones to be
# judged into profiles/ (tools/ncu_extract.sh turns the .ncu-rep into the committed text extract).
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
python -m pytest tests -m gpu -q 2>&1 | tail -3 > "$OUT/${TAG}_pytest_gpu.txt"
cat "$OUT/${TAG}_pytest_gpu.txt"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee "$OUT/${TAG}_smoke.txt"
python bench.py > "$OUT/${TAG}_bench_cfg3_n1.json" 2> "$OUT/${TAG}_bench_err.txt"
python - "$OUT/${TAG}_bench_cfg3_n1.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("bench:", round(d["value"], 1), d["unit"], "|", round(d["ms_per_step"], 3), "ms/step | stream kernel", round(d["roofline"]["kernel_ms"], 3), "ms =",
      round(100 * d["roofline"]["frac"], 1), "% of", d["roofline"]["peak"], d["roofline"]["unit"], "| e2e", round(d["e2e"]["value"], 1), "| clocks", d["clocks"])
PY
# launch list of the same command (short run: ncu serialises and replays)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file "$OUT/${TAG}_cfg3_launches.csv" \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2> "$OUT/${TAG}_ncu_launches_err.txt"
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 3 -c 1 -f -o "$OUT/${TAG}_stream_kernel" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> "$OUT/${TAG}_ncu_full_err.txt"
# the second half of the metric: NUTS ESS/s on config 2, sampler loop in native code (csrc/nuts.cpp)
python tools/nuts_ess.py --driver native --warmup 200 --samples 300 > "$OUT/${TAG}_nuts_cfg2.json" 2> "$OUT/${TAG}_nuts_err.txt"
tail -c 600 "$OUT/${TAG}_nuts_cfg2.json"; echo
ls -la "$OUT" | tail -12
