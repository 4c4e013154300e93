#!/usr/bin/env bash
# 8 GPUs of one box: the library-owned exchange among 8 processes under torchrun, with the at-size parity check on rank 0
set -u
OUT=gpurun_out
TAG=r02c13
mkdir -p $OUT
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 5 \
    > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n${n}_err.txt
tail -c 300 $OUT/${TAG}_bench_n${n}_err.txt
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02c13_bench_n{n}.json"))
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "parity_at_size", "result")}, default=str)[:900])
    print("kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
except Exception as e:
    print("N =", n, "FAILED:", e)
PY
done
