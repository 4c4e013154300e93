#!/usr/bin/env bash
# 2 GPUs: the library-owned exchange between two PROCESSES (CUDA IPC peer memory) under torchrun, with the at-size parity check
set -u
OUT=gpurun_out
TAG=r02c8
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 \
    > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2_err.txt
tail -c 400 $OUT/${TAG}_bench_n2_err.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c8_bench_n2.json"))
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e", "parity_at_size", "result")}, default=str)[:1500])
print("kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "clocks", d["clocks"])
PY
