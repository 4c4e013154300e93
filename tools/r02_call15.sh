#!/usr/bin/env bash
set -u
OUT=gpurun_out
TAG=r02c15
mkdir -p $OUT
# full captures of the SMALL kernels of one evaluation (8-way shard geometry): where do ~15 us per tiny kernel go?
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"prologue_kernel|reduce_kernel|finish_kernel|partial_tail_kernel" -s 24 -c 6 -f -o $OUT/${TAG}_tail_kernels \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --emulate-world 8 > /dev/null 2> $OUT/${TAG}_ncu_err.txt
tail -3 $OUT/${TAG}_ncu_err.txt
ls -la $OUT/${TAG}_tail_kernels.ncu-rep
