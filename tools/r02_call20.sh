#!/usr/bin/env bash
# r02 call 20: multi-chain NUTS with two alternating groups (batch_hint = chains / 2) against all chains together
set -u
OUT=gpurun_out
TAG=r02c20
mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -q -x -k "chains_advance" 2>&1 | tail -2 | tee $OUT/${TAG}_pytest.txt
timeout 900 python tools/chains_probe.py --batches "" --nuts 8,16,32 --hint-div 2 2>&1 | tee $OUT/${TAG}_chains_probe.txt
timeout 600 python tools/chains_probe.py --batches "" --nuts 16 --hint-div 4 2>&1 | tee -a $OUT/${TAG}_chains_probe.txt
