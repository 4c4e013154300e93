#!/usr/bin/env python
"""Dynamic path statistics of the stream kernel from the host warp emulator (no GPU, no timing):
how often a warp iteration (32 lanes x 2 samples) has at least one lane on the sequential
piece-change path -- on lock-step hardware such an iteration issues BOTH the staged and the
sequential code -- plus piece changes per sample and samples per record flush.

    make -C tests/emu -j8 VARIANT=stats EXTRA=-DGWI_EMU_STATS=1
    python tools/emu_path_stats.py --workload cfg3 --scale 0.1 [--emulate-world 8]

Test infrastructure (tests/emu); the numbers depend on the catalog size (piece-sorted runs get longer
with more samples), so use the scale of the question being asked."""

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--emulate-world", type=int, default=1)
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--n-deep", type=int, default=-1, help="force the number of deep dims (-1 = the plan's cost model)")
    args = ap.parse_args()
    os.environ.setdefault("GWI_EMU_SMS", str(args.sms))
    from tests import emu

    lib = emu.activate("stats")
    stats = (C.c_ulonglong * 16).in_dll(lib, "gwi_emu_stats")
    from gwinferno_b200 import workloads
    from gwinferno_b200.likelihood import PopulationLikelihood

    t0 = time.time()
    pe, inj, const, z_range = workloads.shard_catalog(args.workload, 0, args.emulate_world, scale=args.scale, all_reduce_minmax=lambda lo, hi: (lo, hi))
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, lam, _ = workloads.lower_workload(weights, params_fn, pe, inj)
    eng = PopulationLikelihood(low, const["total_inj"], n_deep=args.n_deep)
    info = eng.info()
    t1 = time.time()
    for i in range(16):
        stats[i] = 0
    log_l, grad, head = eng.loglike(lam, Nobs=const["E"])
    t2 = time.time()
    s = [int(stats[i]) for i in range(16)]
    iters = max(1, s[0])
    samples = info["n_padded"]
    print(json.dumps({
        "workload": args.workload, "scale": args.scale, "emulate_world": args.emulate_world, "n_padded": samples, "n_chunks": info["n_chunks"], "n_deep": info["n_deep"],
        "warp_iterations": s[0], "frac_iterations_with_a_lane_on_the_sequential_path": s[1] / iters,
        "mean_lanes_on_it_when_taken": s[2] / max(1, s[1]), "frac_lane_pairs_on_it": s[2] / (32.0 * iters),
        "piece_changes_per_sample_by_leading_dim": [s[3 + d] / samples for d in range(4)],
        "samples_per_flush": samples / max(1, s[7]), "log_l": log_l, "passed": head["passed"],
        "setup_s": round(t1 - t0, 1), "emulated_eval_s": round(t2 - t1, 1)}))


if __name__ == "__main__":
    main()
