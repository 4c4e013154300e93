"""Throughput of gwi_loglike_batch_host on BASELINE.json configs[1] for small chain batches (what the multi-chain NUTS
driver calls once per round of leapfrog steps), and the multi-chain sampler itself.  GPU only:
    python tools/chains_probe.py [--nuts 8,16]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from gwinferno_b200 import nuts, pipeline, workloads  # noqa: E402
from gwinferno_b200.likelihood import PopulationLikelihood  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,2,4,8,16,32")
ap.add_argument("--nuts", default="")
ap.add_argument("--warmup", type=int, default=1000)
ap.add_argument("--samples", type=int, default=800)
ap.add_argument("--hint-div", type=int, default=2, help="model batch_hint = chains / this (2: two alternating groups)")
args = ap.parse_args()
pe, inj, const, z_range = workloads.shard_catalog("cfg2", 0, 1)
weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
low, lam0, p0 = workloads.lower_workload(weights, params_fn, pe, inj)
P = low.spec.n_params
rng = np.random.default_rng(0)
for K in [int(x) for x in args.batches.split(",") if x]:
    eng = PopulationLikelihood(low, const["total_inj"], batch_hint=K)
    lams = np.stack([lam0 * (1.0 + 0.01 * rng.standard_normal(P)) for _ in range(K)])
    for _ in range(20):
        eng.model.loglike_batch_host(lams, const["E"]) if K > 1 else eng.model.loglike_host(lams[0], const["E"])
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        eng.model.loglike_batch_host(lams, const["E"]) if K > 1 else eng.model.loglike_host(lams[0], const["E"])
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"batch": K, "us_per_call": dt * 1e6, "chain_evals_per_s": K / dt, "info": {k: eng.info()[k] for k in ("n_chunks", "grid_blocks", "block_threads")}}), flush=True)
    eng.model.close()
blocks = pipeline.bspline_prior_blocks(low.slots_for, p0)
dim = P - 1
for K in [int(x) for x in args.nuts.split(",") if x]:
    eng = PopulationLikelihood(low, const["total_inj"], batch_hint=max(1, K // args.hint_div))
    th0 = 0.1 * np.random.default_rng(1).standard_normal((K, dim))
    t0 = time.perf_counter()
    s, infos = nuts.nuts_native_chains(eng, blocks, th0, args.warmup, args.samples, Nobs=const["E"], seed=100, max_depth=8, flags=7)
    wall = time.perf_counter() - t0
    ess = np.array([sum(nuts.effective_sample_size(s[c, :, i]) for c in range(K)) for i in range(dim)])
    rhat = np.array([nuts.split_rhat(s[:, :, i]) for i in range(dim)])
    t_s = max(i["sampling_seconds"] for i in infos)
    print(json.dumps({"nuts_chains": K, "batch_hint": max(1, K // args.hint_div), "ess_min": ess.min(), "ess_median": float(np.median(ess)), "ess_min_per_s": ess.min() / t_s, "ess_median_per_s": float(np.median(ess)) / t_s,
                      "grad_evals_per_s": sum(i["leapfrogs_sampling"] for i in infos) / t_s, "rhat_max": float(np.nanmax(rhat)), "sampling_s": t_s, "wall_s": wall,
                      "accept": [round(i["mean_accept"], 3) for i in infos]}), flush=True)
    eng.model.close()
