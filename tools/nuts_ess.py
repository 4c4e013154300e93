#!/usr/bin/env python
"""NUTS ESS/s of the B-spline population model through the fused GPU likelihood (the second half of
BASELINE.json's metric).  Priors and smoothing penalties follow examples/simple_bspline_example.py
(bspline_mass_prior m_tau=1 q_tau=1, bspline_spin_prior a_tau=25 ct_tau=25, bspline_redshift_prior
z_tau=1, lamb ~ N(0,3)); the sampler is gwinferno_b200/nuts.py.  Prints one JSON line."""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gwinferno_b200 import nuts, pipeline, workloads  # noqa: E402
from gwinferno_b200.likelihood import PopulationLikelihood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--warmup", type=int, default=150)
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--max-depth", type=int, default=8)
    ap.add_argument("--driver", choices=["native", "numpy"], default="native",
                    help="native: the whole NUTS loop in libgwi (csrc/nuts.cpp); numpy: gwinferno_b200/nuts.py around gwi_loglike_host")
    ap.add_argument("--flags", type=int, default=0,
                    help="native driver only: GWI_NUTS_* bits (1 multinomial, 2 windowed adaptation, 4 dense mass matrix)")
    args = ap.parse_args()
    pe, inj, const, z_range = workloads.shard_catalog(args.workload, 0, 1, scale=args.scale)
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, lam0, p0 = workloads.lower_workload(weights, params_fn, pe, inj)
    eng = PopulationLikelihood(low, const["total_inj"])
    Nobs = const["E"]

    def loglike(lam):
        log_l, grad, head = eng.loglike(lam, Nobs=Nobs)
        return log_l, grad

    blocks = pipeline.bspline_prior_blocks(low.slots_for, p0)  # pipeline/utils.py:163-216 defaults, example taus
    post = nuts.BSplinePosterior(loglike, low.spec.n_params, blocks)
    rng = np.random.default_rng(0)
    theta0 = 0.1 * rng.standard_normal(post.dim)
    t0 = time.perf_counter()
    if args.driver == "native":
        samples, info = nuts.nuts_native(eng, blocks, theta0, args.warmup, args.samples, Nobs=Nobs, seed=0, max_depth=args.max_depth, flags=args.flags)
        post.n_evals = info["n_evals"]
    else:
        samples, info = nuts.nuts(post, theta0, args.warmup, args.samples, rng, max_depth=args.max_depth)
    wall = time.perf_counter() - t0
    ess = np.array([nuts.effective_sample_size(samples[:, i]) for i in range(samples.shape[1])])
    out = {
        "metric": "NUTS ESS/s", "workload": args.workload, "E": const["E"], "S": const["S"], "I": const["I"], "dim": post.dim,
        "warmup": args.warmup, "samples": args.samples, "max_tree_depth": args.max_depth, "driver": args.driver, "flags": args.flags,
        "ess_min": float(ess.min()), "ess_median": float(np.median(ess)),
        "ess_min_per_s": float(ess.min() / info["sampling_seconds"]), "ess_median_per_s": float(np.median(ess) / info["sampling_seconds"]),
        "grad_evals_total": post.n_evals, "grad_evals_per_s_overall": post.n_evals / wall, "wall_s": wall, **info,
        "sampler": ("csrc/nuts.cpp (native NUTS loop, diagonal mass)" if args.driver == "native" else "gwinferno_b200/nuts.py (NumPy NUTS, diagonal mass)")
        + "; likelihood+gradient on the GPU via gwi_loglike_host",
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
