#!/usr/bin/env python
"""Set-up of a BASELINE workload with the found injections GENERATED ON THE DEVICE (gwi_synth_injections: Philox keyed by the
global injection index) and a device-resident catalog: nothing but the PE samples (3e6 of 1.03e8 samples at cfg3) and the chunk
descriptors crosses PCIe.  Prints one JSON line with the set-up times and the parity of one evaluation against the plain-C oracle
on the SAME catalog (downloaded from the device for the check).

    python tools/device_catalog_probe.py --workload cfg3
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gwinferno_b200 import capi, lowering, synthetic, workloads  # noqa: E402
from gwinferno_b200.likelihood import PopulationLikelihood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg2", "cfg3", "cfg5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    _, family, E, S, I = workloads.WORKLOADS[args.workload]
    cfg = int(args.workload[3:])
    S, I = max(8, int(round(S * args.scale))), max(64, int(round(I * args.scale)))
    capi.load_library()
    # warm-up: CUDA context, module load of libgwi's kernels (a tiny model through the same calls), so that the times below are
    # the work itself and not the process start-up
    wpe, winj, wconst = synthetic.make_catalog(4, 64, 2000, cfg=9)
    wd_pe, wd_inj = {k: capi.DeviceArray(v) for k, v in wpe.items()}, capi.synth_injections_device(1, 0, 2000)
    ww, wp = workloads.build_model(family, wd_pe, wd_inj, z_range=(0.01, 1.5))
    wlow, wlam, _ = workloads.lower_workload(ww, wp, wd_pe, wd_inj)
    weng = PopulationLikelihood(wlow, 8000.0)
    weng.loglike(wlam, Nobs=4)
    weng.model.close()
    t = {}
    t0 = time.perf_counter()
    pe = synthetic.make_events(E, S, 1000 * cfg + 0)
    t["generate_pe_host"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pe_d = {k: capi.DeviceArray(v) for k, v in pe.items()}
    t["upload_pe"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    inj_d = capi.synth_injections_device(1000 * cfg + 1, 0, I)
    zlo, zhi = inj_d["redshift"].minmax()  # (synchronises)
    t["generate_injections_device"] = time.perf_counter() - t0
    z_range = (max(float(pe["redshift"].min()), zlo), min(float(pe["redshift"].max()), zhi))
    t0 = time.perf_counter()
    weights, params_fn = workloads.build_model(family, pe_d, inj_d, z_range=z_range)
    low, lam, _ = workloads.lower_workload(weights, params_fn, pe_d, inj_d)
    eng = PopulationLikelihood(low, float(4 * I))
    t["lower_and_gwi_model_create"] = time.perf_counter() - t0
    info = eng.info()
    t0 = time.perf_counter()
    log_l, grad, head = eng.loglike(lam, Nobs=E)
    t["first_evaluation"] = time.perf_counter() - t0
    line = {"workload": args.workload, "E": E, "S": S, "I": I, "setup_s": t, "setup_total_s": sum(v for k, v in t.items() if k != "first_evaluation"),
            "gwi_model_create": info["plan_seconds"], "plan_on_device": info["plan_on_device"], "log_l": log_l, "passed": head["passed"],
            "h2d_bytes": int(sum(v.nbytes for v in pe.values())), "n_padded": info["n_padded"]}
    if not args.no_parity:
        from oracle import c_oracle, popmodel

        t0 = time.perf_counter()
        inj = {k: v.download() for k, v in inj_d.items()}
        t["download_for_parity"] = time.perf_counter() - t0
        w2, p2 = workloads.build_model(family, pe, inj, z_range=z_range)
        low2, lam2, _ = workloads.lower_workload(w2, p2, pe, inj)
        assert np.array_equal(lam, lam2)
        ev = c_oracle.evaluate(low2.spec, low2.pe_cols, low2.inj_cols, float(4 * I), lam2, want_jac=True, want_neff_jac=False, n_threads=os.cpu_count() or 1)
        l_o, g_o, _ = popmodel.hierarchical_log_likelihood(ev, E, min_neff_cut=True)
        line["parity"] = {"rel_log_l": float(abs(log_l - l_o) / abs(l_o)), "rel_grad": float(np.max(np.abs(grad - g_o)) / np.max(np.abs(g_o))), "log_l_oracle": float(l_o),
                          "oracle_seconds": time.perf_counter() - t0}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
