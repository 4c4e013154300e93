"""The BASELINE.json workloads, written the way a GWInferno user writes them: reference-style
model objects over the (sharded) synthetic catalog and a ``get_weights`` product
(examples/simple_bspline_example.py:25-89, examples/simple_powerlaw_peak_example.py:29-113).
Used by bench.py and the tests; the heavy lifting is in libgwi.so."""

import numpy as np

from . import lowering, synthetic
from . import models as M

NSPLINES = dict(m1=50, q=30, a=16, t=16, z=20)  # gwinferno/pipeline/utils.py:29-33
MMIN, MMAX = 3.0, 100.0  # gwinferno/pipeline/utils.py:34-35

WORKLOADS = {
    # name: (BASELINE.json config index, model family, E, S, I)
    "cfg1": (0, "plpeak", 69, 3000, 100_000),
    "cfg2": (1, "bspline", 70, 4000, 500_000),
    "cfg3": (2, "bspline", 300, 10_000, 100_000_000),
    "cfg4": (3, "bspline", 70, 4000, 500_000),  # cfg-2 catalog, 1024 chains per step (N_CHAINS)
    "cfg5": (4, "bspline_iid", 200, 8000, 20_000_000),
}
N_CHAINS = {"cfg4": 1024}


def shard_catalog(name, rank=0, world=1, scale=1.0, all_reduce_minmax=None, shard_by="bucket"):
    """This rank's share of workload ``name``: whole events ``e % world == rank`` and either whole
    (m1-piece, q-piece) buckets of the found injections (``shard_by="bucket"``, default) or the index
    range ``[rank I/world, (rank+1) I/world)`` (``shard_by="index"``).  Returns ``(pe, inj, constants, z_range)``;
    ``z_range`` is the redshift range of the WHOLE catalog (parametric.py:114-115), obtained with
    ``all_reduce_minmax(lo, hi) -> (min over ranks of lo, max over ranks of hi)`` when sharded."""
    cfg_idx, family, E, S, I = WORKLOADS[name]
    cfg = 2 if name == "cfg4" else int(name[3:])  # cfg4 = the cfg-2 catalog
    S = max(8, int(round(S * scale)))
    I = max(64 * world, int(round(I * scale)))
    pe_all = synthetic.make_events(E, S, 1000 * cfg + 0)
    zpe = (float(pe_all["redshift"].min()), float(pe_all["redshift"].max()))
    pe = {k: np.ascontiguousarray(v[rank::world]) for k, v in pe_all.items()}
    del pe_all
    if world == 1 or shard_by == "index":
        a, b = rank * I // world, (rank + 1) * I // world
        inj = synthetic.make_injections(I, 1000 * cfg + 1, a, b)
    else:
        # Shard by PIECE BUCKET: the found injections are grouped by the spline pieces of their two
        # leading coordinates (log m1: 48 pieces, q: 28 pieces at the default spline counts) and whole
        # buckets are dealt to the ranks (largest first, to the least loaded rank).  Any partition
        # gives the same sums; this one keeps every rank's piece-sorted runs exactly as long as on
        # one GPU and spreads sparse and dense regions of parameter space evenly over the ranks.
        n1, n2 = NSPLINES["m1"] - 3, NSPLINES["q"] - 3
        cols_all = synthetic.make_injections(I, 1000 * cfg + 1)
        j1 = np.clip(np.floor((np.log(cols_all["mass_1"]) - np.log(MMIN)) / (np.log(MMAX) - np.log(MMIN)) * n1), 0, n1 - 1).astype(np.int64)
        qmin = MMIN / MMAX
        j2 = np.clip(np.floor((cols_all["mass_ratio"] - qmin) / (1.0 - qmin) * n2), 0, n2 - 1).astype(np.int64)
        bucket = j1 * n2 + j2
        counts = np.bincount(bucket, minlength=n1 * n2)
        owner = np.zeros(n1 * n2, dtype=np.int64)
        load = np.zeros(world, dtype=np.int64)
        for b in np.argsort(-counts, kind="stable"):
            r = int(np.argmin(load))
            owner[b] = r
            load[r] += counts[b]
        keep = owner[bucket] == rank
        inj = {k: np.ascontiguousarray(v[keep]) for k, v in cols_all.items()}
        del cols_all, bucket, j1, j2, keep
    zlo, zhi = float(inj["redshift"].min()), float(inj["redshift"].max())
    if world > 1:
        zlo, zhi = all_reduce_minmax(zlo, zhi)
    z_range = (max(zpe[0], zlo), min(zpe[1], zhi))
    const = {"total_inj": float(4 * I), "obs_time": 1.0, "nObs": E, "n_events_local": pe["redshift"].shape[0], "E": E, "S": S, "I": I, "family": family}
    return pe, inj, const, z_range


def build_model(family, pe, inj, z_range=None, nsplines=None):
    """Reference-style model objects + ``weights(datadict, pe_samples, params)`` + a parameter
    factory ``params(seed)`` for the given family."""
    ns = dict(NSPLINES, **(nsplines or {}))
    if family == "bspline":
        mm = M.BSplinePrimaryBSplineRatio(ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=MMIN, m2min=MMIN, mmax=MMAX,
                                          kwargs_m={"basis": M.LogXLogYBSpline}, kwargs_q={"basis": M.LogYBSpline})  # pipeline/utils.py:104-118
        ma = M.BSplineIndependentSpinMagnitudes(ns["a"], ns["a"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
        mt = M.BSplineIndependentSpinTilts(ns["t"], ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
        mz = M.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"], z_range=z_range)

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ma(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
            w = w * mt(p["tilt1_cs"], p["tilt2_cs"], pe_samples=pe_samples) * mz(d["redshift"], p["lamb"], p["z_cs"])
            return w / d["prior"]

        def params(seed, scale=0.3):
            rng = np.random.default_rng(seed)
            p = dict(mass_cs=scale * rng.standard_normal(ns["m1"]), q_cs=scale * rng.standard_normal(ns["q"]), a1_cs=scale * rng.standard_normal(ns["a"]),
                     a2_cs=scale * rng.standard_normal(ns["a"]), tilt1_cs=scale * rng.standard_normal(ns["t"]), tilt2_cs=scale * rng.standard_normal(ns["t"]),
                     lamb=np.float64(2.7), z_cs=scale * rng.standard_normal(ns["z"]))
            p["z_cs"][0] = 0.0  # pipeline/utils.py:213-214
            return p

        return weights, params
    if family == "bspline_iid":
        mm = M.BSplineIIDComponentMasses(ns["m1"], pe["mass_1"], pe["mass_2"], inj["mass_1"], inj["mass_2"], mmin=MMIN, mmax=MMAX)
        ma = M.BSplineIIDSpinMagnitudes(ns["a"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
        mt = M.BSplineIIDSpinTilts(ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
        mz = M.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"], z_range=z_range)

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], beta=p["beta"], pe_samples=pe_samples) * ma(p["a_cs"], pe_samples=pe_samples) * mt(p["tilt_cs"], pe_samples=pe_samples)
            return w * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        def params(seed, scale=0.3):
            rng = np.random.default_rng(seed)
            p = dict(mass_cs=scale * rng.standard_normal(ns["m1"]), beta=np.float64(1.1), a_cs=scale * rng.standard_normal(ns["a"]),
                     tilt_cs=scale * rng.standard_normal(ns["t"]), lamb=np.float64(2.7), z_cs=scale * rng.standard_normal(ns["z"]))
            p["z_cs"][0] = 0.0
            return p

        return weights, params
    if family == "plpeak":
        mz = M.PowerlawRedshiftModel(pe["redshift"], inj["redshift"], z_range=z_range)

        def weights(d, pe_samples, p):
            p_m1q = M.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], MMIN, MMAX, p["mu_peak"], p["sig_peak"], p["lambda_m"])
            p_a = M.independent_spin_magnitude_beta_dist(d["a_1"], d["a_2"], p["alpha_a1"], p["beta_a1"], p["alpha_a2"], p["beta_a2"])
            p_ct = M.independent_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["lambda_ct1"], p["lambda_ct2"], p["sig_ct1"], p["sig_ct2"])
            return p_m1q * p_a * p_ct * mz(d["redshift"], p["lamb"]) / d["prior"]

        def params(seed, scale=0.02):
            rng = np.random.default_rng(seed)
            base = dict(alpha=-2.8, beta=1.4, mu_peak=34.0, sig_peak=4.5, lambda_m=0.08, alpha_a1=1.8, beta_a1=4.2, alpha_a2=2.1, beta_a2=3.3,
                        lambda_ct1=0.6, lambda_ct2=0.35, sig_ct1=1.2, sig_ct2=0.8, lamb=2.4)
            return {k: np.float64(v * (1.0 + scale * rng.standard_normal())) for k, v in base.items()}

        return weights, params
    raise KeyError(family)


def lower_workload(weights, params_fn, pe, inj, seed=0):
    p = params_fn(seed)
    pe_w, inj_w = weights(pe, True, p), weights(inj, False, p)
    low = lowering.lower(pe_w, inj_w)
    lam = lowering.flatten_params(pe_w, low.spec.n_params)
    return low, lam, p
