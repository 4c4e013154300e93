"""Context number for BASELINE.md section 4.1: the REFERENCE's own hot-path source (under oracle/jax_shim.py: NumPy, not XLA) on
BASELINE.json configs[1], forward only, on this host's cores.  Needs /root/reference (build container only).  Output kept in
profiles/r02_reference_under_shim_cfg2.json."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import jax_shim
R = jax_shim.load_reference()
from gwinferno_b200 import synthetic
A, SEP, SPL, INT = R["analysis"], R["separable"], R["spline_perturbation"], R["interpolation"]
E, S, I = 70, 4000, 500_000
t0 = time.perf_counter()
pe, inj, const = synthetic.make_catalog(E, S, I, cfg=2)
t_gen = time.perf_counter() - t0
ns = dict(m1=50, q=30, a=16, t=16, z=20)
t0 = time.perf_counter()
rm = SEP.BSplinePrimaryBSplineRatio(ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=3.0, m2min=3.0, mmax=100.0,
                                    kwargs_m={"basis": INT.LogXLogYBSpline}, kwargs_q={"basis": INT.LogYBSpline})
ra = SEP.BSplineIndependentSpinMagnitudes(ns["a"], ns["a"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
rt = SEP.BSplineIndependentSpinTilts(ns["t"], ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
t_setup = time.perf_counter() - t0
rng = np.random.default_rng(0)
def params():
    return dict(m=0.3*rng.standard_normal(50), q=0.3*rng.standard_normal(30), a1=0.3*rng.standard_normal(16), a2=0.3*rng.standard_normal(16),
                t1=0.3*rng.standard_normal(16), t2=0.3*rng.standard_normal(16), lamb=2.7, z=0.3*rng.standard_normal(20))
def weights(d, pe_samples, p):
    w = rm(p["m"], p["q"], pe_samples=pe_samples) * ra(p["a1"], p["a2"], pe_samples=pe_samples) * rt(p["t1"], p["t2"], pe_samples=pe_samples)
    return w * rz(d["redshift"], p["lamb"], p["z"]) / d["prior"]
def forward(p):
    pw, iw = weights(pe, True, p), weights(inj, False, p)
    lb, ln, v = A.per_event_log_bayes_factors(pw)
    lm, lni, vi = A.detection_efficiency(iw, const["total_inj"])
    return float(np.sum(lb) - E * lm)
forward(params())
ts = []
for _ in range(3):
    p = params(); t0 = time.perf_counter(); v = forward(p); ts.append(time.perf_counter() - t0)
print({"reference_source_under_numpy_shim": True, "workload": "cfg2 (70 x 4000 + 5e5)", "forward_only_s": min(ts), "forward_evals_per_s": 1/min(ts), "model_setup_s": t_setup,
       "cores": os.cpu_count(), "note": "NumPy, not XLA; FORWARD only (the reference's gradient is jax reverse-mode: >= 2x this)", "value": v})
