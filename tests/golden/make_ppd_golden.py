"""Golden vectors for gwinferno_b200/postprocess.py FROM THE REFERENCE'S OWN CODE.

Run in the build container (needs /root/reference; cannot run on the GPU box):

    python tests/golden/make_ppd_golden.py

Loads gwinferno/postprocess/calculations.py unmodified under oracle/jax_shim.py (fp64) and calls its
functions (calculations.py:20-60, 63-93, 133-178, 181-242, 244-276) for a few seeded "posterior draws";
inputs and outputs go to tests/golden/ppd_reference.npz.
"""

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import jax_shim  # noqa: E402

R = jax_shim.load_reference()
pkg = types.ModuleType("gwinferno.postprocess")
pkg.__path__ = [os.path.join(jax_shim.REFERENCE_ROOT, "gwinferno", "postprocess")]
sys.modules["gwinferno.postprocess"] = pkg
spec = importlib.util.spec_from_file_location("gwinferno.postprocess.calculations", os.path.join(pkg.__path__[0], "calculations.py"))
CALC = importlib.util.module_from_spec(spec)
sys.modules[spec.name] = CALC
spec.loader.exec_module(CALC)

rng = np.random.default_rng(77)
n = 3
ns = {"m1": 14, "q": 9, "a": 8, "tilt": 7, "a1": 8, "a2": 6, "tilt1": 7, "tilt2": 9}
mmin, mmax = 4.0, 90.0
inp = dict(
    n=n, mmin=mmin, mmax=mmax, **{f"ns_{k}": v for k, v in ns.items()},
    m_cs=rng.normal(0, 1.0, (n, ns["m1"])), q_cs=rng.normal(0, 1.0, (n, ns["q"])),
    a_cs=rng.normal(0, 1.0, (n, ns["a"])), t_cs=rng.normal(0, 1.0, (n, ns["tilt"])),
    a1_cs=rng.normal(0, 1.0, (n, ns["a1"])), a2_cs=rng.normal(0, 1.0, (n, ns["a2"])),
    t1_cs=rng.normal(0, 1.0, (n, ns["tilt1"])), t2_cs=rng.normal(0, 1.0, (n, ns["tilt2"])),
    rate=rng.uniform(10, 40, n), frac=rng.uniform(0.2, 1.0, n),
    alpha=rng.uniform(2.5, 4.0, n), beta=rng.uniform(0.5, 2.0, n), mu_peak=rng.uniform(30, 38, n), sig_peak=rng.uniform(2, 6, n), lam=rng.uniform(0.02, 0.2, n),
    alpha_a=rng.uniform(1.2, 3.0, n), beta_a=rng.uniform(2.0, 6.0, n), sig_ct=rng.uniform(0.3, 2.0, n), lambda_ct=rng.uniform(0.1, 0.9, n),
    lamb_z=rng.uniform(0.5, 4.0, n), z_cs=rng.normal(0, 0.5, (n, 7)),
    z_pe=rng.uniform(0.02, 1.2, (4, 30)), z_inj=rng.uniform(0.01, 1.4, 200),
)
out = {}
out["bs_mpdfs"], out["bs_ms"], out["bs_qpdfs"], out["bs_qs"] = CALC.calculate_bspline_mass_ppds(inp["m_cs"], inp["q_cs"], ns, mmin, mmax, rate=inp["rate"], pop_frac=inp["frac"])
out["pp_mpdfs"], _, out["pp_qpdfs"], _ = CALC.calculate_powerlaw_peak_mass_ppds(inp["alpha"], inp["beta"], inp["mu_peak"], inp["sig_peak"], inp["lam"], mmin, mmax, rate=inp["rate"])
out["iid_apdfs"], out["aa"], out["iid_ctpdfs"], out["cc"] = CALC.calculate_bspline_spin_ppds(inp["a_cs"], inp["t_cs"], ns, rate=inp["rate"], pop_frac=inp["frac"])
r = CALC.calculate_bspline_spin_ppds(inp["a1_cs"], inp["t1_cs"], ns, a2_cs=inp["a2_cs"], tilt2_cs=inp["t2_cs"])
out["ind_apdfs_1"], out["ind_apdfs_2"], _, out["ind_ctpdfs_1"], out["ind_ctpdfs_2"], _ = r
out["beta_apdfs"], _ = CALC.calculate_beta_spin_mag(inp["alpha_a"], inp["beta_a"], rate=inp["rate"])
out["iso_ctpdfs"], _ = CALC.calculate_mixture_iso_aligned_spin_tilt(inp["sig_ct"], inp["lambda_ct"], pop_frac=inp["frac"])
zm = R["spline_perturbation"].PowerlawSplineRedshiftModel(8, inp["z_pe"], inp["z_inj"])
out["plz_rs"], out["zs"] = CALC.calculate_powerlaw_rate_of_z_ppds(inp["lamb_z"], inp["rate"], zm)
out["plsz_rs"], _ = CALC.calculate_powerlaw_spline_rate_of_z_ppds(inp["lamb_z"], inp["z_cs"], inp["rate"], zm, pop_frac=inp["frac"])
np.savez_compressed(os.path.join(HERE, "ppd_reference.npz"), **{f"in_{k}": np.asarray(v) for k, v in inp.items()}, **{f"out_{k}": np.asarray(v, dtype=np.float64) for k, v in out.items()})
print({k: np.asarray(v).shape for k, v in out.items()})
