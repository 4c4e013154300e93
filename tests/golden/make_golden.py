"""Generate the golden vectors in this directory FROM THE REFERENCE'S OWN CODE.

Run in the build container (needs /root/reference; cannot run on the GPU box):

    python tests/golden/make_golden.py

For each case it builds a small seeded synthetic catalog, constructs the REFERENCE's model classes
(gwinferno/models/**, executed unmodified under oracle/jax_shim.py in fp64), evaluates the
weights exactly like the reference's example/test models do
(examples/simple_bspline_example.py:58-68, tests/inference_test.py:168-172,256-260), reduces them
with the reference's ``per_event_log_bayes_factors`` / ``detection_efficiency`` /
``hierarchical_likelihood`` (gwinferno/pipeline/analysis.py:50-136,139-319) and differentiates
every output with respect to every hyper-parameter by COMPLEX-STEP through that same code
(h = 1e-30, error ~1e-16).  Inputs and outputs are stored in ``<case>.npz``.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import jax_shim  # noqa: E402

R = jax_shim.load_reference()
from gwinferno_b200 import synthetic  # noqa: E402

A = R["analysis"]
SEP = R["separable"]
SPL = R["spline_perturbation"]
PAR = R["parametric"]
INT = R["interpolation"]

COLS = ["mass_1", "mass_ratio", "mass_2", "a_1", "a_2", "cos_tilt_1", "cos_tilt_2", "redshift", "prior"]
EXTRA_COLS = ["chi_eff", "chi_p", "dVdc"]  # derived coordinates, stored only by the cases that use them
SGL = R["single"]
COSMO = R["cosmology"]


def _record_likelihood(pe_w, inj_w, total_inj, Nobs, Tobs, vt, **kw):
    """Drive the reference's hierarchical_likelihood with numpyro primitives stubbed."""
    import numpyro

    rec = {}
    numpyro.deterministic = lambda name, v: rec.setdefault(name, v) if False else rec.__setitem__(name, v) or v
    numpyro.factor = lambda name, v: rec.__setitem__(name, v)
    numpyro.sample = lambda name, *a, **k: 30.0
    A.numpyro = numpyro
    A.hierarchical_likelihood(pe_w, inj_w, total_inj, Nobs, Tobs, surveyed_hypervolume=vt, **kw)
    return rec


def _reduce(weights_fn, pe, inj, const, params, vt_fn, like_kw):
    """All reference outputs as one flat complex/real vector + names (for complex-step)."""
    pw = weights_fn(pe, True, params)
    iw = weights_fn(inj, False, params)
    logBF, logneff, var = A.per_event_log_bayes_factors(pw)
    logmu, logneff_inj, var_inj = A.detection_efficiency(iw, const["total_inj"])
    vt = vt_fn(params)
    rec = _record_likelihood(pw, iw, const["total_inj"], const["nObs"], const["obs_time"], vt, **like_kw)
    return dict(
        logBF=np.asarray(logBF),
        logNeff=np.asarray(logneff),
        var=np.asarray(var),
        log_mu=np.asarray(logmu),
        logNeff_inj=np.asarray(logneff_inj),
        var_inj=np.asarray(var_inj),
        surveyed_hypervolume=np.asarray(vt),
        log_l=np.asarray(rec["log_likelihood"]),
    )


def _complex_step(weights_fn, pe, inj, const, params, names, vt_fn, like_kw):
    """Jacobian of every output wrt the flat parameter vector (order = ``names``)."""
    h = 1e-30
    base = _reduce(weights_fn, pe, inj, const, params, vt_fn, like_kw)
    flat_names = []
    for n in names:
        flat_names += [(n, i) for i in range(np.size(params[n]))]
    jac = {k: np.zeros(np.shape(v) + (len(flat_names),)) for k, v in base.items() if k in ("logBF", "logNeff", "log_mu", "logNeff_inj", "log_l")}
    for col, (n, i) in enumerate(flat_names):
        p2 = {k: (np.array(v, dtype=np.complex128) if k == n else v) for k, v in params.items()}
        if np.ndim(p2[n]) == 0:
            p2[n] = p2[n] + 1j * h
        else:
            p2[n][i] += 1j * h
        out = _reduce(weights_fn, pe, inj, const, p2, vt_fn, like_kw)
        for k in jac:
            jac[k][..., col] = np.imag(out[k]) / h
    return base, jac


def _save(name, pe, inj, const, params, names, base, jac, meta):
    out = {}
    for c in COLS + [c for c in EXTRA_COLS if c in pe]:
        out[f"pe_{c}"] = pe[c]
        out[f"inj_{c}"] = inj[c]
    for k, v in const.items():
        out[f"const_{k}"] = np.float64(v)
    for n in names:
        out[f"param_{n}"] = np.asarray(params[n], dtype=np.float64)
    out["param_order"] = np.array(names)
    for k, v in base.items():
        out[f"out_{k}"] = np.real(v)
    for k, v in jac.items():
        out[f"jac_{k}"] = v
    for k, v in meta.items():
        out[f"meta_{k}"] = np.asarray(v)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)  log_l = {float(np.real(base['log_l'])):.12f}")


# ----------------------------------------------------------------------------------------------
def case_bspline_full(maxvar_only=False):
    """cfg-2/3 model (examples/simple_bspline_example.py): B-spline m1+q, independent spin
    magnitudes and tilts, power-law x spline redshift.  Default spline counts 50/30/16/16/20."""
    E, S, I = 8, 250, 6000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=102)
    ns = dict(m1=50, q=30, a=16, t=16, z=20)
    mmin, mmax = 3.0, 100.0
    rm = SEP.BSplinePrimaryBSplineRatio(
        ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax,
        kwargs_m={"basis": INT.LogXLogYBSpline}, kwargs_q={"basis": INT.LogYBSpline},
    )  # pipeline/utils.py:104-118
    ra = SEP.BSplineIndependentSpinMagnitudes(ns["a"], ns["a"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
    rt = SEP.BSplineIndependentSpinTilts(ns["t"], ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(102002)
    params = dict(
        mass_cs=rng.standard_normal(ns["m1"]), q_cs=rng.standard_normal(ns["q"]),
        a1_cs=rng.standard_normal(ns["a"]), a2_cs=rng.standard_normal(ns["a"]),
        tilt1_cs=rng.standard_normal(ns["t"]), tilt2_cs=rng.standard_normal(ns["t"]),
        lamb=np.float64(2.7), z_cs=rng.standard_normal(ns["z"]),
    )
    params["z_cs"][0] = 0.0
    names = ["mass_cs", "q_cs", "a1_cs", "a2_cs", "tilt1_cs", "tilt2_cs", "lamb", "z_cs"]
    sharp = {k: (np.array(v) if np.ndim(v) else v) for k, v in params.items()}  # N(0,1): fails the N_eff cut
    for k in names:
        if np.ndim(params[k]):
            params[k] = 0.35 * params[k]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ra(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
        w = w * rt(p["tilt1_cs"], p["tilt2_cs"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]
        return w

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    # max_variance_cut=True (analysis.py:309-317; needs marginalize_selection = min_neff_cut = False):
    # one parameter point below the variance threshold of 1, one (the sharp population) above it
    like_kw = dict(marginalize_selection=False, min_neff_cut=False, max_variance_cut=True)
    for nm, pp in (("bspline_full_maxvar", params), ("bspline_full_maxvar_fail", sharp)):
        base, jac = _complex_step(weights, pe, inj, const, pp, names, vt, like_kw)
        var_l = const["nObs"] ** 2 * float(np.real(base["var_inj"])) + float(np.sum(np.real(base["var"])))
        print(f"   ({nm}) variance_log_likelihood = {var_l:.4f}")
        _save(nm, pe, inj, const, pp, names, base, jac, dict(model="bspline_full", mmin=mmin, mmax=mmax, max_variance_cut=1, **{f"n_{k}": v for k, v in ns.items()}))
    if maxvar_only:
        return
    like_kw = dict(marginalize_selection=False, min_neff_cut=True)
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, like_kw)
    _save("bspline_full", pe, inj, const, params, names, base, jac, dict(model="bspline_full", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))
    # same model, marginalize_selection=True (exercises the N_eff,inj gradient)
    like_kw = dict(marginalize_selection=True, min_neff_cut=True)
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, like_kw)
    _save("bspline_full_margsel", pe, inj, const, params, names, base, jac, dict(model="bspline_full", mmin=mmin, mmax=mmax, marginalize_selection=1, **{f"n_{k}": v for k, v in ns.items()}))
    # sharp population: per-event N_eff <= Nobs => the reference returns its -inf sentinel
    like_kw = dict(marginalize_selection=False, min_neff_cut=True)
    base, jac = _complex_step(weights, pe, inj, const, sharp, names, vt, like_kw)
    print("   (cut-fail case) min N_eff =", float(np.exp(np.min(np.real(base["logNeff"])))), " N_eff,inj =", float(np.exp(np.real(base["logNeff_inj"]))))
    _save("bspline_full_cutfail", pe, inj, const, sharp, names, base, jac, dict(model="bspline_full", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_bspline_iid():
    """cfg-5 model: IID spin magnitudes / tilts (shared coefficients), IID component masses with
    pairing (m2/m1)^beta, power-law x spline redshift (separable.py:17-79,156-218,533-613)."""
    E, S, I = 7, 240, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=105)
    ns = dict(m=24, a=10, t=9, z=8)
    mmin, mmax = 3.0, 100.0
    rm = SEP.BSplineIIDComponentMasses(ns["m"], pe["mass_1"], pe["mass_2"], inj["mass_1"], inj["mass_2"], mmin=mmin, mmax=mmax)
    ra = SEP.BSplineIIDSpinMagnitudes(ns["a"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
    rt = SEP.BSplineIIDSpinTilts(ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(105002)
    params = dict(mass_cs=rng.standard_normal(ns["m"]), beta=np.float64(1.3), a_cs=rng.standard_normal(ns["a"]), tilt_cs=rng.standard_normal(ns["t"]), lamb=np.float64(1.9), z_cs=0.5 * rng.standard_normal(ns["z"]))
    params["z_cs"][0] = 0.0
    names = ["mass_cs", "beta", "a_cs", "tilt_cs", "lamb", "z_cs"]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], beta=p["beta"], pe_samples=pe_samples) * ra(p["a_cs"], pe_samples=pe_samples) * rt(p["tilt_cs"], pe_samples=pe_samples)
        return w * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=True))
    _save("bspline_iid", pe, inj, const, params, names, base, jac, dict(model="bspline_iid", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_bspline_indep_masses():
    """cfg-5 variant: independent per-component mass B-splines (separable.py:616-703)."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=115)
    ns = dict(m1=20, m2=14, z=6)
    mmin, mmax = 3.0, 100.0
    rm = SEP.BSplineIndependentComponentMasses(ns["m1"], ns["m2"], pe["mass_1"], pe["mass_2"], inj["mass_1"], inj["mass_2"], mmin1=mmin, mmax1=mmax, mmin2=mmin, mmax2=mmax)
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(115002)
    params = dict(m1_cs=rng.standard_normal(ns["m1"]), m2_cs=rng.standard_normal(ns["m2"]), beta=np.float64(0.7), lamb=np.float64(3.1), z_cs=0.3 * rng.standard_normal(ns["z"]))
    names = ["m1_cs", "m2_cs", "beta", "lamb", "z_cs"]

    def weights(d, pe_samples, p):
        return rm(p["m1_cs"], p["m2_cs"], beta=p["beta"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("bspline_indep_masses", pe, inj, const, params, names, base, jac, dict(model="bspline_indep_masses", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_inference_test_bspline():
    """tests/inference_test.py:98-117,227-263: m1(10)+q(5) B-splines with mmin=5, mmax=100 (PE and
    injection samples below 5 Msun are masked), PowerlawSplineRedshift(5); test point
    m1_coefs~N(0,1), q_coefs~N(0,1), z_coefs=1, lamb=2.9; weights guarded by where(isnan|isinf,0,.)."""
    E, S, I = 9, 100, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=120)
    mmin, mmax = 5.0, 100.0
    rm = SEP.BSplinePrimaryBSplineRatio(10, 5, pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
    rz = SPL.PowerlawSplineRedshiftModel(5, pe["redshift"], inj["redshift"])
    params = dict(m1_coefs=np.random.default_rng(0).standard_normal(10), q_coefs=np.random.default_rng(1).standard_normal(5), lamb=np.float64(2.9), z_coefs=np.ones(5))
    names = ["m1_coefs", "q_coefs", "lamb", "z_coefs"]

    def weights(d, pe_samples, p):
        w = rm(p["m1_coefs"], p["q_coefs"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_coefs"]) / d["prior"]
        return np.where(np.isnan(w) | np.isinf(w), 0, w)

    vt = lambda p: rz.normalization(lamb=p["lamb"], cs=p["z_coefs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(marginalize_selection=False, min_neff_cut=False))
    _save("inference_test_bspline", pe, inj, const, params, names, base, jac, dict(model="inference_test_bspline", mmin=mmin, mmax=mmax))


def case_inference_test_parametric():
    """tests/inference_test.py:140-172: powerlaw_primary_ratio_pdf x PowerlawRedshiftModel at
    alpha=3.5, beta=1.1, lamb=2.9, mmin=5, mmax=100."""
    E, S, I = 9, 100, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=121)
    mmin, mmax = 5.0, 100.0
    rz = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    params = dict(alpha=np.float64(3.5), beta=np.float64(1.1), lamb=np.float64(2.9))
    names = ["alpha", "beta", "lamb"]

    def weights(d, pe_samples, p):
        w = PAR.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax) * rz(d["redshift"], p["lamb"]) / d["prior"]
        return np.where(np.isnan(w) | np.isinf(w), 0, w)

    vt = lambda p: rz.normalization(lamb=p["lamb"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(marginalize_selection=False, min_neff_cut=False))
    _save("inference_test_parametric", pe, inj, const, params, names, base, jac, dict(model="inference_test_parametric", mmin=mmin, mmax=mmax))


def case_plpeak():
    """cfg-1 model (examples/simple_powerlaw_peak_example.py:52-91): PL+Peak primary x power-law
    ratio, independent Beta spin magnitudes, independent iso+aligned tilts, power-law redshift."""
    E, S, I = 8, 250, 6000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=101)
    mmin, mmax = 3.0, 100.0
    rz = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    params = dict(
        alpha=np.float64(-2.8), beta=np.float64(1.4), mu_peak=np.float64(34.0), sig_peak=np.float64(4.5), lambda_m=np.float64(0.08),
        alpha_a1=np.float64(1.8), beta_a1=np.float64(4.2), alpha_a2=np.float64(2.1), beta_a2=np.float64(3.3),
        lambda_ct1=np.float64(0.6), lambda_ct2=np.float64(0.35), sig_ct1=np.float64(1.2), sig_ct2=np.float64(0.8), lamb=np.float64(2.4),
    )
    names = list(params.keys())

    def weights(d, pe_samples, p):
        p_m1q = PAR.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"])
        p_a = PAR.independent_spin_magnitude_beta_dist(d["a_1"], d["a_2"], p["alpha_a1"], p["beta_a1"], p["alpha_a2"], p["beta_a2"])
        p_ct = PAR.independent_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["lambda_ct1"], p["lambda_ct2"], p["sig_ct1"], p["sig_ct2"])
        return p_m1q * p_a * p_ct * rz(d["redshift"], p["lamb"]) / d["prior"]

    vt = lambda p: rz.normalization(lamb=p["lamb"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=True))
    _save("plpeak", pe, inj, const, params, names, base, jac, dict(model="plpeak", mmin=mmin, mmax=mmax))


def case_plpeak_smooth():
    """PL+Peak with the low-mass window ``delta`` (parametric.py:39-53 -> distributions.py:16-21, as
    the reference evaluates it) x power-law redshift."""
    E, S, I = 7, 200, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=141)
    mmin, mmax = 4.0, 100.0
    rz = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    params = dict(alpha=np.float64(-2.3), beta=np.float64(1.1), mu_peak=np.float64(33.0), sig_peak=np.float64(5.0), lambda_m=np.float64(0.1),
                  delta_m=np.float64(4.7), lamb=np.float64(2.6))
    names = list(params.keys())

    def weights(d, pe_samples, p):
        p_m1q = PAR.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"], delta=p["delta_m"])
        return p_m1q * rz(d["redshift"], p["lamb"]) / d["prior"]

    vt = lambda p: rz.normalization(lamb=p["lamb"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("plpeak_smooth", pe, inj, const, params, names, base, jac, dict(model="plpeak_smooth", mmin=mmin, mmax=mmax))


def case_default_spin():
    """PL primary/ratio x IID Beta spin magnitudes x ``default_spin_tilt`` (both tilts aligned together,
    parametric.py:67-68,97-102) x power-law redshift."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=142)
    mmin, mmax = 3.0, 100.0
    rz = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    params = dict(alpha=np.float64(-2.1), beta=np.float64(0.9), alpha_a=np.float64(1.7), beta_a=np.float64(3.9), xi=np.float64(0.55), sig_t=np.float64(0.9), lamb=np.float64(2.8))
    names = list(params.keys())

    def weights(d, pe_samples, p):
        w = PAR.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax)
        w = w * PAR.iid_spin_magnitude(d["a_1"], d["a_2"], p["alpha_a"], p["beta_a"]) * PAR.default_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"])
        return w * rz(d["redshift"], p["lamb"]) / d["prior"]

    vt = lambda p: rz.normalization(lamb=p["lamb"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("default_spin", pe, inj, const, params, names, base, jac, dict(model="default_spin", mmin=mmin, mmax=mmax))


def case_mixed_mass():
    """The two semi-parametric mass models (separable.py:295-365, 368-443):
    ``mixed_bspline_m1``: B-spline primary x power-law ratio; ``mixed_plpeak_m1``: PL+Peak primary x
    B-spline ratio; both x IID spin-tilt mixture x power-law redshift."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=143)
    mmin, mmax = 3.0, 100.0
    rz = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    vt = lambda p: rz.normalization(lamb=p["lamb"])  # noqa: E731
    rng = np.random.default_rng(143002)
    # (1) B-spline primary x power-law ratio
    r1 = SEP.BSplinePrimaryPowerlawRatio(16, pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax)
    params = dict(mass_cs=0.5 * rng.standard_normal(16), beta=np.float64(1.2), xi=np.float64(0.4), sig_t=np.float64(1.1), lamb=np.float64(2.5))
    names = list(params.keys())

    def weights1(d, pe_samples, p):
        w = r1(d["mass_1"], d["mass_ratio"], p["beta"], mmin, p["mass_cs"], pe_samples=pe_samples)
        return w * PAR.iid_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"]) * rz(d["redshift"], p["lamb"]) / d["prior"]

    base, jac = _complex_step(weights1, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("mixed_bspline_m1", pe, inj, const, params, names, base, jac, dict(model="mixed_bspline_m1", mmin=mmin, mmax=mmax, n_m1=16))
    # (2) PL+Peak primary x B-spline ratio
    r2 = SEP.PLPeakPrimaryBSplineRatio(9, pe["mass_ratio"], inj["mass_ratio"])
    params = dict(alpha=np.float64(-2.6), mu_peak=np.float64(35.0), sig_peak=np.float64(4.0), lambda_m=np.float64(0.12), q_cs=0.5 * rng.standard_normal(9),
                  xi=np.float64(0.7), sig_t=np.float64(0.6), lamb=np.float64(3.0))
    names = list(params.keys())

    def weights2(d, pe_samples, p):
        w = r2(d["mass_1"], p["alpha"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"], p["q_cs"], pe_samples=pe_samples)
        return w * PAR.iid_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"]) * rz(d["redshift"], p["lamb"]) / d["prior"]

    base, jac = _complex_step(weights2, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("mixed_plpeak_m1", pe, inj, const, params, names, base, jac, dict(model="mixed_plpeak_m1", mmin=mmin, mmax=mmax, n_q=9))


def case_bspline_redshift_default():
    """``BSplineRedshift`` exactly as the reference constructs it by default (single.py:398-492): the
    LogXBSpline basis is normalised, so the exponent is ``B.c / trapezoid(B.c)`` while ``Z`` uses the
    raw coefficients; masses from the parametric power law."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=144)
    _with_derived(pe, inj)
    mmin, mmax, zmax = 3.0, 100.0, 2.3
    rz = SGL.BSplineRedshift(8, pe["redshift"], inj["redshift"], pe["dVdc"], inj["dVdc"], zmax=zmax)
    rng = np.random.default_rng(144002)
    params = dict(alpha=np.float64(-2.4), beta=np.float64(1.3), z_cs=np.exp(0.4 * rng.standard_normal(8)))
    names = list(params.keys())

    def weights(d, pe_samples, p):
        w = PAR.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax)
        return w * rz(p["z_cs"], pe_samples=pe_samples) / d["prior"]

    vt = lambda p: rz.normalization(p["z_cs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("bspline_redshift_default", pe, inj, const, params, names, base, jac, dict(model="bspline_redshift_default", mmin=mmin, mmax=mmax, zmax=zmax))


def _with_derived(pe, inj):
    for d in (pe, inj):
        d["chi_eff"], d["chi_p"] = synthetic.effective_spins(d)
        d["dVdc"] = np.asarray(COSMO.PLANCK_2015_LVK_Cosmology.dVcdz(d["redshift"]), dtype=np.float64)


def case_bspline_effspin():
    """Effective-spin model: B-spline m1+q, ``BSplineEffectiveSpinDims`` (the chi_eff / chi_p splines
    ARE the densities: default ``BSpline`` basis, single.py:199-230,287-318; separable.py:706-778)
    and ``BSplineRedshift`` (single.py:398-492; ``normalize=False`` so that the exponent is the plain
    spline)."""
    E, S, I = 7, 220, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=130)
    _with_derived(pe, inj)
    ns = dict(m1=14, q=8, e=12, p=9, z=7)
    mmin, mmax, zmax = 3.0, 100.0, 2.3
    rm = SEP.BSplinePrimaryBSplineRatio(ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
    rs = SEP.BSplineEffectiveSpinDims(ns["e"], ns["p"], pe["chi_eff"], pe["chi_p"], inj["chi_eff"], inj["chi_p"], normalize=True)
    rz = SGL.BSplineRedshift(ns["z"], pe["redshift"], inj["redshift"], pe["dVdc"], inj["dVdc"], zmax=zmax, normalize=False)
    rng = np.random.default_rng(130002)
    params = dict(
        mass_cs=0.4 * rng.standard_normal(ns["m1"]), q_cs=0.4 * rng.standard_normal(ns["q"]),
        chieff_cs=np.exp(0.6 * rng.standard_normal(ns["e"])), chip_cs=np.exp(0.6 * rng.standard_normal(ns["p"])),
        z_cs=0.4 * rng.standard_normal(ns["z"]),
    )
    names = ["mass_cs", "q_cs", "chieff_cs", "chip_cs", "z_cs"]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * rs(p["chieff_cs"], p["chip_cs"], pe_samples=pe_samples)
        return w * rz(p["z_cs"], pe_samples=pe_samples) / d["prior"]

    vt = lambda p: rz.normalization(p["z_cs"])  # noqa: E731
    for name, kw in (("bspline_effspin", dict(min_neff_cut=False)), ("bspline_effspin_margsel", dict(marginalize_selection=True, min_neff_cut=False))):
        base, jac = _complex_step(weights, pe, inj, const, params, names, vt, kw)
        _save(name, pe, inj, const, params, names, base, jac, dict(model="bspline_effspin", mmin=mmin, mmax=mmax, zmax=zmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_bspline_symchieff():
    """``BSplineSymmetricChiEffective`` (spline density in |chi_eff|, x 1/2, single.py:233-284), chi_p
    with a LogY basis, the primary mass as a ``LogXBSpline`` DENSITY (linear in y, log in x) and the
    power-law x spline redshift model."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=131)
    _with_derived(pe, inj)
    ns = dict(m1=12, e=8, p=7, z=6)
    mmin, mmax = 3.0, 100.0
    rm = SGL.BSplineMass(ns["m1"], pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax, basis=INT.LogXBSpline)
    re_ = SGL.BSplineSymmetricChiEffective(ns["e"], pe["chi_eff"], inj["chi_eff"], normalize=True)
    rp = SGL.BSplineChiPrecess(ns["p"], pe["chi_p"], inj["chi_p"], basis=INT.LogYBSpline)
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(131002)
    params = dict(
        mass_cs=np.exp(0.8 * rng.standard_normal(ns["m1"])), chieff_cs=np.exp(0.5 * rng.standard_normal(ns["e"])),
        chip_cs=0.5 * rng.standard_normal(ns["p"]), lamb=np.float64(2.2), z_cs=0.3 * rng.standard_normal(ns["z"]),
    )
    names = ["mass_cs", "chieff_cs", "chip_cs", "lamb", "z_cs"]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], pe_samples=pe_samples) * re_(p["chieff_cs"], pe_samples=pe_samples) * rp(p["chip_cs"], pe_samples=pe_samples)
        return w * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("bspline_symchieff", pe, inj, const, params, names, base, jac, dict(model="bspline_symchieff", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


from tests.cases import knots_case_kwargs, knots_density_kwargs  # noqa: E402  (the same numbers build the mirror models)


def case_bspline_knots():
    """Explicit knot vectors and degrees != 3 through every basis flavour (interpolation.py:72-106,128-149): the legal
    API the named configs do not use."""
    E, S, I = 7, 240, 5000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=140)
    mmin, mmax = 3.0, 100.0
    K = knots_case_kwargs(mmin, mmax)
    ns = K["ns"]
    assert len(K["kwargs_t"]["knots"]) == ns["t"] + 4 and len(K["kwargs_m"]["knots"]) == ns["m1"] + 4
    rm = SEP.BSplinePrimaryBSplineRatio(
        ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax,
        kwargs_m={"basis": INT.LogXLogYBSpline, **K["kwargs_m"]}, kwargs_q={"basis": INT.LogYBSpline, **K["kwargs_q"]},
    )
    ra = SEP.BSplineIndependentSpinMagnitudes(ns["a1"], ns["a2"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], kwargs1=K["kwargs_a1"], kwargs2=K["kwargs_a2"], normalize=True)
    rt = SEP.BSplineIIDSpinTilts(ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True, **K["kwargs_t"])
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(140002)
    params = dict(
        mass_cs=0.4 * rng.standard_normal(ns["m1"]), q_cs=0.4 * rng.standard_normal(ns["q"]), a1_cs=0.4 * rng.standard_normal(ns["a1"]),
        a2_cs=0.4 * rng.standard_normal(ns["a2"]), tilt_cs=0.4 * rng.standard_normal(ns["t"]), lamb=np.float64(2.4), z_cs=0.3 * rng.standard_normal(ns["z"]),
    )
    names = ["mass_cs", "q_cs", "a1_cs", "a2_cs", "tilt_cs", "lamb", "z_cs"]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ra(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
        return w * rt(p["tilt_cs"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    for name, kw in (("bspline_knots", dict(min_neff_cut=False)), ("bspline_knots_margsel", dict(marginalize_selection=True, min_neff_cut=False))):
        base, jac = _complex_step(weights, pe, inj, const, params, names, vt, kw)
        _save(name, pe, inj, const, params, names, base, jac, dict(model="bspline_knots", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_bspline_knots_density():
    """Spline DENSITIES (BSpline / LogXBSpline bases: the spline is the pdf, interpolation.py:293-317) on explicit knots and
    with degree 2: the linear normaliser (BSpline.norm, :280-291) over per-piece polynomials."""
    E, S, I = 6, 200, 4000
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=141)
    _with_derived(pe, inj)
    ns = dict(m1=11, e=8, p=7, z=6)
    mmin, mmax = 3.0, 100.0
    K = knots_density_kwargs(mmin, mmax)
    rm = SGL.BSplineMass(ns["m1"], pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax, basis=INT.LogXBSpline, **K["kwargs_m"])
    re_ = SGL.BSplineChiEffective(ns["e"], pe["chi_eff"], inj["chi_eff"], normalize=True, **K["kwargs_e"])
    rp = SGL.BSplineChiPrecess(ns["p"], pe["chi_p"], inj["chi_p"], basis=INT.LogYBSpline)
    rz = SPL.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])
    rng = np.random.default_rng(141002)
    params = dict(
        mass_cs=np.exp(0.8 * rng.standard_normal(ns["m1"])), chieff_cs=np.exp(0.5 * rng.standard_normal(ns["e"])),
        chip_cs=0.5 * rng.standard_normal(ns["p"]), lamb=np.float64(2.2), z_cs=0.3 * rng.standard_normal(ns["z"]),
    )
    names = ["mass_cs", "chieff_cs", "chip_cs", "lamb", "z_cs"]

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], pe_samples=pe_samples) * re_(p["chieff_cs"], pe_samples=pe_samples) * rp(p["chip_cs"], pe_samples=pe_samples)
        return w * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    vt = lambda p: rz.normalization(p["lamb"], p["z_cs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(min_neff_cut=False))
    _save("bspline_knots_density", pe, inj, const, params, names, base, jac, dict(model="bspline_knots_density", mmin=mmin, mmax=mmax, **{f"n_{k}": v for k, v in ns.items()}))


def case_inference_test_gwtc3():
    """The reference's OWN test scenario on its OWN data (tests/inference_test.py:74-117,140-172,227-263): the vendored GWTC-3
    posterior-sample file (69 events x 1000 samples, read here with gwinferno_b200/catalog_io.py -- xarray is not installed),
    100 samples per event drawn without replacement (:78-82; seeded here), the B-spline and the power-law models of that
    test at its parameter points.  The found injections of the reference test (tests/data/injections.h5) are not vendored:
    synthetic ones stand in."""
    from gwinferno_b200 import catalog_io

    path = os.path.join(jax_shim.REFERENCE_ROOT, "tests", "data", "xarray_GWTC3_BBH_69evs_downsampled_1000samps_nospin.h5")
    pe, events, _ = catalog_io.load_pe_dataset(path, n_samples=100, rng=np.random.default_rng(2021))
    inj = synthetic.make_injections(6000, 150001)
    const = {"total_inj": float(4 * 6000), "obs_time": 1.0, "nObs": len(events)}
    mmin, mmax = 5.0, 100.0
    # B-spline model (:98-117, 227-263)
    rm = SEP.BSplinePrimaryBSplineRatio(10, 5, pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
    rz = SPL.PowerlawSplineRedshiftModel(5, pe["redshift"], inj["redshift"])
    params = dict(m1_coefs=np.random.default_rng(0).standard_normal(10), q_coefs=np.random.default_rng(1).standard_normal(5), lamb=np.float64(2.9), z_coefs=np.ones(5))
    names = ["m1_coefs", "q_coefs", "lamb", "z_coefs"]

    def weights(d, pe_samples, p):
        w = rm(p["m1_coefs"], p["q_coefs"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_coefs"]) / d["prior"]
        return np.where(np.isnan(w) | np.isinf(w), 0, w)

    vt = lambda p: rz.normalization(lamb=p["lamb"], cs=p["z_coefs"])  # noqa: E731
    base, jac = _complex_step(weights, pe, inj, const, params, names, vt, dict(marginalize_selection=False, min_neff_cut=False))
    _save("inference_test_bspline_gwtc3", pe, inj, const, params, names, base, jac, dict(model="inference_test_bspline", mmin=mmin, mmax=mmax))
    # power-law model (:140-172)
    rz2 = PAR.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])
    params = dict(alpha=np.float64(3.5), beta=np.float64(1.1), lamb=np.float64(2.9))
    names = ["alpha", "beta", "lamb"]

    def weights2(d, pe_samples, p):
        w = PAR.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax) * rz2(d["redshift"], p["lamb"]) / d["prior"]
        return np.where(np.isnan(w) | np.isinf(w), 0, w)

    vt2 = lambda p: rz2.normalization(lamb=p["lamb"])  # noqa: E731
    base, jac = _complex_step(weights2, pe, inj, const, params, names, vt2, dict(marginalize_selection=False, min_neff_cut=False))
    _save("inference_test_parametric_gwtc3", pe, inj, const, params, names, base, jac, dict(model="inference_test_parametric", mmin=mmin, mmax=mmax))


CASES = dict(
    bspline_full=case_bspline_full,
    bspline_full_maxvar=lambda: case_bspline_full(maxvar_only=True),
    bspline_iid=case_bspline_iid,
    bspline_indep_masses=case_bspline_indep_masses,
    inference_test_bspline=case_inference_test_bspline,
    inference_test_parametric=case_inference_test_parametric,
    plpeak=case_plpeak,
    bspline_effspin=case_bspline_effspin,
    bspline_symchieff=case_bspline_symchieff,
    plpeak_smooth=case_plpeak_smooth,
    default_spin=case_default_spin,
    mixed_mass=case_mixed_mass,
    bspline_redshift_default=case_bspline_redshift_default,
    bspline_knots=case_bspline_knots,
    bspline_knots_density=case_bspline_knots_density,
    inference_test_gwtc3=case_inference_test_gwtc3,
)

if __name__ == "__main__":
    import warnings

    warnings.filterwarnings("ignore")
    for name in sys.argv[1:] or list(CASES):  # optional: only the named cases
        CASES[name]()
