"""Shared test helpers: rebuild the golden cases (tests/golden/*.npz) with THIS repo's mirror of
the reference model classes, giving the lazy weights, the lowered spec and the flat Lambda."""

import os

import numpy as np

from gwinferno_b200 import lowering
from gwinferno_b200 import models as M

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COLS = ["mass_1", "mass_ratio", "mass_2", "a_1", "a_2", "cos_tilt_1", "cos_tilt_2", "redshift", "prior"]
EXTRA_COLS = ["chi_eff", "chi_p", "dVdc"]  # derived coordinates stored by the cases that use them
GOLDEN_CASES = [
    "bspline_full",
    "bspline_full_margsel",
    "bspline_full_cutfail",
    "bspline_full_maxvar",
    "bspline_full_maxvar_fail",
    "bspline_iid",
    "bspline_indep_masses",
    "inference_test_bspline",
    "inference_test_parametric",
    "plpeak",
    "bspline_effspin",
    "bspline_effspin_margsel",
    "bspline_symchieff",
    "plpeak_smooth",
    "default_spin",
    "mixed_bspline_m1",
    "mixed_plpeak_m1",
    "bspline_knots",
    "bspline_knots_margsel",
    "bspline_knots_density",
    "inference_test_bspline_gwtc3",
    "inference_test_parametric_gwtc3",
]
# Cases whose model carries host-side glue (a parameter map / a host normaliser, lowering.pull_back and
# lowering.host_log_norm): the DEVICE model is an ordinary one, so they are checked on the CPU with the
# oracle standing in for the device (tests/test_host_glue_cpu.py), not in the GPU-parametrised list.
HOST_GLUE_CASES = ["bspline_redshift_default"]
LIKE_KW = {
    "bspline_full": dict(marginalize_selection=False, min_neff_cut=True),
    "bspline_full_margsel": dict(marginalize_selection=True, min_neff_cut=True),
    "bspline_full_cutfail": dict(marginalize_selection=False, min_neff_cut=True),
    "bspline_full_maxvar": dict(marginalize_selection=False, min_neff_cut=False, max_variance_cut=True),
    "bspline_full_maxvar_fail": dict(marginalize_selection=False, min_neff_cut=False, max_variance_cut=True),
    "bspline_iid": dict(min_neff_cut=True),
    "bspline_indep_masses": dict(min_neff_cut=False),
    "inference_test_bspline": dict(marginalize_selection=False, min_neff_cut=False),
    "inference_test_parametric": dict(marginalize_selection=False, min_neff_cut=False),
    "plpeak": dict(min_neff_cut=True),
    "bspline_effspin": dict(min_neff_cut=False),
    "bspline_effspin_margsel": dict(marginalize_selection=True, min_neff_cut=False),
    "bspline_symchieff": dict(min_neff_cut=False),
    "plpeak_smooth": dict(min_neff_cut=False),
    "default_spin": dict(min_neff_cut=False),
    "mixed_bspline_m1": dict(min_neff_cut=False),
    "mixed_plpeak_m1": dict(min_neff_cut=False),
    "bspline_redshift_default": dict(min_neff_cut=False),
    "bspline_knots": dict(min_neff_cut=False),
    "bspline_knots_margsel": dict(marginalize_selection=True, min_neff_cut=False),
    "bspline_knots_density": dict(min_neff_cut=False),
    "inference_test_bspline_gwtc3": dict(marginalize_selection=False, min_neff_cut=False),
    "inference_test_parametric_gwtc3": dict(marginalize_selection=False, min_neff_cut=False),
}


# knot vectors / degrees of the "bspline_knots" case; tests/cases.py rebuilds the mirror models from the same numbers
def knots_case_kwargs(mmin=3.0, mmax=100.0):
    n_m1, n_q, n_a1, n_a2, n_t = 13, 9, 8, 7, 10
    # (a) explicit NON-UNIFORM knot vector for the mass (given in x units: the log-x bases take its logarithm,
    #     interpolation.py:425-428): clamped cubic knots, denser at low mass
    inner = mmin * (mmax / mmin) ** (np.linspace(0.0, 1.0, n_m1 - 2) ** 1.6)
    m_knots = np.concatenate([[mmin * 0.7, mmin * 0.8, mmin * 0.9], inner, [mmax * 1.1, mmax * 1.2, mmax * 1.3]])
    # (b) quadratic basis on default knots for q; (c) `interior_knots=` for a_1 (only their count and first spacing
    #     matter, interpolation.py:99-101: the basis does NOT sum to one near the upper end); (d) linear basis for a_2;
    # (e) a cubic vector with a repeated knot for the tilts (C^1 only there)
    a1_interior = np.array([0.0, 0.08, 0.2, 0.45, 0.7, 1.0])
    t_knots = np.concatenate([[-1.6, -1.4, -1.2], [-1.0, -0.5, 0.0, 0.0, 0.4, 0.8, 1.0], [1.2, 1.4, 1.6, 1.8]])
    return dict(
        ns=dict(m1=n_m1, q=n_q, a1=n_a1, a2=n_a2, t=n_t, z=6),
        kwargs_m=dict(knots=m_knots), kwargs_q=dict(degree=2), kwargs_a1=dict(interior_knots=a1_interior), kwargs_a2=dict(degree=1), kwargs_t=dict(knots=t_knots),
    )


def knots_density_kwargs(mmin=3.0, mmax=100.0):
    """Knots of the "bspline_knots_density" case: a non-uniform cubic vector for the LogXBSpline mass density (in x units),
    a quadratic chi_eff density on default knots."""
    inner = mmin * (mmax / mmin) ** (np.linspace(0.0, 1.0, 9) ** 0.7)
    return dict(kwargs_m=dict(knots=np.concatenate([[2.0, 2.3, 2.6], inner, [120.0, 140.0, 160.0]])), kwargs_e=dict(degree=2))


class Case:
    pass


def build_weight_fn(model, pe, inj, meta):
    """Return ``weights(datadict, pe_samples, params) -> LazyWeight`` written like the reference's
    example models (examples/simple_bspline_example.py:58-68 etc.)."""
    mmin, mmax = float(meta["mmin"]), float(meta["mmax"])
    if model == "bspline_full":
        mm = M.BSplinePrimaryBSplineRatio(
            int(meta["n_m1"]), int(meta["n_q"]), pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax,
            kwargs_m={"basis": M.LogXLogYBSpline}, kwargs_q={"basis": M.LogYBSpline},
        )
        ma = M.BSplineIndependentSpinMagnitudes(int(meta["n_a"]), int(meta["n_a"]), pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
        mt = M.BSplineIndependentSpinTilts(int(meta["n_t"]), int(meta["n_t"]), pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
        mz = M.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ma(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
            w = w * mt(p["tilt1_cs"], p["tilt2_cs"], pe_samples=pe_samples) * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]
            return w

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    if model == "bspline_iid":
        mm = M.BSplineIIDComponentMasses(int(meta["n_m"]), pe["mass_1"], pe["mass_2"], inj["mass_1"], inj["mass_2"], mmin=mmin, mmax=mmax)
        ma = M.BSplineIIDSpinMagnitudes(int(meta["n_a"]), pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
        mt = M.BSplineIIDSpinTilts(int(meta["n_t"]), pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
        mz = M.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], beta=p["beta"], pe_samples=pe_samples) * ma(p["a_cs"], pe_samples=pe_samples) * mt(p["tilt_cs"], pe_samples=pe_samples)
            return w * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    if model == "bspline_indep_masses":
        mm = M.BSplineIndependentComponentMasses(int(meta["n_m1"]), int(meta["n_m2"]), pe["mass_1"], pe["mass_2"], inj["mass_1"], inj["mass_2"], mmin1=mmin, mmax1=mmax, mmin2=mmin, mmax2=mmax)
        mz = M.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            return mm(p["m1_cs"], p["m2_cs"], beta=p["beta"], pe_samples=pe_samples) * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    if model == "inference_test_bspline":
        mm = M.BSplinePrimaryBSplineRatio(10, 5, pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
        mz = M.PowerlawSplineRedshiftModel(5, pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            return mm(p["m1_coefs"], p["q_coefs"], pe_samples=pe_samples) * mz(d["redshift"], p["lamb"], p["z_coefs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"], cs=p["z_coefs"]))
    if model == "inference_test_parametric":
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            return M.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax) * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "plpeak":
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            p_m1q = M.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"])
            p_a = M.independent_spin_magnitude_beta_dist(d["a_1"], d["a_2"], p["alpha_a1"], p["beta_a1"], p["alpha_a2"], p["beta_a2"])
            p_ct = M.independent_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["lambda_ct1"], p["lambda_ct2"], p["sig_ct1"], p["sig_ct2"])
            return p_m1q * p_a * p_ct * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "plpeak_smooth":
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            p_m1q = M.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"], delta=p["delta_m"])
            return p_m1q * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "default_spin":
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            w = M.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax)
            w = w * M.iid_spin_magnitude(d["a_1"], d["a_2"], p["alpha_a"], p["beta_a"]) * M.default_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"])
            return w * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "mixed_bspline_m1":
        mm = M.BSplinePrimaryPowerlawRatio(int(meta["n_m1"]), pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax)
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(d["mass_1"], d["mass_ratio"], p["beta"], mmin, p["mass_cs"], pe_samples=pe_samples)
            return w * M.iid_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"]) * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "mixed_plpeak_m1":
        mm = M.PLPeakPrimaryBSplineRatio(int(meta["n_q"]), pe["mass_ratio"], inj["mass_ratio"])
        mz = M.PowerlawRedshiftModel(z_pe=pe["redshift"], z_inj=inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(d["mass_1"], p["alpha"], mmin, mmax, p["mu_peak"], p["sig_peak"], p["lambda_m"], p["q_cs"], pe_samples=pe_samples)
            return w * M.iid_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["sig_t"]) * mz(d["redshift"], p["lamb"]) / d["prior"]

        return weights, (lambda p: mz.normalization(lamb=p["lamb"]))
    if model == "bspline_redshift_default":
        mz = M.BSplineRedshift(8, pe["redshift"], inj["redshift"], pe["dVdc"], inj["dVdc"], zmax=float(meta["zmax"]))

        def weights(d, pe_samples, p):
            w = M.powerlaw_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], alpha=p["alpha"], beta=p["beta"], mmin=mmin, mmax=mmax)
            return w * mz(p["z_cs"], pe_samples=pe_samples) / d["prior"]

        return weights, (lambda p: mz.normalization(p["z_cs"]))
    if model == "bspline_effspin":
        mm = M.BSplinePrimaryBSplineRatio(int(meta["n_m1"]), int(meta["n_q"]), pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
        ms = M.BSplineEffectiveSpinDims(int(meta["n_e"]), int(meta["n_p"]), pe["chi_eff"], pe["chi_p"], inj["chi_eff"], inj["chi_p"], normalize=True)
        mz = M.BSplineRedshift(int(meta["n_z"]), pe["redshift"], inj["redshift"], pe["dVdc"], inj["dVdc"], zmax=float(meta["zmax"]), normalize=False)

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ms(p["chieff_cs"], p["chip_cs"], pe_samples=pe_samples)
            return w * mz(p["z_cs"], pe_samples=pe_samples) / d["prior"]

        return weights, (lambda p: mz.normalization(p["z_cs"]))
    if model == "bspline_symchieff":
        mm = M.BSplineMass(int(meta["n_m1"]), pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax, basis=M.LogXBSpline)
        me = M.BSplineSymmetricChiEffective(int(meta["n_e"]), pe["chi_eff"], inj["chi_eff"], normalize=True)
        mp = M.BSplineChiPrecess(int(meta["n_p"]), pe["chi_p"], inj["chi_p"], basis=M.LogYBSpline)
        mz = M.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], pe_samples=pe_samples) * me(p["chieff_cs"], pe_samples=pe_samples) * mp(p["chip_cs"], pe_samples=pe_samples)
            return w * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    if model == "bspline_knots":
        K = knots_case_kwargs(mmin, mmax)
        ns = K["ns"]
        mm = M.BSplinePrimaryBSplineRatio(
            ns["m1"], ns["q"], pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax,
            kwargs_m={"basis": M.LogXLogYBSpline, **K["kwargs_m"]}, kwargs_q={"basis": M.LogYBSpline, **K["kwargs_q"]},
        )
        ma = M.BSplineIndependentSpinMagnitudes(ns["a1"], ns["a2"], pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], kwargs1=K["kwargs_a1"], kwargs2=K["kwargs_a2"], normalize=True)
        mt = M.BSplineIIDSpinTilts(ns["t"], pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True, **K["kwargs_t"])
        mz = M.PowerlawSplineRedshiftModel(ns["z"], pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ma(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
            return w * mt(p["tilt_cs"], pe_samples=pe_samples) * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    if model == "bspline_knots_density":
        K = knots_density_kwargs(mmin, mmax)
        mm = M.BSplineMass(int(meta["n_m1"]), pe["mass_1"], inj["mass_1"], mmin=mmin, mmax=mmax, basis=M.LogXBSpline, **K["kwargs_m"])
        me = M.BSplineChiEffective(int(meta["n_e"]), pe["chi_eff"], inj["chi_eff"], normalize=True, **K["kwargs_e"])
        mp = M.BSplineChiPrecess(int(meta["n_p"]), pe["chi_p"], inj["chi_p"], basis=M.LogYBSpline)
        mz = M.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

        def weights(d, pe_samples, p):
            w = mm(p["mass_cs"], pe_samples=pe_samples) * me(p["chieff_cs"], pe_samples=pe_samples) * mp(p["chip_cs"], pe_samples=pe_samples)
            return w * mz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

        return weights, (lambda p: mz.normalization(p["lamb"], p["z_cs"]))
    raise KeyError(model)


def load_case(name):
    """Load a golden case and lower it with the mirror classes."""
    d = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    c = Case()
    c.name = name
    cols = COLS + [k for k in EXTRA_COLS if f"pe_{k}" in d.files]
    c.pe = {k: d[f"pe_{k}"] for k in cols}
    c.inj = {k: d[f"inj_{k}"] for k in cols}
    c.total_inj = float(d["const_total_inj"])
    c.Nobs = int(d["const_nObs"])
    c.Tobs = float(d["const_obs_time"])
    c.names = [str(n) for n in d["param_order"]]
    c.params = {n: (np.array(d[f"param_{n}"]) if d[f"param_{n}"].ndim else np.float64(d[f"param_{n}"])) for n in c.names}
    c.meta = {k[5:]: d[k] for k in d.files if k.startswith("meta_")}
    c.out = {k[4:]: d[k] for k in d.files if k.startswith("out_")}
    c.jac = {k[4:]: d[k] for k in d.files if k.startswith("jac_")}
    c.like_kw = LIKE_KW[name]
    model = str(c.meta["model"])
    c.weights, c.vt = build_weight_fn(model, c.pe, c.inj, c.meta)
    c.pe_w = c.weights(c.pe, True, c.params)
    c.inj_w = c.weights(c.inj, False, c.params)
    c.low = lowering.lower(c.pe_w, c.inj_w)
    c.Lam = lowering.flatten_params(c.pe_w, c.low.spec.n_params)
    # golden Jacobian columns follow ``names`` flattened; map them to Lambda slots
    perm = []
    for n in c.names:
        sl = c.low.slots_for(c.params[n])
        perm += list(range(sl.start, sl.stop))
    c.golden_to_lambda = np.array(perm)
    return c


def golden_jac_in_lambda_order(c, key):
    """Golden Jacobian ``key`` re-ordered to this repo's Lambda layout (shape (..., P))."""
    J = c.jac[key]
    out = np.zeros(J.shape[:-1] + (c.low.spec.n_params,))
    out[..., c.golden_to_lambda] = J
    return out
