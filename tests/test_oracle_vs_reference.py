"""CPU, build container only: the oracle against the reference's OWN code executed live under
oracle/jax_shim.py at randomised hyper-parameters (the committed golden vectors are fixed points
of the same comparison).  Skipped where /root/reference does not exist (the GPU box)."""

import os

import numpy as np
import pytest

from oracle import jax_shim, popmodel
from tests import cases

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(jax_shim.REFERENCE_ROOT, "gwinferno")), reason="reference tree not mounted")


def _ref_models(R, c):
    SEP, SPL, INT = R["separable"], R["spline_perturbation"], R["interpolation"]
    pe, inj, meta = c.pe, c.inj, c.meta
    mmin, mmax = float(meta["mmin"]), float(meta["mmax"])
    rm = SEP.BSplinePrimaryBSplineRatio(int(meta["n_m1"]), int(meta["n_q"]), pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax,
                                        kwargs_m={"basis": INT.LogXLogYBSpline}, kwargs_q={"basis": INT.LogYBSpline})
    ra = SEP.BSplineIndependentSpinMagnitudes(int(meta["n_a"]), int(meta["n_a"]), pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
    rt = SEP.BSplineIndependentSpinTilts(int(meta["n_t"]), int(meta["n_t"]), pe["cos_tilt_1"], pe["cos_tilt_2"], inj["cos_tilt_1"], inj["cos_tilt_2"], normalize=True)
    rz = SPL.PowerlawSplineRedshiftModel(int(meta["n_z"]), pe["redshift"], inj["redshift"])

    def weights(d, pe_samples, p):
        w = rm(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * ra(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
        return w * rt(p["tilt1_cs"], p["tilt2_cs"], pe_samples=pe_samples) * rz(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"]

    return weights


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_points_against_reference_code(seed):
    R = jax_shim.load_reference()
    A = R["analysis"]
    c = cases.load_case("bspline_full")
    ref_w = _ref_models(R, c)
    rng = np.random.default_rng(seed)
    p = {k: (0.5 * rng.standard_normal(np.shape(v)) if np.ndim(v) else np.float64(rng.uniform(-1, 4))) for k, v in c.params.items()}
    pw, iw = np.asarray(ref_w(c.pe, True, p)), np.asarray(ref_w(c.inj, False, p))
    logBF, logneff, _ = A.per_event_log_bayes_factors(pw)
    logmu, logneff_inj, _ = A.detection_efficiency(iw, c.total_inj)
    from gwinferno_b200 import lowering

    pe_w = c.weights(c.pe, True, p)
    lam = lowering.flatten_params(pe_w, c.low.spec.n_params)
    ev = popmodel.evaluate(c.low.spec, c.low.pe_cols, c.low.inj_cols, c.total_inj, lam, want_jac=False)
    assert np.max(np.abs(ev["logBF"] - np.asarray(logBF))) < 1e-11
    assert np.max(np.abs(ev["logNeff"] - np.asarray(logneff))) < 1e-11
    assert abs(ev["log_mu"] - float(logmu)) < 1e-11 and abs(ev["logNeff_inj"] - float(logneff_inj)) < 1e-11
    # log-space branch of the reference reducers (analysis.py:76-80,123-127) agrees too
    with np.errstate(divide="ignore"):
        lb2, ln2, _ = A.per_event_log_bayes_factors(np.log(pw), log=True)
    assert np.max(np.abs(ev["logBF"] - np.asarray(lb2))) < 1e-10


def test_reference_component_known_answers():
    """The reference's own component pins (tests/interpolation_test.py:50-55,
    tests/distributions_test.py:30-88) hold for the oracle's closed forms."""
    from scipy.interpolate import BSpline as SciBSpline
    from scipy.stats import beta as sbeta
    from scipy.stats import truncnorm

    n = 10
    x = np.linspace(0, 1, 2001)
    j, w, inside = popmodel.spline_taps(x, 0.0, 1.0, n)
    n_int = n - 2
    dx = 1.0 / (n_int - 1)
    knots = np.linspace(-3 * dx, 1 + 3 * dx, n_int + 6)
    dense = np.zeros((n, x.size))
    for k in range(4):
        dense[j + k, np.arange(x.size)] += w[:, k]
    ref = SciBSpline(knots, np.eye(n), 3)(x).T
    assert np.max(np.abs(dense[:, :-1] - ref[:, :-1])) < 1e-14
    xs = np.linspace(-0.9, 0.9, 50)
    lp, _, _ = popmodel._truncnorm_logpdf(xs, 0.3, 0.7, -1.0, 1.0)
    assert np.allclose(np.exp(lp), truncnorm.pdf(xs, (-1 - 0.3) / 0.7, (1 - 0.3) / 0.7, loc=0.3, scale=0.7), rtol=1e-12)
    a = np.linspace(0.05, 0.95, 30)
    f = (2.5 - 1) * np.log(a) + (4.0 - 1) * np.log(1 - a) - popmodel.betaln(2.5, 4.0)
    assert np.allclose(np.exp(f), sbeta.pdf(a, 2.5, 4.0), rtol=1e-12)
