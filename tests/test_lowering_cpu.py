"""CPU: the reference-style Python surface (gwinferno_b200/models.py) and its lowering to the
C-ABI description -- host logic only, no GPU."""

import numpy as np
import pytest

from gwinferno_b200 import lowering
from gwinferno_b200 import models as M
from gwinferno_b200 import spec as S
from gwinferno_b200 import synthetic


@pytest.fixture(scope="module")
def cat():
    return synthetic.make_catalog(4, 50, 400, cfg=301)


def test_iid_models_share_coefficient_slots(cat):
    pe, inj, _ = cat
    mag = M.BSplineIIDSpinMagnitudes(8, pe["a_1"], pe["a_2"], inj["a_1"], inj["a_2"], normalize=True)
    c = np.linspace(-1, 1, 8)
    low = lowering.lower(mag(c, pe_samples=True), mag(c, pe_samples=False))
    assert low.spec.n_params == 8  # one coefficient vector for two columns (separable.py:77-79)
    assert [t.slots[0] for t in low.spec.terms] == [0, 0]
    assert len(low.spec.groups) == 2 and len(low.pe_cols) == 2
    assert low.slots_for(c) == slice(0, 8)


def test_independent_models_get_distinct_slots_and_norm_groups(cat):
    pe, inj, _ = cat
    mass = M.BSplinePrimaryBSplineRatio(10, 6, pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=3.0, m2min=3.0, mmax=100.0)
    z = M.PowerlawSplineRedshiftModel(5, pe["redshift"], inj["redshift"])
    cm, cq, cz, lamb = np.zeros(10), np.zeros(6), np.zeros(5), np.float64(2.0)

    def w(d, ps):
        return mass(cm, cq, pe_samples=ps) * z(d["redshift"], lamb, cz) / d["prior"]

    low = lowering.lower(w(pe, True), w(inj, False))
    assert low.spec.n_params == 10 + 6 + 1 + 5
    kinds = [t.kind for t in low.spec.terms]
    assert kinds.count(S.TERM_SPLINE) == 3 and kinds.count(S.TERM_LINEAR) == 1 and kinds.count(S.TERM_STATIC) == 2
    # the redshift spline and the (1+z)^(lamb-1) term share one norm group (spline_perturbation.py:323-336)
    zterms = [t for t in low.spec.terms if t.cols == low.spec.terms[-2].cols and t.norm_group >= 0]
    assert len({t.norm_group for t in zterms}) == 1 and len(zterms) == 2
    # z <= zmax cut (spline_perturbation.py:368-372) and the data-derived range (parametric.py:114-115)
    assert len(low.spec.cuts) == 1 and low.spec.cuts[0].hi == z.zmax
    assert z.zmin == max(pe["redshift"].min(), inj["redshift"].min())
    # LogXLogY mass grid: 1500 points linear in m, LogY ratio grid: 1000 points (interpolation.py:378,433)
    assert mass.primary_model.grid.size == 1500 and mass.ratio_model.grid.size == 1000
    lam = lowering.flatten_params(w(pe, True), low.spec.n_params)
    assert lam[low.slots_for(lamb)] == 2.0


def test_weight_products_must_match(cat):
    pe, inj, _ = cat
    a = M.BSplineSpinMagnitude(8, pe["a_1"], inj["a_1"], normalize=True)
    b = M.BSplineSpinTilt(8, pe["cos_tilt_1"], inj["cos_tilt_1"], normalize=True)
    c = np.zeros(8)
    with pytest.raises(ValueError):
        lowering.lower(a(c, pe_samples=True) * b(c, pe_samples=True), a(c, pe_samples=False))
    with pytest.raises(ValueError):
        lowering.lower(a(c, pe_samples=False), a(c, pe_samples=True))  # swapped sample sets
    with pytest.raises(ValueError):
        a(np.zeros(7), pe_samples=True)  # wrong number of coefficients
    with pytest.raises(ValueError):
        _ = a(c, pe_samples=True) * b(c, pe_samples=False)  # mixing PE and injection weights


def test_unsupported_reference_options_fail_loudly(cat):
    pe, inj, _ = cat
    with pytest.raises(NotImplementedError):
        M.BSplineRedshift(8, pe["redshift"], inj["redshift"], pe["redshift"], inj["redshift"], basis=M.LogYBSpline)
    with pytest.raises(NotImplementedError):
        M.BSplineSpinMagnitude(8, pe["a_1"], inj["a_1"], degree=4)  # 5 coefficients per piece: not on the CUDA path
    with pytest.raises(AssertionError):
        M.BSplineSpinMagnitude(8, pe["a_1"], inj["a_1"], knots=np.linspace(0, 1, 11))  # interpolation.py:106
    with pytest.raises(NotImplementedError):
        M.BSplineRedshift(8, pe["redshift"], inj["redshift"], pe["redshift"], inj["redshift"], degree=2)
    m = M.BSplineSpinMagnitude(8, pe["a_1"], inj["a_1"], degree=2)  # explicit knots / other degrees lower to a knot vector
    assert m.order == 3 and len(m.knots) == 8 + 3
    m = M.BSplineMass(9, pe["mass_1"], inj["mass_1"], mmin=3, mmax=100, knots=np.geomspace(2.0, 150.0, 13))
    assert np.allclose(m.knots, np.log(np.geomspace(2.0, 150.0, 13)))  # log-x bases take the log of user knots (:425-428)


def test_parametric_free_functions_lower_to_parametric_terms(cat):
    pe, inj, _ = cat
    z = M.PowerlawRedshiftModel(pe["redshift"], inj["redshift"])
    p = dict(alpha=np.float64(-2.5), beta=np.float64(1.0), mu=np.float64(30.0), sig=np.float64(5.0), lam=np.float64(0.1), xi=np.float64(0.5), st=np.float64(1.0),
             aa=np.float64(2.0), ab=np.float64(3.0), lamb=np.float64(2.0))

    def w(d):
        return (M.plpeak_primary_ratio_pdf(d["mass_1"], d["mass_ratio"], p["alpha"], p["beta"], 3.0, 100.0, p["mu"], p["sig"], p["lam"])
                * M.iid_spin_magnitude(d["a_1"], d["a_2"], p["aa"], p["ab"]) * M.iid_spin_tilt(d["cos_tilt_1"], d["cos_tilt_2"], p["xi"], p["st"])
                * z(d["redshift"], p["lamb"]) / d["prior"])

    low = lowering.lower(w(pe), w(inj))
    kinds = [t.kind for t in low.spec.terms]
    assert kinds.count(S.TERM_PLPEAK) == 1 and kinds.count(S.TERM_POWERLAW_RATIO) == 1 and kinds.count(S.TERM_BETA) == 2 and kinds.count(S.TERM_ISOALIGN) == 2
    # iid models: both columns share the parameter objects => 10 parameters in total
    assert low.spec.n_params == 10
    beta_terms = [t for t in low.spec.terms if t.kind == S.TERM_BETA]
    assert beta_terms[0].slots == beta_terms[1].slots


def test_sharded_catalog_is_an_exact_partition():
    from gwinferno_b200 import workloads

    full = synthetic.make_injections(500_000, 2001)
    key = lambda d: np.sort(d["mass_1"] * 1e3 + d["redshift"])  # noqa: E731
    for by in ("bucket", "index"):
        parts = [workloads.shard_catalog("cfg2", r, 4, all_reduce_minmax=lambda a, b: (a, b), shard_by=by)[1] for r in range(4)]
        assert sum(p["mass_1"].size for p in parts) == 500_000
        assert np.array_equal(np.sort(np.concatenate([key(p) for p in parts])), key(full))
    ev = [workloads.shard_catalog("cfg2", r, 4, all_reduce_minmax=lambda a, b: (a, b))[0]["mass_1"].shape[0] for r in range(4)]
    assert sum(ev) == 70


def test_spline_density_bases_lower_to_linear_spline_terms(cat):
    """BSpline / LogXBSpline bases: the spline itself is the density (interpolation.py:293-317)."""
    pe, inj, _ = cat
    a = M.BSplineChiEffective(8, pe["cos_tilt_1"], inj["cos_tilt_1"], normalize=True)
    b = M.BSplineSymmetricChiEffective(6, pe["cos_tilt_2"], inj["cos_tilt_2"])
    z = M.BSplineRedshift(7, pe["redshift"], inj["redshift"], pe["prior"], inj["prior"], normalize=False)
    ca, cb, cz = np.ones(8), np.ones(6), np.zeros(7)
    low = lowering.lower(a(ca, pe_samples=True) * b(cb, pe_samples=True) * z(cz, pe_samples=True),
                         a(ca, pe_samples=False) * b(cb, pe_samples=False) * z(cz, pe_samples=False))
    kinds = [t.kind for t in low.spec.terms]
    assert kinds.count(S.TERM_SPLINE_LINEAR) == 2 and kinds.count(S.TERM_SPLINE) == 1 and kinds.count(S.TERM_STATIC) == 3
    assert low.spec.n_params == 21 and len(low.spec.groups) == 2  # chi_eff normalised, symmetric one not, redshift Z
    t = [t for t in low.spec.terms if t.kind == S.TERM_SPLINE][0]
    assert t.outside == S.OUTSIDE_ZERO and t.logx and t.xrange == (1e-4, 2.3)
    assert np.array_equal(low.pe_cols[[t for t in low.spec.terms if t.kind == S.TERM_SPLINE_LINEAR][1].cols[0]], np.abs(pe["cos_tilt_2"]))


def test_log_domain_weights_lower_to_the_same_model(cat):
    """``log_prob`` terms combined with + / - log(prior) (analysis.py:401-402, log=True) give the same
    device model as the product form."""
    pe, inj, _ = cat
    z = M.PowerlawRedshiftModel(pe["redshift"], inj["redshift"])
    lamb = np.float64(2.3)
    prod = lowering.lower(z(pe["redshift"], lamb) / pe["prior"], z(inj["redshift"], lamb) / inj["prior"])
    lw_pe = z.log_prob(pe["redshift"], lamb) - np.log(pe["prior"])
    lw_inj = z.log_prob(inj["redshift"], lamb) - np.log(inj["prior"])
    assert lw_pe.log_domain and lw_inj.log_domain
    logd = lowering.lower(lw_pe, lw_inj)
    assert [(t.kind, t.feature) for t in logd.spec.terms] == [(t.kind, t.feature) for t in prod.spec.terms]
    prior_col = [t for t in logd.spec.terms if t.feature == S.FEAT_NEG_LOG][0].cols[0]
    assert np.allclose(logd.pe_cols[prior_col], pe["prior"], rtol=1e-15)
    with pytest.raises(TypeError):
        _ = lw_pe * z(pe["redshift"], lamb)
    with pytest.raises(TypeError):
        _ = z(pe["redshift"], lamb) + z(pe["redshift"], lamb)
    s = z.log_prob(pe["redshift"], lamb) + z.log_prob(pe["redshift"], lamb)
    assert s.log_domain and len(s.terms) == 2
