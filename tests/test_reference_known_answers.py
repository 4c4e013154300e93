"""CPU: the known-answer tests the reference holds for this path, restated against BOTH oracles
(NumPy oracle/popmodel.py and the plain-C oracle) through the model description the product uses:

  tests/distributions_test.py:30-88          power law == SciPy truncpareto, truncated normal, beta
  tests/interpolation_test.py:57-85          normalised spline densities integrate to one (all four bases)
  tests/models/bsplines/separable_test.py    exactly zero density outside the mass support
  tests/models/parametric/parametric_test.py:44-71   redshift power law normalised, zero above zmax

The reference asserts these at rtol 1e-5 / 3 decimal places; here the closed forms are required to
1e-12 (the trapezoid identities hold to rounding because the same grid is used on both sides)."""

import numpy as np
import pytest
from scipy.stats import beta as sbeta
from scipy.stats import truncnorm, truncpareto

from gwinferno_b200 import lowering
from gwinferno_b200 import models as M
from oracle import c_oracle, popmodel

ORACLES = ["numpy", "c"]


def _log_density(oracle, weight_fn, x, params):
    """log p(x) of a 1-D lazy model on the points ``x`` (one 'event' holding the points, and the same
    points as the 'injections'), including its grid normaliser."""
    pe, inj = np.ascontiguousarray(x[None, :]), np.ascontiguousarray(x)
    low = lowering.lower(weight_fn(pe, True, params), weight_fn(inj, False, params))
    lam = lowering.flatten_params(weight_fn(pe, True, params), low.spec.n_params)
    if oracle == "numpy":
        xw, valid, _ = popmodel.log_weights(low.spec, low.inj_cols, lam)
        logZ, _ = popmodel.log_normalisers(low.spec, lam)
        return np.where(valid, xw - np.sum(logZ), -np.inf)
    # C oracle: one injection per call would be slow; use the per-event path with ONE sample per event
    pe1 = {k: np.ascontiguousarray(v.reshape(-1, 1)) for k, v in low.inj_cols.items()}
    ev = c_oracle.evaluate(low.spec, pe1, low.inj_cols, float(x.size), lam, want_jac=False, n_threads=2)
    return ev["logBF"]  # log(w / 1) of the single sample


@pytest.mark.parametrize("oracle", ORACLES)
def test_powerlaw_matches_scipy_truncpareto(oracle):
    x = np.linspace(2, 55, 1000)
    alpha, xmin, xmax = -3.2, 3.0, 50.0
    a = np.float64(alpha)
    lp = _log_density(oracle, lambda d, pe, p: M._powerlaw_term(d, p, xmin, xmax, "t"), x, a)
    expect = truncpareto.pdf(x, -alpha - 1, xmax / xmin, loc=0.0, scale=xmin)
    inside = (x >= xmin) & (x <= xmax)
    assert np.allclose(np.exp(lp[inside]), expect[inside], rtol=1e-12)
    assert np.all(np.exp(lp[~inside]) == 0.0) and np.all(expect[~inside] == 0.0)


@pytest.mark.parametrize("oracle", ORACLES)
def test_truncnorm_matches_scipy(oracle):
    x = np.linspace(-1, 1.2, 50)
    mu, sigma, lo, hi = np.float64(0.3), np.float64(1.4), -0.8, 1.0

    def w(d, pe, p):
        from gwinferno_b200 import spec as S

        col = M._col_of(d)
        return M.LazyWeight([M._LazyTerm(("tn", id(d)), [col], [p[0], p[1]],
                                         lambda slots, g, cols: ([S.Term(S.TERM_TRUNCNORM, [cols[0]], slots=list(slots[:2]), cst=[lo, hi])], [], []))], pe)

    lp = _log_density(oracle, w, x, (mu, sigma))
    expect = truncnorm.pdf(x, (lo - mu) / sigma, (hi - mu) / sigma, loc=mu, scale=sigma)
    assert np.allclose(np.exp(lp), expect, rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("oracle", ORACLES)
def test_betadist_matches_scipy(oracle):
    x = np.linspace(0, 1, 50)
    a, b = np.float64(2.0), np.float64(3.0)
    lp = _log_density(oracle, lambda d, pe, p: M.beta_spin_magnitude(d, p[0], p[1]), x, (a, b))
    interior = (x > 0) & (x < 1)  # the end points are 0 in SciPy and dropped by the model
    assert np.allclose(np.exp(lp[interior]), sbeta.pdf(x[interior], 2, 3), rtol=1e-12)
    assert np.all(np.exp(lp[~interior]) == 0.0)


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("basis,xrange,positive", [(M.BSpline, (0.0, 1.0), True), (M.LogYBSpline, (0.0, 1.0), False),
                                                   (M.LogXBSpline, (0.001, 1.0), True), (M.LogXLogYBSpline, (0.001, 1.0), False)])
def test_normalised_spline_densities_integrate_to_one(oracle, basis, xrange, positive):
    """interpolation_test.py:57-85 (there: 3 decimal places on a 1000-point grid)."""
    rng = np.random.default_rng(7)
    N = 10
    cs = rng.uniform(size=N) if positive else rng.normal(size=N)
    grid = np.linspace(xrange[0], xrange[1], basis.n_grid)
    m = M.Base1DBSplineModel(N, grid[None, :], grid, xrange=xrange, basis=basis, normalize=True)
    low = lowering.lower(m(cs, pe_samples=True), m(cs, pe_samples=False))
    lam = lowering.flatten_params(m(cs, pe_samples=True), low.spec.n_params)
    if oracle == "numpy":
        xw, valid, _ = popmodel.log_weights(low.spec, low.inj_cols, lam)
        logZ, _ = popmodel.log_normalisers(low.spec, lam)
        dens = np.where(valid, np.exp(xw - np.sum(logZ)), 0.0)
    else:
        pe1 = {k: np.ascontiguousarray(v.reshape(-1, 1)) for k, v in low.inj_cols.items()}
        dens = np.exp(c_oracle.evaluate(low.spec, pe1, low.inj_cols, float(grid.size), lam, want_jac=False)["logBF"])
    assert abs(np.trapezoid(dens, grid) - 1.0) < 1e-12


@pytest.mark.parametrize("oracle", ORACLES)
def test_zero_density_outside_the_mass_support(oracle):
    """separable_test.py:93-97,118-122,141-145: samples outside [mmin, mmax] get exactly zero."""
    rng = np.random.default_rng(11)
    mmin, mmax = 5.0, 60.0
    m1 = rng.uniform(2.0, 80.0, 400)
    q = rng.uniform(0.05, 1.0, 400)
    coefs = rng.normal(size=12)
    mm = M.BSplinePrimaryBSplineRatio(12, 12, m1[None, :], m1, q[None, :], q, m1min=mmin, m2min=mmin, mmax=mmax)
    low = lowering.lower(mm(coefs, coefs, pe_samples=True), mm(coefs, coefs, pe_samples=False))
    lam = lowering.flatten_params(mm(coefs, coefs, pe_samples=True), low.spec.n_params)
    if oracle == "numpy":
        xw, valid, _ = popmodel.log_weights(low.spec, low.inj_cols, lam)
        dens = np.where(valid, np.exp(xw), 0.0)
    else:
        pe1 = {k: np.ascontiguousarray(v.reshape(-1, 1)) for k, v in low.inj_cols.items()}
        dens = np.exp(c_oracle.evaluate(low.spec, pe1, low.inj_cols, 400.0, lam, want_jac=False)["logBF"])
        dens = np.where(np.isfinite(dens), dens, 0.0)
    outside = (m1 < mmin) | (m1 > mmax)
    assert outside.any() and np.all(dens[outside] == 0.0) and np.all(dens[~outside & (q >= mmin / mmax)] > 0.0)


@pytest.mark.parametrize("oracle", ORACLES)
def test_powerlaw_redshift_normalised_and_truncated(oracle):
    """parametric_test.py:44-71: trapezoid(prob / normalization) == 1 on the model's own grid, zero above zmax."""
    rng = np.random.default_rng(13)
    z_pe = rng.uniform(0.01, 1.5, (3, 50))
    z_inj = rng.uniform(0.02, 1.2, 500)
    lamb = np.float64(2.7)
    zm = M.PowerlawRedshiftModel(z_pe, z_inj)
    zs = np.concatenate([zm.zs, [zm.zmax * 1.01, zm.zmax * 1.5]])  # the grid plus two points above zmax
    probe = M.PowerlawRedshiftModel(zs[None, :], zs, z_range=(zm.zmin, zm.zmax))
    low = lowering.lower(probe(probe.column.pe, lamb), probe(probe.column.inj, lamb))
    lam = lowering.flatten_params(probe(probe.column.pe, lamb), low.spec.n_params)
    if oracle == "numpy":
        xw, valid, _ = popmodel.log_weights(low.spec, low.inj_cols, lam)
        logZ, _ = popmodel.log_normalisers(low.spec, lam)
        dens = np.where(valid, np.exp(xw - np.sum(logZ)), 0.0)
    else:
        pe1 = {k: np.ascontiguousarray(v.reshape(-1, 1)) for k, v in low.inj_cols.items()}
        dens = np.exp(c_oracle.evaluate(low.spec, pe1, low.inj_cols, float(zs.size), lam, want_jac=False)["logBF"])
        dens = np.where(np.isfinite(dens), dens, 0.0)
    assert abs(np.trapezoid(dens[:-2], zm.zs) - 1.0) < 1e-12
    assert np.all(dens[-2:] == 0.0)
    assert abs(zm.normalization(lamb) - np.exp(popmodel.log_normalisers(low.spec, lam)[0][0])) < 1e-9 * zm.normalization(lamb)
