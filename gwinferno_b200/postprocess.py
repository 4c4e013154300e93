"""Posterior-predictive densities on grids (mirror of ``gwinferno/postprocess/calculations.py:20-276``), evaluated by
the SAME device path as the likelihood.

The reference builds a population model on an 800 x 800 mesh and, per posterior draw, evaluates the dense density and
integrates it with the trapezoid rule along one axis.  A trapezoid integral over a grid row is a weighted sum of
per-sample densities -- exactly the per-event Monte-Carlo sum of the hot path with

    event i  = one grid point of the axis that is kept,        samples j = the grid of the axis integrated out,
    1/prior  = the trapezoid weight of sample j (0 where the reference masks the density),

so ``sum_j omega_j p(x_i, y_j | Lambda) = S exp(logBF_i)`` comes straight out of ``gwi_eval``: the plan (piece words
of the mesh) is built once per grid, every draw is one fused evaluation.  1-D densities on a grid are the S = 1 case.
The final ``rate * frac * p / trapezoid(p)`` normalisation is O(grid) host arithmetic as in the reference.

There is no CPU fallback (``capi`` raises without libgwi.so / a GPU); the CPU test suite runs these functions on the
host warp emulator and compares them with the reference's own functions executed under the NumPy ``jax`` shim
(``tests/golden/ppd_*.npz``).
"""

import numpy as np

from . import lowering
from . import models as M
from .likelihood import PopulationLikelihood

__all__ = [
    "calculate_bspline_mass_ppds",
    "calculate_powerlaw_peak_mass_ppds",
    "calculate_bspline_spin_ppds",
    "calculate_beta_spin_mag",
    "calculate_mixture_iso_aligned_spin_tilt",
    "calculate_powerlaw_rate_of_z_ppds",
    "calculate_powerlaw_spline_rate_of_z_ppds",
    "GridDensity",
]

N_GRID = 800  # calculations.py:22-23,135,163,...


def _trapezoid_weights(x):
    x = np.asarray(x, dtype=np.float64)
    w = np.empty_like(x)
    w[0] = 0.5 * (x[1] - x[0])
    w[-1] = 0.5 * (x[-1] - x[-2])
    w[1:-1] = 0.5 * (x[2:] - x[:-2])
    return w


def _trapezoid(y, x):
    return float(np.sum(0.5 * (y[1:] + y[:-1]) * np.diff(x)))


class GridDensity:
    """Row sums ``sum_j w_ij`` of a lazy weight product ``w`` on the device: one plan per grid, one evaluation per
    draw.  ``pe_w`` / ``inj_w``: the weights of ANY draw (they fix the structure); the injection side is only there
    because a model description always has both sample sets (a handful of grid points)."""

    def __init__(self, pe_w, inj_w, device=0):
        self.low = lowering.lower(pe_w, inj_w)
        n_inj = next(iter(self.low.inj_cols.values())).shape[0]
        self.engine = PopulationLikelihood(self.low, float(max(1, n_inj)), device=device)
        self.n_rows, self.n_samples = next(iter(self.low.pe_cols.values())).shape

    def rows(self, pe_w):
        lam = lowering.flatten_params(pe_w, self.low.spec.n_params)
        r = self.engine.evaluate(lam, jacobians=False)
        return self.n_samples * np.exp(np.asarray(r["logBF"], dtype=np.float64))  # exp(-inf) = 0: rows with no support

    def close(self):
        self.engine.model.close()


def _ones(n, default):
    return np.ones(n) if default is None else np.asarray(default, dtype=np.float64)


def _marginals(make_weights, n_draws, ms, qs, keep, rate, pop_frac):
    """Shared body of the two mass PPD functions.  ``make_weights(i, M, Q, m1d, q1d) -> (pe_w, inj_w)`` for draw i on
    the given (mesh, 1-D) arrays; ``keep``: the reference's mask on the [q, m] mesh."""
    Mm, Qm = np.meshgrid(ms, qs)  # [q, m]
    wm, wq = _trapezoid_weights(ms), _trapezoid_weights(qs)
    with np.errstate(divide="ignore"):
        # rows = q grid points, samples along m (-> p_q); and the transposed mesh (-> p_m)
        setups = [
            (np.ascontiguousarray(Mm), np.ascontiguousarray(Qm), np.where(keep, 1.0 / wm[None, :], np.inf)),
            (np.ascontiguousarray(Mm.T), np.ascontiguousarray(Qm.T), np.where(keep.T, 1.0 / wq[None, :], np.inf)),
        ]
    one = np.ones(len(ms))
    mpdfs, qpdfs = np.zeros((n_draws, len(ms))), np.zeros((n_draws, len(qs)))
    grids = []
    try:
        for Mx, Qx, prior in setups:
            def weights(i, Mx=Mx, Qx=Qx, prior=prior):
                pe_w, inj_w = make_weights(i, Mx, Qx)
                return pe_w / prior, inj_w / one

            grids.append((GridDensity(*weights(0)), weights))
        for i in range(n_draws):
            p_q = grids[0][0].rows(grids[0][1](i)[0])
            p_m = grids[1][0].rows(grids[1][1](i)[0])
            mpdfs[i] = rate[i] * p_m * pop_frac[i] / _trapezoid(p_m, ms)
            qpdfs[i] = rate[i] * p_q * pop_frac[i] / _trapezoid(p_q, qs)
    finally:
        for g, _ in grids:
            g.close()
    return mpdfs, qpdfs


def calculate_bspline_mass_ppds(m_cs, q_cs, nspline_dict, mmin, mmax, rate=None, pop_frac=None):
    """calculations.py:20-60: ``(mpdfs [n, 800], ms, qpdfs [n, 800], qs)`` for the B-spline primary-mass x mass-ratio
    model, ``p(m, q)`` masked to ``q > mmin / m``."""
    m_cs, q_cs = np.atleast_2d(np.asarray(m_cs, dtype=np.float64)), np.atleast_2d(np.asarray(q_cs, dtype=np.float64))
    n = m_cs.shape[0]
    rate, pop_frac = _ones(n, rate), _ones(n, pop_frac)
    ms, qs = np.linspace(mmin, mmax, N_GRID), np.linspace(mmin / mmax, 1, N_GRID)
    Mm, Qm = np.meshgrid(ms, qs)
    keep = Qm > mmin / Mm
    models = {}

    def make_weights(i, Mx, Qx):
        model = models.get(id(Mx))
        if model is None:
            model = models[id(Mx)] = M.BSplinePrimaryBSplineRatio(nspline_dict["m1"], nspline_dict["q"], Mx, ms, Qx, qs, m1min=mmin, m2min=mmin, mmax=mmax)
        return model(m_cs[i], q_cs[i], pe_samples=True), model(m_cs[i], q_cs[i], pe_samples=False)

    mpdfs, qpdfs = _marginals(make_weights, n, ms, qs, keep, rate, pop_frac)
    return mpdfs, ms, qpdfs, qs


def calculate_powerlaw_peak_mass_ppds(alpha, beta, mu_peak, sig_peak, lamb, mmin, mmax, rate=None, pop_frac=None):
    """calculations.py:63-93: the same marginals for the power-law + peak model."""
    alpha, beta, mu_peak, sig_peak, lamb = (np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in (alpha, beta, mu_peak, sig_peak, lamb))
    n = alpha.shape[0]
    rate, pop_frac = _ones(n, rate), _ones(n, pop_frac)
    ms, qs = np.linspace(mmin, mmax, N_GRID), np.linspace(mmin / mmax, 1, N_GRID)
    Mm, Qm = np.meshgrid(ms, qs)
    keep = Qm > mmin / Mm

    def make_weights(i, Mx, Qx):
        args = (alpha[i], beta[i], mmin, mmax, mu_peak[i], sig_peak[i], lamb[i])
        return M.plpeak_primary_ratio_pdf(Mx, Qx, *args), M.plpeak_primary_ratio_pdf(ms, qs, *args)

    mpdfs, qpdfs = _marginals(make_weights, n, ms, qs, keep, rate, pop_frac)
    return mpdfs, ms, qpdfs, qs


def _pdf_on_grid(make_weights, n_draws, grid, rate, pop_frac):
    """1-D density on a grid: every grid point is a row with one sample."""
    col, one = np.ascontiguousarray(grid[:, None]), np.ones(len(grid))
    out = np.zeros((n_draws, len(grid)))
    g = None
    try:
        for i in range(n_draws):
            pe_w, inj_w = make_weights(i, col)
            pe_w, inj_w = pe_w / np.ones_like(col), inj_w / one
            if g is None:
                g = GridDensity(pe_w, inj_w)
            p = g.rows(pe_w)
            out[i] = rate[i] * pop_frac[i] * p / _trapezoid(p, grid)
    finally:
        if g is not None:
            g.close()
    return out


def calculate_bspline_spin_ppds(a1_cs, tilt1_cs, nspline_dict, a2_cs=None, tilt2_cs=None, rate=None, pop_frac=None):
    """calculations.py:181-242: spin-magnitude and tilt densities of the B-spline spin models on 800-point grids; IID
    form ``(apdfs, aa, ctpdfs, cc)``, independent form ``(apdfs_1, apdfs_2, aa, ctpdfs_1, ctpdfs_2, cc)``."""
    a1_cs, tilt1_cs = np.atleast_2d(np.asarray(a1_cs, dtype=np.float64)), np.atleast_2d(np.asarray(tilt1_cs, dtype=np.float64))
    n = a1_cs.shape[0]
    rate, pop_frac = _ones(n, rate), _ones(n, pop_frac)
    aa, cc = np.linspace(0, 1, N_GRID), np.linspace(-1, 1, N_GRID)

    def spline_pdf(cls, n_splines, grid, cs):
        holder = {}

        def make(i, col):
            mdl = holder.get("m")
            if mdl is None:
                mdl = holder["m"] = cls(n_splines, col, grid, basis=M.LogYBSpline, normalize=True)
            return mdl(cs[i], pe_samples=True), mdl(cs[i], pe_samples=False)

        return _pdf_on_grid(make, n, grid, rate, pop_frac)

    if a2_cs is None:
        return spline_pdf(M.BSplineSpinMagnitude, nspline_dict["a"], aa, a1_cs), aa, spline_pdf(M.BSplineSpinTilt, nspline_dict["tilt"], cc, tilt1_cs), cc
    a2_cs, tilt2_cs = np.atleast_2d(np.asarray(a2_cs, dtype=np.float64)), np.atleast_2d(np.asarray(tilt2_cs, dtype=np.float64))
    return (
        spline_pdf(M.BSplineSpinMagnitude, nspline_dict["a1"], aa, a1_cs),
        spline_pdf(M.BSplineSpinMagnitude, nspline_dict["a2"], aa, a2_cs),
        aa,
        spline_pdf(M.BSplineSpinTilt, nspline_dict["tilt1"], cc, tilt1_cs),
        spline_pdf(M.BSplineSpinTilt, nspline_dict["tilt2"], cc, tilt2_cs),
        cc,
    )


def calculate_beta_spin_mag(alpha_a, beta_a, amax=1, rate=None, pop_frac=None):
    """calculations.py:133-154."""
    alpha_a, beta_a = np.atleast_1d(np.asarray(alpha_a, dtype=np.float64)), np.atleast_1d(np.asarray(beta_a, dtype=np.float64))
    n = alpha_a.shape[0]
    aa = np.linspace(0, amax, N_GRID)
    make = lambda i, col: (M.beta_spin_magnitude(col, alpha_a[i], beta_a[i], amax), M.beta_spin_magnitude(aa, alpha_a[i], beta_a[i], amax))
    return _pdf_on_grid(make, n, aa, _ones(n, rate), _ones(n, pop_frac)), aa


def calculate_mixture_iso_aligned_spin_tilt(sig_ct, lambda_ct, rate=None, pop_frac=None):
    """calculations.py:157-178."""
    sig_ct, lambda_ct = np.atleast_1d(np.asarray(sig_ct, dtype=np.float64)), np.atleast_1d(np.asarray(lambda_ct, dtype=np.float64))
    n = sig_ct.shape[0]
    ct = np.linspace(-1, 1, N_GRID)
    make = lambda i, col: (M.mixture_isoalign_spin_tilt(col, lambda_ct[i], sig_ct[i]), M.mixture_isoalign_spin_tilt(ct, lambda_ct[i], sig_ct[i]))
    return _pdf_on_grid(make, n, ct, _ones(n, rate), _ones(n, pop_frac)), ct


def calculate_powerlaw_rate_of_z_ppds(lamb, rate, z_model, pop_frac=None):
    """calculations.py:244-258: ``R(z) = rate * frac * (1+z)^lamb`` on the model's redshift grid (O(grid) host arithmetic)."""
    lamb, rate = np.atleast_1d(np.asarray(lamb, dtype=np.float64)), np.atleast_1d(np.asarray(rate, dtype=np.float64))
    pop_frac = _ones(lamb.shape[0], pop_frac)
    zs = np.asarray(z_model.zs, dtype=np.float64)
    return rate[:, None] * pop_frac[:, None] * np.power(1.0 + zs[None, :], lamb[:, None]), zs


def calculate_powerlaw_spline_rate_of_z_ppds(lamb, z_cs, rate, z_model, pop_frac=None):
    """calculations.py:261-276: the power law times ``exp(B(log z) . [0, cs])`` on the model's grid."""
    lamb, rate = np.atleast_1d(np.asarray(lamb, dtype=np.float64)), np.atleast_1d(np.asarray(rate, dtype=np.float64))
    z_cs = np.atleast_2d(np.asarray(z_cs, dtype=np.float64))
    pop_frac = _ones(z_cs.shape[0], pop_frac)
    zs = np.asarray(z_model.zs, dtype=np.float64)
    D = lowering.host_spline_design(np.log(zs), z_model.xi_range, z_model.n_splines)  # [G, n_splines]
    cs = np.concatenate([np.zeros((z_cs.shape[0], 1)), z_cs], axis=1)
    return rate[:, None] * pop_frac[:, None] * np.power(1.0 + zs[None, :], lamb[:, None]) * np.exp(cs @ D.T), zs
