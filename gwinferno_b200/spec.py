"""Declarative description of a population model: the boundary between the reference-facing
Python classes (gwinferno_b200/models.py) and the C-ABI (include/gwi.h, ``gwi_term`` /
``gwi_norm_group`` / ``gwi_cut``).

A model is a list of additive log-density TERMS over named catalog columns, a list of static
CUTS (samples whose population density is exactly zero in the reference, e.g. outside
``[xmin, xmax]`` -- gwinferno/models/bsplines/single.py:54-55,90-92), and a list of NORM GROUPS
(sample-independent normalisers evaluated on a 1000/1500-point trapezoid grid --
gwinferno/interpolation.py:290,378,433; gwinferno/models/spline_perturbation.py:323-336).
The per-sample log-weight is

    x_j = sum_terms t(theta_j; Lambda)  -  sum_groups log Z_g(Lambda)         (x_j = -inf if cut)

Numeric values of the enums below are part of the C-ABI (include/gwi.h).
"""

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# ---- gwi_term_kind ------------------------------------------------------------------------
TERM_SPLINE = 1  # sum_k B_k(xi) c_k ; xi = col or log(col); uniform cubic knots
TERM_LINEAR = 2  # (Lambda[slot] + offset) * F(col)
TERM_STATIC = 3  # F(col), no parameter
TERM_POWERLAW = 4  # log powerlaw_pdf(col; alpha, lo, hi)            distributions.py:100-119
TERM_POWERLAW_RATIO = 5  # log powerlaw_pdf(q; beta, mmin/m1, 1)     parametric.py:28,40
TERM_PLPEAK = 6  # log[(1-lam) PL(m1) + lam TN(m1)]                  parametric.py:49-53
TERM_BETA = 7  # log betadist(col; alpha, beta, scale)               distributions.py:146-162
TERM_ISOALIGN = 8  # log[(1-xi)/2 + xi TN(ct; 1, sigma, -1, 1)]      parametric.py:84-86
TERM_TRUNCNORM = 9  # log truncnorm_pdf(col; mu, sig, lo, hi)        distributions.py:122-143
TERM_SPLINE_LINEAR = 10  # log sum_k B_k(xi) c_k : the spline is the density  interpolation.py:280-317
TERM_ISOALIGN_PAIR = 12  # log[(1-xi)/4 + xi TN(ct1) TN(ct2)]   (default_spin_tilt, parametric.py:97-102)
TERM_SMOOTH = 11  # log smooth(delta; col0 [* col1], xmin)           distributions.py:16-21

# ---- gwi_feature --------------------------------------------------------------------------
FEAT_LOG1P = 1  # log(1 + col0)
FEAT_LOG = 2  # log(col0)
FEAT_LOG_RATIO = 3  # log(col0 / col1)
FEAT_LOG_DVDZ = 4  # log dVc/dz(col0)   (Planck15-LVK table, cosmology.py:95-120)
FEAT_NEG_LOG = 5  # -log(col0)          (division by the sampling prior)
FEAT_NEG_LOG1P = 6  # -log(1 + col0)    (the 1/(1+z) of BSplineRedshift, single.py:488-491)
FEAT_CONST = 7  # cst[0]                (a constant factor, e.g. the 0.5 of single.py:284)

# ---- gwi_outside (SPLINE only) ------------------------------------------------------------
OUTSIDE_DROP = 0  # LogY-type bases are -inf outside xrange => pdf 0 (interpolation.py:407,449)
OUTSIDE_ZERO = 1  # B/LogX bases are 0 outside xrange => spline term contributes 0 (:175)

# ---- gwi_cut_kind -------------------------------------------------------------------------
CUT_RANGE = 1  # keep lo <= col0 <= hi
CUT_RATIO_RANGE = 2  # keep lo <= col0/col1 <= hi


@dataclass
class Term:
    kind: int
    cols: List[str]
    slots: List[int] = field(default_factory=list)  # offsets into Lambda
    cst: List[float] = field(default_factory=list)
    feature: int = 0
    n_splines: int = 0
    logx: bool = False
    outside: int = OUTSIDE_DROP
    xrange: tuple = (0.0, 1.0)  # SPLINE: domain in x (not log x)
    xi_range: Optional[tuple] = None  # SPLINE: domain in the spline coordinate (log x if logx)
    norm_group: int = -1
    # grids (one entry per grid point of ``norm_group``)
    grid_xi: Optional[np.ndarray] = None  # SPLINE: spline coordinate at the grid points
    grid_feat: Optional[np.ndarray] = None  # LINEAR: feature value at the grid points
    # SPLINE: explicit knot vector in spline-coordinate units (len n_splines + order) and order = degree + 1, the
    # reference's knots= / interior_knots= / k= (interpolation.py:72-106); None = default uniform cubic knots
    knots: Optional[np.ndarray] = None
    order: int = 4
    name: str = ""


@dataclass
class NormGroup:
    log_w: np.ndarray  # log(trapezoid weight * static integrand) per grid point (-inf allowed)
    name: str = ""


@dataclass
class Cut:
    kind: int
    cols: List[str]
    lo: float
    hi: float


@dataclass
class ModelSpec:
    terms: List[Term]
    groups: List[NormGroup]
    cuts: List[Cut]
    n_params: int
    param_names: List[str] = field(default_factory=list)

    def columns(self):
        out = []
        for t in self.terms:
            for c in t.cols:
                if c not in out:
                    out.append(c)
        for c in self.cuts:
            for cc in c.cols:
                if cc not in out:
                    out.append(cc)
        return out


def trapezoid_weights(grid):
    """w_g such that sum_g w_g y_g == trapezoid(y, grid)."""
    grid = np.asarray(grid, dtype=np.float64)
    w = np.zeros_like(grid)
    d = np.diff(grid)
    w[:-1] += 0.5 * d
    w[1:] += 0.5 * d
    return w


def uniform_knots(n_splines, lo, hi):
    """(x0, dx, n_int) of the reference's default knot vector (interpolation.py:98-106):
    ``n_int = n_splines - 2`` interior knots linspace(lo, hi), extended by 3 dx on either side."""
    n_int = n_splines - 2
    if n_int < 2:
        raise ValueError("need at least 4 basis functions for a cubic B-spline")
    dx = (hi - lo) / (n_int - 1)
    return lo, dx, n_int
