"""Mirrors of the reference's pipeline helpers that sit right around the hot path
(gwinferno/pipeline/utils.py:104-216, gwinferno/models/bsplines/smoothing.py:8-28): model set-up
from the ``(pedict, injdict)`` dictionaries with the reference's basis choices, and the P-spline
difference prior with its gradient (O(P) host arithmetic; the reference gets the gradient from JAX).

The NumPyro ``sample`` / ``factor`` statements of the reference's prior helpers
(pipeline/utils.py:163-216) have no stand-in here (no NumPyro); :func:`bspline_prior_blocks` gives the
same Normal scales, smoothing strengths and difference orders as block descriptions for the host
sampler (gwinferno_b200/nuts.py: ``BSplinePosterior``).
"""

import numpy as np

from . import models as M


def setup_bspline_mass_models(pedict, injdict, m_nsplines, q_nsplines, mmin, mmax):
    """pipeline/utils.py:104-118."""
    return M.BSplinePrimaryBSplineRatio(
        m_nsplines, q_nsplines, pedict["mass_1"], injdict["mass_1"], pedict["mass_ratio"], injdict["mass_ratio"],
        m1min=mmin, m2min=mmin, mmax=mmax, kwargs_m={"basis": M.LogXLogYBSpline}, kwargs_q={"basis": M.LogYBSpline},
    )


def setup_bspline_spin_models(pedict, injdict, a1_nsplines, ct1_nsplines, IID=False, a2_nsplines=None, ct2_nsplines=None):
    """pipeline/utils.py:121-146: ``(mag_model, tilt_model)``."""
    if IID:
        tilt_model = M.BSplineIIDSpinTilts(ct1_nsplines, pedict["cos_tilt_1"], pedict["cos_tilt_2"], injdict["cos_tilt_1"], injdict["cos_tilt_2"], normalize=True)
        mag_model = M.BSplineIIDSpinMagnitudes(a1_nsplines, pedict["a_1"], pedict["a_2"], injdict["a_1"], injdict["a_2"], normalize=True)
    else:
        tilt_model = M.BSplineIndependentSpinTilts(
            ct1_nsplines, ct2_nsplines, pedict["cos_tilt_1"], pedict["cos_tilt_2"], injdict["cos_tilt_1"], injdict["cos_tilt_2"], normalize=True
        )
        mag_model = M.BSplineIndependentSpinMagnitudes(a1_nsplines, a2_nsplines, pedict["a_1"], pedict["a_2"], injdict["a_1"], injdict["a_2"], normalize=True)
    return mag_model, tilt_model


def setup_powerlaw_spline_redshift_model(pedict, injdict, z_nsplines):
    """pipeline/utils.py:149-155."""
    return M.PowerlawSplineRedshiftModel(z_nsplines, pedict["redshift"], injdict["redshift"])


def apply_difference_prior(coefs, inv_var, degree=1):
    """P-spline difference penalty ``-inv_var/2 |D^degree c|^2`` (smoothing.py:8-28)."""
    delta_c = np.diff(np.asarray(coefs, dtype=np.float64), n=degree)
    return -0.5 * inv_var * float(delta_c @ delta_c)


def difference_prior_grad(coefs, inv_var, degree=1):
    """Gradient of :func:`apply_difference_prior` with respect to ``coefs``: ``-inv_var D^T D c``."""
    c = np.asarray(coefs, dtype=np.float64)
    g = np.diff(c, n=degree)
    for _ in range(degree):  # apply the transpose of one first-difference at a time
        g = np.concatenate([[-g[0]], -np.diff(g), [g[-1]]])
    return -inv_var * g


def bspline_prior_blocks(slots_for, params, m_tau=1.0, q_tau=1.0, a_tau=25.0, ct_tau=25.0, z_tau=1.0, lamb_sig=3.0,
                         m_cs_sig=15.0, q_cs_sig=5.0, a_cs_sig=5.0, ct_cs_sig=5.0, z_cs_sig=1.0, m_deg=1, q_deg=1, a_deg=2, ct_deg=2, z_deg=2):
    """Prior blocks ``(lambda_slice, normal_sigma, smoothing_tau, difference_order, first_fixed_to_zero)``
    of the reference's B-spline priors (pipeline/utils.py:163-216; taus as in
    examples/simple_bspline_example.py) for whichever of the usual parameter names ``params`` holds."""
    table = {
        "mass_cs": (m_cs_sig, m_tau, m_deg, False), "q_cs": (q_cs_sig, q_tau, q_deg, False),
        "a_cs": (a_cs_sig, a_tau, a_deg, False), "a1_cs": (a_cs_sig, a_tau, a_deg, False), "a2_cs": (a_cs_sig, a_tau, a_deg, False),
        "tilt_cs": (ct_cs_sig, ct_tau, ct_deg, False), "tilt1_cs": (ct_cs_sig, ct_tau, ct_deg, False), "tilt2_cs": (ct_cs_sig, ct_tau, ct_deg, False),
        "z_cs": (z_cs_sig, z_tau, z_deg, True),  # z_cs = concatenate([0], sampled)  (pipeline/utils.py:213-214)
        "lamb": (lamb_sig, None, 0, False),
    }
    return [(slots_for(params[k]),) + table[k] for k in params if k in table]
