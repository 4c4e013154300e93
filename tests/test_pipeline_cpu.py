"""CPU: the pipeline-helper mirrors against the reference's own code (run under the NumPy shim
when /root/reference is present) and against the direct construction."""

import os

import numpy as np
import pytest

from gwinferno_b200 import lowering, pipeline, synthetic
from gwinferno_b200 import models as M

HAVE_REF = os.path.isdir("/root/reference/gwinferno")


@pytest.fixture(scope="module")
def cat():
    return synthetic.make_catalog(4, 60, 800, cfg=150)


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_difference_prior_and_gradient(degree):
    rng = np.random.default_rng(degree)
    c = rng.standard_normal(12)
    tau = 7.5
    D = np.diff(np.eye(c.size), n=degree, axis=0)
    assert np.isclose(pipeline.apply_difference_prior(c, tau, degree), -0.5 * tau * (D @ c) @ (D @ c), rtol=1e-14)
    assert np.allclose(pipeline.difference_prior_grad(c, tau, degree), -tau * D.T @ (D @ c), rtol=1e-13, atol=1e-13)


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree (build container only)")
def test_difference_prior_matches_reference_code():
    from oracle import jax_shim

    ref = jax_shim.load_reference()["smoothing"]
    rng = np.random.default_rng(3)
    for degree in (1, 2):
        c = rng.standard_normal(16)
        assert np.isclose(pipeline.apply_difference_prior(c, 25.0, degree), float(ref.apply_difference_prior(c, 25.0, degree=degree)), rtol=1e-14)


def test_setup_helpers_build_the_reference_configuration(cat):
    pe, inj, _ = cat
    mass = pipeline.setup_bspline_mass_models(pe, inj, 12, 8, 3.0, 100.0)
    mag, tilt = pipeline.setup_bspline_spin_models(pe, inj, 6, 7, IID=False, a2_nsplines=6, ct2_nsplines=7)
    zmod = pipeline.setup_powerlaw_spline_redshift_model(pe, inj, 5)
    assert mass.primary_model.basis is M.LogXLogYBSpline and mass.ratio_model.basis is M.LogYBSpline  # pipeline/utils.py:115-116
    assert (mass.ratio_model.xmin, mass.ratio_model.xmax) == (3.0 / 100.0, 1)
    assert mag.primary_model.normalize and tilt.secondary_model.normalize
    p = dict(mass_cs=np.zeros(12), q_cs=np.zeros(8), a1_cs=np.zeros(6), a2_cs=np.zeros(6), tilt1_cs=np.zeros(7), tilt2_cs=np.zeros(7), lamb=np.float64(2.0), z_cs=np.zeros(5))

    def w(d, pe_samples):
        return (mass(p["mass_cs"], p["q_cs"], pe_samples=pe_samples) * mag(p["a1_cs"], p["a2_cs"], pe_samples=pe_samples)
                * tilt(p["tilt1_cs"], p["tilt2_cs"], pe_samples=pe_samples) * zmod(d["redshift"], p["lamb"], p["z_cs"]) / d["prior"])

    low = lowering.lower(w(pe, True), w(inj, False))
    assert low.spec.n_params == 12 + 8 + 6 + 6 + 7 + 7 + 1 + 5
    blocks = pipeline.bspline_prior_blocks(low.slots_for, p)
    assert len(blocks) == 8
    by_start = {b[0].start: b for b in blocks}
    zb = by_start[low.slots_for(p["z_cs"]).start]
    assert zb[1:] == (1.0, 1.0, 2, True)
    mb = by_start[low.slots_for(p["mass_cs"]).start]
    assert mb[1:] == (15.0, 1.0, 1, False)
    iid_mag, iid_tilt = pipeline.setup_bspline_spin_models(pe, inj, 6, 7, IID=True)
    assert isinstance(iid_mag, M.BSplineIIDSpinMagnitudes) and isinstance(iid_tilt, M.BSplineIIDSpinTilts)
