"""CPU: the native NUTS driver (csrc/nuts.cpp).

* gwi_nuts_sample is host code and runs without a GPU: analytic targets through the callback API
  (moments of a correlated Gaussian, rejection of +inf regions, argument checks).
* gwi_posterior_* needs the likelihood, i.e. the device: here it runs on the host warp emulator
  (tests/emu) on a small catalog -- potential and gradient against the NumPy posterior
  (gwinferno_b200/nuts.py: BSplinePosterior) with the oracle-checked likelihood, then a short chain.
"""

import numpy as np
import pytest

from gwinferno_b200 import capi, nuts


@pytest.fixture(scope="module")
def lib():
    try:
        return capi.load_library()
    except capi.GwiError as e:
        pytest.skip(str(e))


def test_correlated_gaussian_moments(lib):
    rng = np.random.default_rng(0)
    dim = 6
    A = rng.standard_normal((dim, dim))
    cov = A @ A.T / dim + 0.5 * np.eye(dim)
    prec = np.linalg.inv(cov)
    mean = rng.standard_normal(dim)

    def U(th):
        d = th - mean
        g = prec @ d
        return 0.5 * d @ g, g

    samples, info = capi.nuts_sample(U, np.zeros(dim), 300, 3000, seed=3)
    assert samples.shape == (3000, dim)
    assert 0.6 < info["mean_accept"] < 0.98 and info["step_size"] > 0.0
    assert info["leapfrogs_total"] > info["leapfrogs_sampling"] > 3000
    ess = np.array([nuts.effective_sample_size(samples[:, i]) for i in range(dim)])
    assert ess.min() > 500
    err = np.abs(samples.mean(0) - mean) / np.sqrt(np.diag(cov) / ess)
    assert err.max() < 4.5  # standard errors
    assert np.allclose(np.cov(samples.T), cov, rtol=0.25, atol=0.08)


def test_same_seed_same_chain_and_infinite_regions_are_rejected(lib):
    def U(th):  # standard normal truncated to th[0] > -0.5 (outside: the likelihood's -inf sentinel)
        if th[0] <= -0.5:
            return np.inf, np.zeros_like(th)
        return 0.5 * th @ th, th

    a, _ = capi.nuts_sample(U, np.ones(3), 100, 400, seed=11)
    b, _ = capi.nuts_sample(U, np.ones(3), 100, 400, seed=11)
    c, _ = capi.nuts_sample(U, np.ones(3), 100, 400, seed=12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a[:, 0].min() > -0.5
    assert abs(a[:, 1].mean()) < 0.3 and 0.7 < a[:, 1].std() < 1.3


def test_argument_checks(lib):
    U = lambda th: (0.5 * th @ th, th)  # noqa: E731
    with pytest.raises(capi.GwiError):
        capi.nuts_sample(U, np.zeros(2), 10, 10, max_depth=0)
    with pytest.raises(capi.GwiError):
        capi.nuts_sample(U, np.zeros(2), 10, 10, target_accept=1.5)
    with pytest.raises(capi.GwiError, match="not finite"):
        capi.nuts_sample(lambda th: (np.inf, th), np.zeros(2), 10, 10)


# ---- the built-in posterior, on the emulated device -------------------------------------------------
@pytest.fixture(scope="module")
def small_posterior():
    from tests import emu

    try:
        emu.build()
    except Exception as e:
        pytest.skip(f"host emulator build failed: {e}")
    emu.activate()
    from gwinferno_b200 import pipeline, synthetic, workloads
    from gwinferno_b200.likelihood import PopulationLikelihood

    pe, inj, const = synthetic.make_catalog(6, 200, 6000, cfg=207)
    weights, params_fn = workloads.build_model("bspline", pe, inj, nsplines=dict(m1=10, q=8, a=6, t=6, z=7))
    low, lam, p = workloads.lower_workload(weights, params_fn, pe, inj, seed=1)
    eng = PopulationLikelihood(low, const["total_inj"])
    blocks = pipeline.bspline_prior_blocks(low.slots_for, p)
    yield dict(eng=eng, blocks=blocks, low=low, lam=lam, E=6)
    eng.model.close()
    emu.deactivate()


def test_native_posterior_equals_numpy_posterior(small_posterior):
    eng, blocks, low = small_posterior["eng"], small_posterior["blocks"], small_posterior["low"]
    ref = nuts.BSplinePosterior(lambda lam: eng.loglike(lam, Nobs=small_posterior["E"])[:2], low.spec.n_params, blocks)
    post = capi.Posterior(eng.model, blocks, small_posterior["E"])
    assert post.dim == ref.dim < low.spec.n_params  # the redshift spline's first coefficient is pinned
    rng = np.random.default_rng(1)
    for _ in range(3):
        th = 0.3 * rng.standard_normal(ref.dim)
        u0, g0 = ref(th)
        u1, g1 = post.potential(th)
        assert abs(u1 - u0) <= 1e-12 * abs(u0)
        assert np.max(np.abs(g1 - g0)) <= 1e-11 * np.max(np.abs(g0))
    post.close()


def test_native_chain_on_the_population_posterior(small_posterior):
    eng, blocks = small_posterior["eng"], small_posterior["blocks"]
    dim = capi.Posterior(eng.model, blocks, small_posterior["E"]).dim
    th0 = 0.05 * np.random.default_rng(2).standard_normal(dim)
    samples, info = nuts.nuts_native(eng, blocks, th0, 12, 12, Nobs=small_posterior["E"], seed=5, max_depth=3)
    assert samples.shape == (12, dim) and np.all(np.isfinite(samples))
    assert info["n_evals"] >= info["leapfrogs_total"] >= 24
    assert 0.2 < info["mean_accept"] <= 1.0
    assert np.std(samples, axis=0).min() > 0.0  # the chain moves in every coordinate
