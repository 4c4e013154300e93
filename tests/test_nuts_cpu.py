"""CPU: the host NUTS driver on an analytic target (no GPU, no oracle)."""

import numpy as np

from gwinferno_b200 import nuts


def test_nuts_samples_a_gaussian():
    rng = np.random.default_rng(0)
    sig = np.array([0.5, 1.0, 3.0, 0.2])

    def U(theta):
        return 0.5 * np.sum((theta / sig) ** 2), theta / sig**2

    samples, info = nuts.nuts(U, np.ones(4), n_warmup=300, n_samples=600, rng=rng)
    assert 0.6 < info["mean_accept"] <= 1.0
    assert np.all(np.abs(samples.mean(axis=0)) < 4 * sig / np.sqrt(100))
    assert np.allclose(samples.std(axis=0), sig, rtol=0.2)
    ess = [nuts.effective_sample_size(samples[:, i]) for i in range(4)]
    assert min(ess) > 100


def test_posterior_wrapper_gradient_matches_finite_difference():
    rng = np.random.default_rng(1)
    P = 12
    A = rng.standard_normal((P, P))
    Q = A @ A.T / P + np.eye(P)

    def loglike(lam):
        return -0.5 * lam @ Q @ lam, -Q @ lam

    blocks = [(slice(0, 6), 5.0, 2.0, 1, False), (slice(6, 11), 1.0, 1.5, 2, True), (slice(11, 12), 3.0, None, 0, False)]
    post = nuts.BSplinePosterior(loglike, P, blocks)
    assert post.dim == P - 1
    th = rng.standard_normal(post.dim)
    u, g = post(th)
    for i in range(post.dim):
        e = np.zeros(post.dim)
        e[i] = 1e-6
        fd = (post(th + e)[0] - post(th - e)[0]) / 2e-6
        assert abs(fd - g[i]) < 1e-5 * max(1.0, abs(g[i]))
