"""Minimal host-side NUTS driver over the fused likelihood.

The reference drives its likelihood with NumPyro's NUTS (examples/utils.py:63-84); NumPyro/JAX are
not available here, so the "NUTS ESS/s" half of the headline metric is measured with this small
NumPy implementation of the No-U-Turn sampler (Hoffman & Gelman 2014, Algorithm 6: dual-averaging
step size, slice variant) with a diagonal mass matrix adapted during warm-up.  Everything that is
O(n_samples) runs on the GPU through ``gwi_loglike_host``; the sampler state, the priors and the
P-spline smoothing penalty (gwinferno/models/bsplines/smoothing.py:8-28,
gwinferno/pipeline/utils.py:163-216) are O(P) host arithmetic, as in the reference.
"""

import time

import numpy as np

from . import capi


class BSplinePosterior:
    """Potential energy U(theta) = -(log L + log prior) of the cfg-2/3 B-spline model
    (examples/simple_bspline_example.py:25-89).  ``blocks``: list of
    ``(lambda_slice, prior_sigma, smoothing_tau or None, difference_degree, fix_first_zero)``."""

    def __init__(self, loglike_fn, n_params, blocks):
        self.loglike_fn = loglike_fn  # Lambda -> (log_l, grad)
        self.n_params = n_params
        self.blocks = blocks
        self.free = np.ones(n_params, dtype=bool)
        for sl, _, _, _, fix0 in blocks:
            if fix0:
                self.free[sl.start] = False  # z_cs = concatenate([0], sampled)  (pipeline/utils.py:213-214)
        self.dim = int(self.free.sum())
        self.n_evals = 0
        # per block Q = I / sigma^2 + tau D^T D (D = difference matrix), so that log prior = -c.Q.c / 2
        self._Q = []
        for sl, sig, tau, deg, _ in blocks:
            n = sl.stop - sl.start
            Q = np.eye(n) / sig**2
            if tau is not None and n > deg:
                D = np.diff(np.eye(n), n=deg, axis=0)
                Q = Q + tau * (D.T @ D)  # apply_difference_prior (smoothing.py:26-28)
            self._Q.append(Q)

    def to_lambda(self, theta):
        lam = np.zeros(self.n_params)
        lam[self.free] = theta
        return lam

    def log_prior(self, lam):
        lp, g = 0.0, np.zeros(self.n_params)
        for (sl, _, _, _, _), Q in zip(self.blocks, self._Q):
            c = lam[sl]
            qc = Q @ c
            lp += -0.5 * (c @ qc)
            g[sl] -= qc
        return lp, g

    def __call__(self, theta):
        lam = self.to_lambda(theta)
        log_l, grad = self.loglike_fn(lam)
        self.n_evals += 1
        lp, gp = self.log_prior(lam)
        if not np.isfinite(log_l) or log_l < -1e300:  # failed N_eff cut: the reference's -inf sentinel
            return np.inf, np.zeros(self.dim)
        return -(log_l + lp), -(grad + gp)[self.free]


def _leapfrog(U, theta, r, grad, eps, inv_mass):
    r = r - 0.5 * eps * grad
    theta = theta + eps * inv_mass * r
    u, grad = U(theta)
    r = r - 0.5 * eps * grad
    return theta, r, u, grad


def _find_reasonable_eps(U, theta, u0, g0, inv_mass, rng):
    eps = 0.05
    r = rng.standard_normal(theta.size) / np.sqrt(inv_mass)
    h0 = u0 + 0.5 * np.sum(inv_mass * r * r)
    _, r1, u1, _ = _leapfrog(U, theta, r, g0, eps, inv_mass)
    h1 = u1 + 0.5 * np.sum(inv_mass * r1 * r1)
    a = 1.0 if (np.isfinite(h1) and h0 - h1 > np.log(0.5)) else -1.0
    for _ in range(30):
        _, r1, u1, _ = _leapfrog(U, theta, r, g0, eps, inv_mass)
        h1 = u1 + 0.5 * np.sum(inv_mass * r1 * r1)
        if not np.isfinite(h1):
            h1 = np.inf
        if a * (h0 - h1) <= -a * np.log(2.0):
            break
        eps *= 2.0**a
    return eps


def nuts(U, theta0, n_warmup, n_samples, rng, target_accept=0.8, max_depth=8):
    """Returns ``(samples[n_samples, dim], info)``."""
    theta = np.array(theta0, dtype=np.float64)
    dim = theta.size
    inv_mass = np.ones(dim)
    u, grad = U(theta)
    eps = _find_reasonable_eps(U, theta, u, grad, inv_mass, rng)
    mu, gamma, t0, kappa = np.log(10.0 * eps), 0.05, 10.0, 0.75
    eps_bar, Hbar = (1.0 if n_warmup > 0 else eps), 0.0  # no warm-up: sample with the step size the search found
    adapt_origin = 0  # the dual averaging restarts after the mass-matrix update
    samples = np.zeros((n_samples, dim))
    warm = []
    n_leapfrog = 0
    accept_stat = []
    t_sampling = None

    def build_tree(theta, r, grad, logu, v, j, eps, h0):
        nonlocal n_leapfrog
        if j == 0:
            th1, r1, u1, g1 = _leapfrog(U, theta, r, grad, v * eps, inv_mass)
            n_leapfrog += 1
            h1 = u1 + 0.5 * np.sum(inv_mass * r1 * r1)
            if not np.isfinite(h1):
                h1 = np.inf
            n1 = int(logu <= -h1)
            s1 = int(logu < 1000.0 - h1)
            alpha = min(1.0, np.exp(min(0.0, h0 - h1))) if np.isfinite(h1) else 0.0
            return th1, r1, g1, th1, r1, g1, th1, g1, u1, n1, s1, alpha, 1
        thm, rm, gm, thp, rp, gp, th1, g1, u1, n1, s1, a1, na1 = build_tree(theta, r, grad, logu, v, j - 1, eps, h0)
        if s1 == 1:
            if v == -1:
                thm, rm, gm, _, _, _, th2, g2, u2, n2, s2, a2, na2 = build_tree(thm, rm, gm, logu, v, j - 1, eps, h0)
            else:
                _, _, _, thp, rp, gp, th2, g2, u2, n2, s2, a2, na2 = build_tree(thp, rp, gp, logu, v, j - 1, eps, h0)
            if n1 + n2 > 0 and rng.uniform() < n2 / (n1 + n2):
                th1, g1, u1 = th2, g2, u2
            a1 += a2
            na1 += na2
            dth = thp - thm
            s1 = s2 * int(dth @ (inv_mass * rm) >= 0) * int(dth @ (inv_mass * rp) >= 0)
            n1 += n2
        return thm, rm, gm, thp, rp, gp, th1, g1, u1, n1, s1, a1, na1

    for m in range(n_warmup + n_samples):
        if m == n_warmup:
            t_sampling = time.perf_counter()
            n_leapfrog_sampling0 = n_leapfrog
        r0 = rng.standard_normal(dim) / np.sqrt(inv_mass)
        h0 = u + 0.5 * np.sum(inv_mass * r0 * r0)
        logu = np.log(rng.uniform()) - h0
        thm = thp = theta
        rm = rp = r0
        gm = gp = grad
        j, n, s = 0, 1, 1
        step = eps if m < n_warmup else eps_bar
        while s == 1 and j < max_depth:
            v = -1 if rng.uniform() < 0.5 else 1
            if v == -1:
                thm, rm, gm, _, _, _, th1, g1, u1, n1, s1, a, na = build_tree(thm, rm, gm, logu, v, j, step, h0)
            else:
                _, _, _, thp, rp, gp, th1, g1, u1, n1, s1, a, na = build_tree(thp, rp, gp, logu, v, j, step, h0)
            if s1 == 1 and rng.uniform() < min(1.0, n1 / n):
                theta, grad, u = th1, g1, u1
            n += n1
            dth = thp - thm
            s = s1 * int(dth @ (inv_mass * rm) >= 0) * int(dth @ (inv_mass * rp) >= 0)
            j += 1
        acc = a / max(na, 1)
        if m < n_warmup:
            mm = m + 1 - adapt_origin  # the dual averaging restarts after the mass-matrix update
            Hbar = (1.0 - 1.0 / (mm + t0)) * Hbar + (target_accept - acc) / (mm + t0)
            eps = np.exp(mu - np.sqrt(mm) / gamma * Hbar)
            eta = mm ** (-kappa)
            eps_bar = np.exp(eta * np.log(eps) + (1.0 - eta) * np.log(eps_bar))
            warm.append(theta.copy())
            # one mass-matrix update in the middle of warm-up (diagonal, regularised sample variance)
            if m + 1 == n_warmup // 2 and len(warm) >= 20:
                w = np.array(warm[len(warm) // 4 :])
                var = np.var(w, axis=0)
                nn = w.shape[0]
                inv_mass = (nn / (nn + 5.0)) * var + 1e-3 * (5.0 / (nn + 5.0))
                u, grad = U(theta)
                eps = _find_reasonable_eps(U, theta, u, grad, inv_mass, rng)
                mu = np.log(10.0 * eps)
                eps_bar, Hbar = 1.0, 0.0
                adapt_origin = m + 1
        else:
            samples[m - n_warmup] = theta
            accept_stat.append(acc)
    dt = time.perf_counter() - t_sampling
    info = dict(step_size=float(eps_bar), mean_accept=float(np.mean(accept_stat)), sampling_seconds=dt,
                leapfrogs_sampling=int(n_leapfrog - n_leapfrog_sampling0), leapfrogs_total=int(n_leapfrog))
    return samples, info


def nuts_native(engine, blocks, theta0, n_warmup, n_samples, Nobs=None, seed=0, target_accept=0.8, max_depth=8, flags=0, **like_kw):
    """The same sampler with the whole transition loop in native code (csrc/nuts.cpp,
    gwi_nuts_sample_posterior): ``engine`` is a :class:`~gwinferno_b200.likelihood.PopulationLikelihood`,
    ``blocks`` as for :class:`BSplinePosterior`.  One gwi_loglike_host call per leapfrog step and no
    interpreter in between -- use this when the likelihood is cheaper than ~1 ms.  ``flags``: 0 = this
    module's algorithm; ``capi.NUTS_MULTINOMIAL | capi.NUTS_WINDOWED_ADAPT | capi.NUTS_DENSE_MASS``
    select multinomial trajectory sampling, Stan-style windowed warm-up and a dense mass matrix.  Returns
    ``(samples[n_samples, dim], info)``; ``info`` has the keys of :func:`nuts` plus ``n_evals``."""
    post = capi.Posterior(engine.model, blocks, engine.n_events if Nobs is None else Nobs, **like_kw)
    try:
        return post.sample(theta0, n_warmup, n_samples, seed=seed, target_accept=target_accept, max_depth=max_depth, flags=flags)
    finally:
        post.close()


def nuts_native_chains(engine, blocks, theta0, n_warmup, n_samples, Nobs=None, seed=0, target_accept=0.8, max_depth=8, flags=0, **like_kw):
    """``theta0[n_chains, dim]``: that many independent chains of :func:`nuts_native`, advanced together -- whenever all
    chains stand at a leapfrog step their Lambda vectors are evaluated in ONE batched GPU call (gwi_loglike_batch_host),
    the reference's ``MCMC(..., chain_method="vectorized")`` on one GPU.  Chain ``c`` draws what :func:`nuts_native`
    draws with ``seed + c``.  Create ``engine`` with ``batch_hint = n_chains``.  Returns
    ``(samples[n_chains, n_samples, dim], [info per chain])``."""
    post = capi.Posterior(engine.model, blocks, engine.n_events if Nobs is None else Nobs, **like_kw)
    try:
        return post.sample_chains(theta0, n_warmup, n_samples, seed=seed, target_accept=target_accept, max_depth=max_depth, flags=flags)
    finally:
        post.close()


def split_rhat(chains):
    """Split-R-hat of ``chains[n_chains, n_samples]`` (Gelman et al. 2013): every chain cut in halves, between- over
    within-half variance.  ~1 when the chains sample the same distribution."""
    x = np.asarray(chains, dtype=np.float64)
    n = x.shape[1] // 2
    if n < 2:
        return float("nan")
    h = np.concatenate([x[:, :n], x[:, n : 2 * n]], axis=0)
    w = h.var(axis=1, ddof=1).mean()
    b = n * h.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * w + b / n) / w)) if w > 0 else float("nan")


def effective_sample_size(x):
    """ESS of a 1-D chain (Geyer's initial positive sequence on the FFT autocorrelation)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    x = x - x.mean()
    if n < 4 or not np.any(x):
        return float(n)
    f = np.fft.rfft(x, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n].real
    acf /= acf[0]
    s = 0.0
    for k in range(1, n - 1, 2):
        pair = acf[k] + acf[k + 1]
        if pair < 0:
            break
        s += pair
    tau = 1.0 + 2.0 * s - 0.0
    return float(n / max(tau, 1e-12))
