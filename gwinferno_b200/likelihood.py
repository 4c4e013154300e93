"""Reference-facing front-end of the fused population likelihood.

:func:`hierarchical_likelihood` keeps the signature of the reference's
``gwinferno.pipeline.analysis.hierarchical_likelihood`` (analysis.py:139-163) but takes the LAZY
weights produced by the mirror model classes (gwinferno_b200/models.py) and evaluates the whole
path -- per-sample densities, per-event and injection Monte-Carlo sums, N_eff / variance cuts and
the gradient with respect to every hyper-parameter -- in one call into libgwi.so.  Without
NumPyro in the loop the "sites" the reference records with ``numpyro.deterministic`` /
``numpyro.factor`` (analysis.py:260-319) are returned in a :class:`LikelihoodResult`.

There is no CPU fallback: a missing library or GPU raises :class:`gwinferno_b200.capi.GwiError`.
"""

import numpy as np

from . import capi, lowering

_ENGINES = {}


class PopulationLikelihood:
    """Device-resident plan for one (catalog, population model) pair; evaluate for many Lambda."""

    def __init__(self, lowered, total_inj, device=0, need_neff_grad=False, chunk_steps=0, n_deep=-1):
        self.lowered = lowered
        self.spec = lowered.spec
        self.n_params = lowered.spec.n_params
        self.catalog = capi.Catalog(lowered.pe_cols, lowered.inj_cols, total_inj, device=device)
        self.model = capi.Model(self.catalog, lowered.spec, need_neff_grad=need_neff_grad, chunk_steps=chunk_steps, n_deep=n_deep)
        self.need_neff_grad = need_neff_grad
        self.n_events = self.catalog.n_events

    @classmethod
    def from_weights(cls, pe_weights, inj_weights, total_inj, **kw):
        return cls(lowering.lower(pe_weights, inj_weights), total_inj, **kw)

    def evaluate(self, lam, jacobians=True):
        """``logBF[E], logNeff[E], log_mu, logNeff_inj`` (+ Jacobians) -- the outputs of the
        reference's per_event_log_bayes_factors / detection_efficiency (analysis.py:50-136)."""
        return self.model.evaluate(lam, jacobians=jacobians)

    def loglike(self, lam, Nobs=None, marginalize_selection=False, min_neff_cut=True, max_variance_cut=False):
        """``(log_l, dlog_l/dLambda, diagnostics)`` through gwi_loglike_host (host buffers)."""
        if marginalize_selection and not self.need_neff_grad:
            raise ValueError("marginalize_selection=True needs PopulationLikelihood(..., need_neff_grad=True)")
        head, grad = self.model.loglike_host(
            lam, self.n_events if Nobs is None else Nobs, marginalize_selection=marginalize_selection, min_neff_cut=min_neff_cut, max_variance_cut=max_variance_cut
        )
        return head["log_l"], grad, head

    def info(self):
        return self.model.info()


class LikelihoodResult:
    """What the reference's hierarchical_likelihood leaves in the NumPyro trace (analysis.py:260-319)."""

    def __init__(self, log_l, grad_flat, head, lowered, pe_weights, surveyed_hypervolume, Tobs, Nobs):
        self.log_likelihood = float(log_l)
        self.grad_flat = grad_flat
        self.sites = {
            "log_l": float(log_l),
            "log_nEff_inj": head["logNeff_inj"],
            "detection_efficiency": float(np.exp(head["log_mu"])),
            "sum_logBFs": head["sum_logBF"],
            "variance_log_likelihood": head["variance"],
            "passed_cuts": bool(head["passed"]),
        }
        if surveyed_hypervolume is not None:
            self.sites["surveyed_hypervolume"] = float(surveyed_hypervolume) / 1.0e9 * Tobs
        self._lowered = lowered
        self._pe_weights = pe_weights
        self.rate = None

    def grad(self, param):
        """Gradient of log_l with respect to a parameter OBJECT that was passed to the model calls."""
        return self.grad_flat[self._lowered.slots_for(param)]


def hierarchical_likelihood(
    pe_weights,
    inj_weights,
    total_inj,
    Nobs,
    Tobs,
    surveyed_hypervolume=None,
    categorical=False,
    marginal_qs=False,
    indv_weights=None,
    rngkey=None,
    pop_frac=None,
    reconstruct_rate=True,
    marginalize_selection=False,
    min_neff_cut=True,
    max_variance_cut=False,
    posterior_predictive_check=False,
    param_names=None,
    pedata=None,
    injdata=None,
    m2min=3.0,
    m1min=5.0,
    mmax=100.0,
    log=False,
    device=0,
):
    """Drop-in for ``gwinferno.pipeline.analysis.hierarchical_likelihood`` (analysis.py:139-356)
    on lazy weights.  ``log=True`` takes lazy LOG-weights (``log_prob`` terms combined with ``+`` and
    ``- jnp.log(prior)``, analysis.py:401-421); the fused path always works in log space, so both forms
    lower to the same device model.  The categorical sub-population branch (:246-254) and posterior-predictive
    resampling (:321-355) are outside the fused hot path and raise NotImplementedError."""
    if max_variance_cut and (marginalize_selection or min_neff_cut):
        raise ValueError(
            "max_variance_cut is True which requires marginalize_selection and min_neff_cut to be False but got "
            f"marginalize_selection = {marginalize_selection} and min_neff_cut = {min_neff_cut}"
        )
    if bool(log) != bool(getattr(pe_weights, "log_domain", False)) or bool(log) != bool(getattr(inj_weights, "log_domain", False)):
        raise ValueError("log=True expects log-weights (built from log_prob with + / -), log=False expects weights (built with * and /)")
    if categorical or marginal_qs:
        raise NotImplementedError("the categorical sub-population branch is not part of the fused path")
    keys, pattern = lowering._structure(pe_weights, inj_weights)
    cache_key = (keys, pattern, float(total_inj), bool(marginalize_selection), int(device))
    eng = _ENGINES.get(cache_key)
    if eng is None:
        eng = PopulationLikelihood.from_weights(pe_weights, inj_weights, total_inj, device=device, need_neff_grad=bool(marginalize_selection))
        _ENGINES[cache_key] = eng
        lowered = eng.lowered
    else:
        # same static structure, new hyper-parameter objects: rebuild only the slot map
        lowered = lowering.Lowered(eng.spec, eng.lowered.pe_cols, eng.lowered.inj_cols, eng.lowered.param_layout, _slot_map(pe_weights))
    lam = lowering.flatten_params(pe_weights, eng.n_params)
    log_l, grad, head = eng.loglike(lam, Nobs=Nobs, marginalize_selection=marginalize_selection, min_neff_cut=min_neff_cut, max_variance_cut=max_variance_cut)
    # O(P) host glue: chain rule through parameter maps; per-sample constants kept off the device
    grad = lowering.pull_back(pe_weights, grad)
    log_l, grad, head = apply_host_norm(log_l, grad, head, *lowering.host_log_norm(pe_weights, eng.n_params), n_events=eng.n_events, Nobs=Nobs)
    return LikelihoodResult(log_l, grad, head, lowered, pe_weights, surveyed_hypervolume, Tobs, Nobs)


def apply_host_norm(log_l, grad, head, logZ, dlogZ, n_events, Nobs):
    """Every sample weight is ``exp(-logZ)`` times what the device model evaluated: shift the sites
    (``logBF_i`` and ``log mu`` by ``-logZ``; N_eff and the variances are scale-free) and the
    likelihood ``log_l = sum_i logBF_i - Nobs log mu`` (analysis.py:257-319) by ``(Nobs - E) logZ``
    -- nothing when ``Nobs`` equals the number of events, as in every reference example."""
    if logZ == 0.0 and not np.any(dlogZ):
        return log_l, grad, head
    head = dict(head)
    head["log_mu"] = head["log_mu"] - logZ
    head["sum_logBF"] = head["sum_logBF"] - n_events * logZ
    if head["passed"]:
        log_l = log_l + (Nobs - n_events) * logZ
        grad = grad + (Nobs - n_events) * dlogZ
    return log_l, grad, head


def _slot_map(pe_w):
    slot_of, off = {}, 0
    for t in pe_w.terms:
        for p, k in zip(t.params, t.param_keys):
            if k not in slot_of:
                slot_of[k] = off
                off += p.size
    return slot_of


def clear_cache():
    for e in _ENGINES.values():
        e.model.close()
        e.catalog.close()
    _ENGINES.clear()
