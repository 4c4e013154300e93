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

import collections
import os
import warnings

import numpy as np

from . import capi, lowering

# device-resident plans, keyed by the uids of the model objects / sample arrays of a weight product
# (models._uid_of: never re-used, unlike id()); least-recently-used plans are closed when more than
# GWI_MAX_ENGINES (default 8) are alive -- a cfg-3 plan is 7.4 GB of device memory
_ENGINES = collections.OrderedDict()
_MAX_ENGINES = max(1, int(os.environ.get("GWI_MAX_ENGINES", "8")))
_BUILDS = collections.Counter()  # (term kinds, sharing pattern) -> plans built; warns when one model shape keeps missing the cache


class PopulationLikelihood:
    """Device-resident plan for one (catalog, population model) pair; evaluate for many Lambda."""

    def __init__(self, lowered, total_inj, device=0, need_neff_grad=False, chunk_steps=0, n_deep=-1, batch_hint=0, catalog_on_device=False):
        self.lowered = lowered
        self.spec = lowered.spec
        self.n_params = lowered.spec.n_params
        # catalog_on_device=True: the sample columns are placed in device memory first and the plan is built from there
        # (capi.Catalog(on_device=True)): the setting of a caller whose arrays already live on the GPU
        self.catalog = capi.Catalog(lowered.pe_cols, lowered.inj_cols, total_inj, device=device, on_device=catalog_on_device)
        self.model = capi.Model(self.catalog, lowered.spec, need_neff_grad=need_neff_grad, chunk_steps=chunk_steps, n_deep=n_deep, batch_hint=batch_hint)
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

    def last_sites(self):
        """Per-event sites of the LAST evaluation (analysis.py:260-264): ``logBFs``, ``log_nEffs``,
        ``variance_log_BFs`` [E] and ``variance_log_detection_efficiency``, read back from the device."""
        seg = self.model.last_sites()
        return {"logBFs": seg[1:, 0].copy(), "log_nEffs": seg[1:, 1].copy(), "variance_log_BFs": seg[1:, 2].copy(), "variance_log_detection_efficiency": float(seg[0, 2])}

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
        # per-event sites of analysis.py:260-264 (read back from the device results of this evaluation)
        for k in ("logBFs", "log_nEffs", "variance_log_BFs", "variance_log_detection_efficiency"):
            if k in head:
                self.sites[k] = head[k]
        if surveyed_hypervolume is not None:
            self.sites["surveyed_hypervolume"] = float(surveyed_hypervolume) / 1.0e9 * Tobs
        self._lowered = lowered
        self._pe_weights = pe_weights
        self.rate = None

    def grad(self, param):
        """Gradient of log_l with respect to a parameter OBJECT that was passed to the model calls
        (the total derivative: summed over every argument position the object was passed to)."""
        if self._lowered is None:  # cached engine: the object -> Lambda slots map of THIS call's parameter objects
            e = self._engine_lowered
            self._lowered = lowering.Lowered(e.spec, e.pe_cols, e.inj_cols, e.param_layout, lowering.object_slot_map(self._pe_weights))
        sl = self._lowered.all_slots_for(param)
        g = self.grad_flat[sl[0]]
        for extra in sl[1:]:
            g = g + self.grad_flat[extra]
        return g


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
    resampling (:321-355, with it ``param_names`` / ``pedata`` / ``injdata`` / ``m2min`` / ``m1min`` / ``mmax``)
    are outside the fused hot path and raise NotImplementedError.  ``reconstruct_rate`` (:265-268): NumPyro
    draws ``unscaled_rate ~ Gamma(Nobs)`` from the trace's key; here it is drawn from
    ``numpy.random.default_rng(rngkey)`` (``rngkey``: an int seed or ``None``) when a surveyed hypervolume is
    given, and returned -- as in the reference -- as the function's ``rate`` (``LikelihoodResult.rate``)."""
    if max_variance_cut and (marginalize_selection or min_neff_cut):
        raise ValueError(
            "max_variance_cut is True which requires marginalize_selection and min_neff_cut to be False but got "
            f"marginalize_selection = {marginalize_selection} and min_neff_cut = {min_neff_cut}"
        )
    if bool(log) != bool(getattr(pe_weights, "log_domain", False)) or bool(log) != bool(getattr(inj_weights, "log_domain", False)):
        raise ValueError("log=True expects log-weights (built from log_prob with + / -), log=False expects weights (built with * and /)")
    if categorical or marginal_qs:
        raise NotImplementedError("the categorical sub-population branch is not part of the fused path")
    if posterior_predictive_check:
        raise NotImplementedError("posterior-predictive resampling (analysis.py:321-355) is not part of the fused path: draw from the per-sample weights on the host")
    keys, pattern = lowering._structure(pe_weights, inj_weights)
    cache_key = (keys, pattern, float(total_inj), bool(marginalize_selection), int(device))
    eng = _ENGINES.get(cache_key)
    if eng is None:
        shape_key = (tuple(k[0] if isinstance(k[0], str) else k[1] for k in keys), pattern)
        _BUILDS[shape_key] += 1
        if _BUILDS[shape_key] == 4:
            warnings.warn("hierarchical_likelihood built a new device plan for the same model shape 4 times: a model object or sample array is "
                          "re-created on every call (construct the models once, outside the sampled function)", RuntimeWarning, stacklevel=2)
        eng = PopulationLikelihood.from_weights(pe_weights, inj_weights, total_inj, device=device, need_neff_grad=bool(marginalize_selection))
        _ENGINES[cache_key] = eng
        while len(_ENGINES) > _MAX_ENGINES:
            _, old = _ENGINES.popitem(last=False)
            old.model.close()
            old.catalog.close()
        lowered = eng.lowered
    else:
        _ENGINES.move_to_end(cache_key)
        # same static structure, new hyper-parameter objects: the slot map is rebuilt on demand (LikelihoodResult.grad)
        lowered = None
    lam = lowering.flatten_params(pe_weights, eng.n_params)
    log_l, grad, head = eng.loglike(lam, Nobs=Nobs, marginalize_selection=marginalize_selection, min_neff_cut=min_neff_cut, max_variance_cut=max_variance_cut)
    head = dict(head)
    head.update(eng.last_sites())
    # O(P) host glue: chain rule through parameter maps; per-sample constants kept off the device
    grad = lowering.pull_back(pe_weights, grad)
    log_l, grad, head = apply_host_norm(log_l, grad, head, *lowering.host_log_norm(pe_weights, eng.n_params), n_events=eng.n_events, Nobs=Nobs)
    res = LikelihoodResult(log_l, grad, head, lowered, pe_weights, surveyed_hypervolume, Tobs, Nobs)
    res._engine_lowered = eng.lowered
    if reconstruct_rate and surveyed_hypervolume is not None:
        # analysis.py:265-268
        total_vt = float(surveyed_hypervolume) / 1.0e9 * Tobs
        unscaled = float(np.random.default_rng(rngkey).gamma(Nobs))
        res.sites["unscaled_rate"] = unscaled
        res.rate = res.sites["rate"] = unscaled / float(np.exp(head["log_mu"])) / total_vt
    return res


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
    if "logBFs" in head:
        head["logBFs"] = head["logBFs"] - logZ
    if head["passed"]:
        log_l = log_l + (Nobs - n_events) * logZ
        grad = grad + (Nobs - n_events) * dlogZ
    return log_l, grad, head


def clear_cache():
    for e in _ENGINES.values():
        e.model.close()
        e.catalog.close()
    _ENGINES.clear()
