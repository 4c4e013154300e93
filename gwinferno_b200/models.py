"""Host-side mirror of the reference's population-model classes (drop-in call surface).

Same class names, constructor signatures and ``__call__`` arguments as
``gwinferno/models/bsplines/{single,separable}.py``, ``gwinferno/models/spline_perturbation.py``
and ``gwinferno/models/parametric/parametric.py`` -- but instead of building dense design
matrices (single.py:56-57) and returning a dense pdf array, ``__call__`` returns a
:class:`LazyWeight`: a symbolic product of per-dimension terms that
:func:`gwinferno_b200.likelihood.hierarchical_likelihood` lowers to ONE fused CUDA evaluation
through the C-ABI (include/gwi.h).  ``LazyWeight`` supports ``*`` and ``/`` exactly as the
reference's weight assembly uses them (examples/simple_bspline_example.py:58-68).

Nothing here evaluates a density on the CPU: there is no CPU fallback.
"""

import collections
import hashlib
import itertools

import numpy as np

from . import spec as S
from .cosmology import Planck15

__all__ = [
    "BSpline",
    "LogXBSpline",
    "LogYBSpline",
    "LogXLogYBSpline",
    "LazyWeight",
    "Base1DBSplineModel",
    "BSplineSpinMagnitude",
    "BSplineSpinTilt",
    "BSplineRatio",
    "BSplineMass",
    "BSplineChiEffective",
    "BSplineSymmetricChiEffective",
    "BSplineChiPrecess",
    "BSplineRedshift",
    "BSplineEffectiveSpinDims",
    "BSplineIIDSpinMagnitudes",
    "BSplineIndependentSpinMagnitudes",
    "BSplineIIDSpinTilts",
    "BSplineIndependentSpinTilts",
    "BSplinePrimaryBSplineRatio",
    "BSplinePrimaryPowerlawRatio",
    "PLPeakPrimaryBSplineRatio",
    "BSplineIIDComponentMasses",
    "BSplineIndependentComponentMasses",
    "PowerlawRedshiftModel",
    "PowerlawSplineRedshiftModel",
    "powerlaw_primary_ratio_pdf",
    "plpeak_primary_ratio_pdf",
    "plpeak_primary_pdf",
    "beta_spin_magnitude",
    "iid_spin_magnitude",
    "independent_spin_magnitude_beta_dist",
    "mixture_isoalign_spin_tilt",
    "iid_spin_tilt",
    "default_spin_tilt",
    "independent_spin_tilt",
    "weight_from_prior",
]


# ---- basis "classes": only used as tags selecting the projection, like the reference's
#      ``basis=`` keyword (single.py:42; interpolation.py:240,320,360,410) -------------------
class BSpline:
    logx, logy, default_normalize, n_grid = False, False, False, 1000


class LogXBSpline(BSpline):
    logx, logy, default_normalize, n_grid = True, False, True, 1000


class LogYBSpline(BSpline):
    logx, logy, default_normalize, n_grid = False, True, True, 1000


class LogXLogYBSpline(LogYBSpline):
    logx, logy, default_normalize, n_grid = True, True, True, 1500


# ---- stable identities ---------------------------------------------------------------------------
# The static plan (device-resident, expensive) is cached by WHICH model objects / sample arrays a
# weight product is built from.  ``id()`` alone is not an identity: CPython re-uses the address of a
# collected object, so a new model over a different catalog could hit the old plan.  Every keyed
# object therefore gets a process-unique, never re-used uid: model objects carry it as an attribute,
# arrays (which take no attributes) are looked up in a bounded registry that holds a strong
# reference while the entry lives (an evicted array that shows up again simply gets a new uid: a
# cache miss, never a wrong hit).
_UID = itertools.count(1)
_ARRAY_UIDS = collections.OrderedDict()  # id(array) -> (array, uid, fingerprint), LRU
_ARRAY_BY_PRINT = {}  # fingerprint -> id(array) of the registered representative
_ARRAY_UIDS_MAX = 256


def _fingerprint(a):
    """Cheap content fingerprint of an array (shape, dtype, <= 4096 strided elements)."""
    flat = a.reshape(-1) if a.flags.c_contiguous else np.ravel(a)
    n = flat.size
    probe = flat[:: max(1, n // 4096)][:4096] if n else flat
    return (a.shape, a.dtype.str, hashlib.blake2b(np.ascontiguousarray(probe).tobytes(), digest_size=16).digest())


def _uid_of(obj):
    """Process-unique identity of a keyed object.  Arrays: the SAME object, or an array with EQUAL
    content (an array the user recomputes on every call, e.g. ``jnp.log(prior)`` or ``m1 * q`` inside
    the model function, analysis.py:401-402), maps to the same uid; the content comparison only runs
    on an identity miss and is exact (fingerprint first, then ``np.array_equal``)."""
    if obj is None:
        return 0
    d = getattr(obj, "__dict__", None)
    if d is not None:
        u = d.get("_gwi_uid")
        if u is None:
            u = d["_gwi_uid"] = next(_UID)
        return u
    key = id(obj)
    hit = _ARRAY_UIDS.get(key)
    if hit is not None and hit[0] is obj:
        _ARRAY_UIDS.move_to_end(key)
        return hit[1]
    fp = _fingerprint(obj) if isinstance(obj, np.ndarray) else None
    if fp is not None:
        rep = _ARRAY_UIDS.get(_ARRAY_BY_PRINT.get(fp))
        if rep is not None and rep[2] == fp and np.array_equal(rep[0], obj, equal_nan=True):
            return rep[1]  # equal content: the representative stays the registered one (nothing new is pinned)
    u = next(_UID)
    _ARRAY_UIDS[key] = (obj, u, fp)
    if fp is not None:
        _ARRAY_BY_PRINT[fp] = key
    while len(_ARRAY_UIDS) > _ARRAY_UIDS_MAX:
        k_old, (_, _, fp_old) = _ARRAY_UIDS.popitem(last=False)
        if fp_old is not None and _ARRAY_BY_PRINT.get(fp_old) == k_old:
            del _ARRAY_BY_PRINT[fp_old]
    return u


def _share():
    """Token for a hyper-parameter that one model function hands to several of its terms on purpose
    (``iid_*``: one Lambda slot, gradient = total derivative)."""
    return ("shared", next(_UID))


def is_device_array(a):
    """Sample arrays may live in DEVICE memory: any object exposing ``__cuda_array_interface__`` (jax / torch / cupy arrays,
    ``capi.DeviceArray``).  They flow through the model classes untouched and reach the library as device pointers
    (``gwi_catalog_desc.columns_on_device``): the plan is then built on the GPU without a host round trip."""
    return not isinstance(a, np.ndarray) and hasattr(a, "__cuda_array_interface__")


def _as_samples(a):
    if is_device_array(a):
        ci = a.__cuda_array_interface__
        if ci.get("typestr") != "<f8" or ci.get("strides") is not None:
            raise ValueError("device-resident sample arrays must be C-contiguous float64")
        return a
    return np.asarray(a, dtype=np.float64)


class _Column:
    """A (PE array, injection array) pair of one physical sample coordinate."""

    def __init__(self, pe, inj):
        self.pe = None if pe is None else _as_samples(pe)
        self.inj = None if inj is None else _as_samples(inj)


_NO_MAPS = {}  # shared, never mutated


class _LazyTerm:
    """One additive log-density term; ``params`` are the hyper-parameter VALUES of this call,
    ``build(slots, group_base)`` produces the static spec entries (terms, groups, cuts)."""

    def __init__(self, key, columns, params, build, maps=None, host_norm=None, param_keys=None):
        # ``maps``: {index in params: ParamMap} -- an O(P) host-side change of variables applied when
        # Lambda is assembled; gradients are pulled back through it (lowering.pull_back)
        self.maps = dict(maps) if maps else _NO_MAPS
        # ``host_norm(params) -> (log Z, [dlog Z/dparam_i])``: a per-sample CONSTANT -log Z that is
        # not part of the device model (it cancels in log L when Nobs = number of events); the
        # front-end applies it to the reported sites (lowering.host_log_norm)
        self.host_norm = host_norm
        self.key = key  # hashable identity of the STATIC part (model object id, dimension)
        self.columns = columns  # list[_Column]
        # parameter OBJECTS passed by the caller.  An ARRAY given to two terms (IID models pass one
        # coefficient vector to two dimensions, separable.py:77-79) shares its Lambda slots.  Scalars
        # never share by identity -- equal Python floats are routinely the same object (literals,
        # small ints, one variable passed twice) without meaning one parameter -- every scalar
        # argument position is its own slot unless the calling model function shares it explicitly
        # (``param_keys`` from ``_share()``); gradients come back per object as total derivatives
        # (``LikelihoodResult.grad``).  Keys only need to be unique within one lowering call; the
        # objects stay alive in ``_orig`` for that long.
        self._orig = list(params)
        if param_keys is None:
            self.param_keys = [("arr", id(p)) if (p.ndim if type(p) is np.ndarray else np.ndim(p)) >= 1 else ("pos", next(_UID)) for p in self._orig]
        else:
            self.param_keys = [k if k is not None else (("arr", id(p)) if np.ndim(p) >= 1 else ("pos", next(_UID))) for p, k in zip(self._orig, param_keys)]
        # (fast path: coefficient vectors arrive as float64 arrays on every sampler step)
        self.params = [p if (type(p) is np.ndarray and p.dtype == np.float64 and p.ndim >= 1) else np.atleast_1d(np.asarray(p, dtype=np.float64)) for p in self._orig]
        self.build = build


class ParamMap:
    """Host-side change of variables for one parameter block: ``fwd(value) -> value seen by the
    kernel`` and ``vjp(value, grad_wrt_mapped) -> grad_wrt_value``."""

    def __init__(self, fwd, vjp):
        self.fwd, self.vjp = fwd, vjp


_EXP_CACHE = collections.OrderedDict()  # uid of a log-prior array -> exp of it (stable identity for the plan cache)


class LazyWeight:
    """Symbolic per-sample weight  prod_d p_d(theta_d; Lambda) / prior  for one sample set
    (``pe_samples=True`` or ``False``)."""

    def __init__(self, terms, pe_samples, log_domain=False):
        self.terms = list(terms)
        self.pe_samples = pe_samples
        # log_domain: the object stands for the LOG of the weight (what the reference passes with
        # ``hierarchical_likelihood(..., log=True)``, analysis.py:401-421): factors combine with + / -
        self.log_domain = bool(log_domain)

    def _merge(self, other):
        if self.pe_samples is None:
            return other.pe_samples
        if other.pe_samples is None or other.pe_samples == self.pe_samples:
            return self.pe_samples
        raise ValueError("cannot mix PE-sample and injection weights in one product")

    def __mul__(self, other):
        if isinstance(other, LazyWeight):
            if self.log_domain or other.log_domain:
                raise TypeError("log-weights combine with + and -, not *")
            return LazyWeight(self.terms + other.terms, self._merge(other))
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, LazyWeight):
            raise TypeError("division by a population term is not supported")
        if self.log_domain:
            raise TypeError("log-weights combine with + and -, not /")
        # division by the sampling-prior array (examples/simple_bspline_example.py:66)
        arr = _as_samples(other)
        return self * weight_from_prior(arr, pe_samples=self.pe_samples)

    def __add__(self, other):
        if isinstance(other, LazyWeight) and self.log_domain and other.log_domain:
            return LazyWeight(self.terms + other.terms, self._merge(other), log_domain=True)
        if isinstance(other, LazyWeight):
            raise TypeError("only log-weights (log_prob) can be added")
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        """``log_weight - jnp.log(prior)`` (analysis.py:401-402)."""
        if not self.log_domain or isinstance(other, LazyWeight):
            raise TypeError("only the log of the sampling prior can be subtracted, and only from a log-weight")
        log_prior = np.asarray(other, dtype=np.float64)
        u = _uid_of(log_prior)
        prior = _EXP_CACHE.get(u)
        if prior is None:
            prior = _EXP_CACHE[u] = np.exp(log_prior)
            while len(_EXP_CACHE) > 16:
                _EXP_CACHE.popitem(last=False)
        w = weight_from_prior(prior, pe_samples=self.pe_samples)
        return LazyWeight(self.terms + w.terms, self._merge(w), log_domain=True)


def weight_from_prior(prior, pe_samples=None):
    """``1/prior`` as a lazy static term.  ``pe_samples`` is inferred from the dimensionality when
    omitted (2-D = PE samples, 1-D = injections; same convention as parametric.py:130)."""
    prior = _as_samples(prior)
    if pe_samples is None:
        pe_samples = len(prior.shape) == 2
    col = _Column(prior if pe_samples else None, None if pe_samples else prior)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_STATIC, [cols[0]], feature=S.FEAT_NEG_LOG, name="1/prior")], [], []

    return LazyWeight([_LazyTerm(("prior", _uid_of(prior)), [col], [], build)], pe_samples)


def _is_pe(arr):
    return (len(arr.shape) if is_device_array(arr) else np.ndim(arr)) == 2


# ================================================================================================
# 1-D B-spline models (gwinferno/models/bsplines/single.py)
# ================================================================================================
class Base1DBSplineModel:
    """Mirror of ``Base1DBSplineModel`` (single.py:16-128): holds the static samples and the
    basis description; ``__call__(coefs, pe_samples)`` returns a :class:`LazyWeight`."""

    def __init__(self, n_splines, xx, xx_inj, xrange=(0.0, 1.0), degree=3, basis=BSpline, **kwargs):
        if not 0 <= int(degree) <= 3:
            raise NotImplementedError("B-spline degrees 0..3 are implemented on the CUDA path (4 coefficients per polynomial piece)")
        self.n_splines = int(n_splines)
        self.xmin, self.xmax = float(xrange[0]), float(xrange[1])
        self.degree = degree
        self.basis = basis
        self.normalize = bool(kwargs.get("normalize", basis.default_normalize))
        self.column = _Column(xx, xx_inj)
        # spline-coordinate range and the normalisation grid, exactly as the reference builds them
        if basis.logx:
            self.xi_range = (float(np.log(self.xmin)), float(np.log(self.xmax)))  # interpolation.py:428
            grid = np.linspace(*np.exp(np.array(self.xi_range)), basis.n_grid)  # :433
            grid_xi = np.log(grid)  # :447
        else:
            self.xi_range = (self.xmin, self.xmax)
            grid = np.linspace(self.xmin, self.xmax, basis.n_grid)  # interpolation.py:378
            grid_xi = grid
        # explicit knot vector / order: built exactly as BasisSpline.__init__ does (interpolation.py:94-106; the log-x
        # bases take the logarithm of user-given knots first, :335-338,425-428).  ``interior_knots`` only sets the COUNT of
        # knots and the spacing of the extension (:99-101) -- the reference lays a uniform vector over the extended range.
        self.knots, self.order = None, int(degree) + 1
        knots, interior = kwargs.get("knots"), kwargs.get("interior_knots")
        if int(degree) != 3 or knots is not None or interior is not None:
            k = self.order
            knots = None if knots is None else np.asarray(knots, dtype=np.float64)
            interior = None if interior is None else np.asarray(interior, dtype=np.float64)
            if basis.logx:
                knots = None if knots is None else np.log(knots)
                interior = None if interior is None else np.log(interior)
            if knots is None:
                if interior is None:
                    interior = np.linspace(self.xi_range[0], self.xi_range[1], self.n_splines - k + 2)
                dx = interior[1] - interior[0]
                knots = np.linspace(self.xi_range[0] - dx * (k - 1), self.xi_range[1] + dx * (k - 1), len(interior) + (k - 1) * 2)
            if len(knots) != self.n_splines + k:
                raise AssertionError("len(knots) must be n_splines + degree + 1 (interpolation.py:106)")
            if self.n_splines < 4:
                raise NotImplementedError("the CUDA path needs at least 4 basis functions per spline")
            self.knots = np.ascontiguousarray(knots, dtype=np.float64)
        self.grid = grid
        inside = (grid_xi >= self.xi_range[0]) & (grid_xi <= self.xi_range[1])
        with np.errstate(divide="ignore"):
            log_w = np.log(S.trapezoid_weights(grid))
        # LogY bases are -inf outside the range => integrand exp(-inf) = 0 (interpolation.py:393-394,449)
        self.grid_log_w = np.where(inside, log_w, -np.inf)
        self.grid_xi = np.clip(grid_xi, *self.xi_range)

    def _lazy(self, coefs, pe_samples, tag=""):
        if (coefs.shape if type(coefs) is np.ndarray else np.shape(coefs)) != (self.n_splines,):
            raise ValueError(f"expected {self.n_splines} coefficients, got shape {np.shape(coefs)}")
        static = self.__dict__.setdefault("_lazy_static", {}).get(tag)
        if static is None:  # the static part of the term (identity key, column, spec builder): once per model and tag
            static = self._lazy_static[tag] = ((_uid_of(self), tag), [self.column], self._make_build())
        return LazyWeight([_LazyTerm(static[0], static[1], [coefs], static[2])], pe_samples)

    def _make_build(self):
        model = self

        def build(slots, group_base, cols):
            groups = []
            g = -1
            if model.normalize:
                groups = [S.NormGroup(model.grid_log_w, name=f"Z[{cols[0]}]")]
                g = group_base
            term = S.Term(
                # LogY bases: exp(B.c) (interpolation.py:381-394); BSpline/LogXBSpline: B.c itself (:293-317)
                S.TERM_SPLINE if model.basis.logy else S.TERM_SPLINE_LINEAR,
                [cols[0]],
                slots=[slots[0]],
                n_splines=model.n_splines,
                logx=model.basis.logx,
                outside=S.OUTSIDE_DROP,
                xrange=(model.xmin, model.xmax),
                norm_group=g,
                grid_xi=model.grid_xi if model.normalize else None,
                name=f"spline[{cols[0]}]",
            )
            term.xi_range = model.xi_range
            term.knots, term.order = model.knots, model.order
            return [term], groups, []

        return build

    def __call__(self, coefs, pe_samples=True):
        return self._lazy(coefs, pe_samples)


class BSplineSpinMagnitude(Base1DBSplineModel):
    def __init__(self, n_splines, a, a_inj, basis=LogYBSpline, **kwargs):
        xrange = kwargs.pop("xrange", (0.0, 1.0))
        super().__init__(n_splines, a, a_inj, basis=basis, xrange=xrange, **kwargs)


class BSplineSpinTilt(Base1DBSplineModel):
    def __init__(self, n_splines, ct, ct_inj, basis=LogYBSpline, **kwargs):
        xrange = kwargs.pop("xrange", (-1.0, 1.0))
        super().__init__(n_splines, ct, ct_inj, basis=basis, xrange=xrange, **kwargs)


class BSplineRatio(Base1DBSplineModel):
    def __init__(self, n_splines, q, q_inj, qmin=0, basis=LogYBSpline, **kwargs):
        xrange = kwargs.pop("xrange", (qmin, 1))
        super().__init__(n_splines, q, q_inj, basis=basis, xrange=xrange, **kwargs)


class BSplineMass(Base1DBSplineModel):
    def __init__(self, n_splines, m, m_inj, mmin=2, mmax=100, basis=LogXLogYBSpline, **kwargs):
        xrange = kwargs.pop("xrange", (mmin, mmax))
        super().__init__(n_splines, m, m_inj, basis=basis, xrange=xrange, **kwargs)


class BSplineChiEffective(Base1DBSplineModel):
    """single.py:199-230 (default basis: the spline itself is the density)."""

    def __init__(self, n_splines, chieff, chieff_inj, basis=BSpline, **kwargs):
        xrange = kwargs.pop("xrange", (-1.0, 1.0))
        super().__init__(n_splines, chieff, chieff_inj, basis=basis, xrange=xrange, **kwargs)


def _const_factor(value, column, pe_samples, tag):
    """A constant factor of the weight as a lazy static term."""

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_STATIC, [cols[0]], cst=[float(np.log(value))], feature=S.FEAT_CONST, name=f"const[{tag}]")], [], []

    return LazyWeight([_LazyTerm(("const", tag, float(value), _uid_of(column)), [column], [], build)], pe_samples)


class BSplineSymmetricChiEffective(Base1DBSplineModel):
    """single.py:233-284: a spline in |chi_eff| on [0, 1], times 1/2."""

    def __init__(self, n_splines, chieff, chieff_inj, basis=BSpline, **kwargs):
        xrange = kwargs.pop("xrange", (0.0, 1.0))
        if is_device_array(chieff) or is_device_array(chieff_inj):
            raise NotImplementedError("device-resident arrays: pass |chi_eff| to BSplineChiEffective(xrange=(0, 1)) yourself (the abs() of single.py:262-263 is a host operation here)")
        super().__init__(n_splines, np.abs(chieff), np.abs(chieff_inj), basis=basis, xrange=xrange, **kwargs)

    def __call__(self, coefs, pe_samples=True):
        return self._lazy(coefs, pe_samples) * _const_factor(0.5, self.column, pe_samples, "symmetric")


class BSplineChiPrecess(Base1DBSplineModel):
    """single.py:287-318."""

    def __init__(self, n_splines, chip, chip_inj, basis=BSpline, **kwargs):
        xrange = kwargs.pop("xrange", (0.0, 1.0))
        super().__init__(n_splines, chip, chip_inj, basis=basis, xrange=xrange, **kwargs)


class BSplineRedshift(Base1DBSplineModel):
    """single.py:398-492: ``R(z) = exp(f(z)) dVc/dz / (1+z) / Z(c)``, ``Z`` by trapezoid over 1000
    points between the data's redshift extremes, ``f`` = the projection of the LogXBSpline basis.

    The reference builds that basis with its DEFAULT ``normalize=True``, so ``f(z) = B(log z).c /
    trapezoid(B.c)`` (single.py:77-92 -> interpolation.py:280-317): the exponent is the spline of the
    RESCALED coefficients ``c' = c / (a.c)``, ``a_k = trapezoid(B_k)`` over the basis' own 1000-point
    grid.  The kernel evaluates the plain spline of ``c'``; the rescaling and its chain rule are an
    O(P) host-side :class:`ParamMap`.  With ``normalize=False`` the map is the identity."""

    def __init__(self, n_splines, z, z_inj, dVdc, dVdc_inj, zmax=2.3, basis=LogXBSpline, z_range=None, **kwargs):
        if basis is not LogXBSpline:
            raise NotImplementedError("only the LogXBSpline basis is implemented for BSplineRedshift")
        xrange = kwargs.pop("xrange", (1e-4, zmax))
        super().__init__(n_splines, z, z_inj, xrange=xrange, basis=basis, **kwargs)
        if self.knots is not None:
            raise NotImplementedError("BSplineRedshift with an explicit knot vector / degree != 3 is not implemented (its host-side normaliser map assumes the default knots)")
        self._coef_map = None
        if self.normalize:
            from .lowering import host_spline_design  # tiny, grid-only helper

            # BSpline.norm (interpolation.py:280-291): 1 / trapezoid(grid_bases . c, grid)
            a = np.exp(self.grid_log_w) @ host_spline_design(self.grid_xi, self.xi_range, self.n_splines)

            def fwd(c, a=a):
                return c / (a @ c)

            def vjp(c, g, a=a):
                n = a @ c
                return g / n - a * ((g @ c) / (n * n))

            self._coef_map = ParamMap(fwd, vjp)
        self.dvdc_column = _Column(dVdc, dVdc_inj)
        if z_range is None:
            if is_device_array(self.column.pe) or is_device_array(self.column.inj):
                raise ValueError("device-resident redshift arrays: pass z_range=(zmin, zmax)")
            self.zmin = float(max(np.min(self.column.pe), np.min(self.column.inj)))  # single.py:445
            self.zmax = float(min(np.max(self.column.pe), np.max(self.column.inj)))  # :446
        else:  # a process holding a SHARD passes the whole catalog's range
            self.zmin, self.zmax = float(z_range[0]), float(z_range[1])
        self.zgrid = np.linspace(self.zmin, self.zmax, 1000)  # :447
        with np.errstate(divide="ignore"):
            self._grid_log_w = np.log(S.trapezoid_weights(self.zgrid) * Planck15.dVcdz(self.zgrid) / (1.0 + self.zgrid))  # :465-468
        gxi = np.log(self.zgrid)
        inside = (gxi >= self.xi_range[0]) & (gxi <= self.xi_range[1])
        self._grid_xi = np.where(inside, np.clip(gxi, *self.xi_range), np.nan)  # LogX bases are 0 outside (:175)

    def __call__(self, coefs, pe_samples=True):
        model = self

        def build(slots, group_base, cols):
            rescaled = model._coef_map is not None
            t = S.Term(
                S.TERM_SPLINE, [cols[0]], slots=[slots[0]], n_splines=model.n_splines, logx=True,
                outside=S.OUTSIDE_ZERO,  # funcs() is 0 outside the mask => exp(0) = 1 (single.py:90-92, 488)
                xrange=(model.xmin, model.xmax), norm_group=-1 if rescaled else group_base,
                grid_xi=None if rescaled else model._grid_xi, name="spline[log z]",
            )
            t.xi_range = model.xi_range
            terms = [
                t,
                S.Term(S.TERM_STATIC, [cols[1]], feature=S.FEAT_LOG, name="dVc/dz (given)"),
                S.Term(S.TERM_STATIC, [cols[0]], feature=S.FEAT_NEG_LOG1P, name="1/(1+z)"),
            ]
            return terms, ([] if rescaled else [S.NormGroup(model._grid_log_w, name="Z[redshift]")]), []

        if np.shape(coefs) != (self.n_splines,):
            raise ValueError(f"expected {self.n_splines} coefficients, got shape {np.shape(coefs)}")
        if self._coef_map is None:
            return LazyWeight([_LazyTerm((_uid_of(self), "bsz"), [self.column, self.dvdc_column], [coefs], build)], pe_samples)
        # normalised basis: the exponent uses c' = c / (a.c) (device, through the ParamMap) while the
        # reference's normaliser uses the raw c (single.py:465-468): a per-sample constant, kept on the host
        return LazyWeight(
            [_LazyTerm((_uid_of(self), "bsz-rescaled"), [self.column, self.dvdc_column], [coefs], build, maps={0: self._coef_map},
                       host_norm=self._host_norm)],
            pe_samples,
        )

    def _host_norm(self, params):
        logZ, dlogZ = self._log_normalization(params[0])
        return logZ, [dlogZ]

    def _log_normalization(self, cs):
        """``(log Z(cs), dlog Z/dcs)`` of single.py:453-469 on the host (1000-point sum)."""
        from .lowering import host_spline_design

        D = host_spline_design(self._grid_xi, self.xi_range, self.n_splines)
        li = self._grid_log_w + D @ np.asarray(cs, dtype=np.float64)
        m = np.max(li)
        e = np.exp(li - m)
        return float(m + np.log(np.sum(e))), (e / np.sum(e)) @ D

    def normalization(self, cs):
        """Host-side ``Z(cs)`` (single.py:453-469), e.g. for the merger rate.  As in the reference
        this is the trapezoid of ``dVc/dz / (1+z) exp(B.cs)`` with the UN-rescaled coefficients."""
        from .lowering import host_spline_on_grid

        s = host_spline_on_grid(self._grid_xi, self.xi_range, self.n_splines, np.asarray(cs, dtype=np.float64))
        return float(np.sum(np.exp(self._grid_log_w + s)))


# ================================================================================================
# separable products (gwinferno/models/bsplines/separable.py)
# ================================================================================================
class _IIDPair:
    _cls = None

    def __init__(self, n_splines, x1, x2, x1_inj, x2_inj, **kwargs):
        self.primary_model = self._cls(n_splines, x1, x1_inj, **kwargs)
        self.secondary_model = self._cls(n_splines, x2, x2_inj, **kwargs)

    def __call__(self, coefs, pe_samples=True):
        # same coefficient vector on both columns (separable.py:77-79, 216-218)
        return self.primary_model(coefs, pe_samples=pe_samples) * self.secondary_model(coefs, pe_samples=pe_samples)


class _IndependentPair:
    _cls = None

    def __init__(self, n_splines1, n_splines2, x1, x2, x1_inj, x2_inj, kwargs1={}, kwargs2={}, **kwargs):
        self.primary_model = self._cls(n_splines1, x1, x1_inj, **kwargs1, **kwargs)
        self.secondary_model = self._cls(n_splines2, x2, x2_inj, **kwargs2, **kwargs)

    def __call__(self, pcoefs, scoefs, pe_samples=True):
        return self.primary_model(pcoefs, pe_samples=pe_samples) * self.secondary_model(scoefs, pe_samples=pe_samples)


class BSplineIIDSpinMagnitudes(_IIDPair):
    _cls = BSplineSpinMagnitude


class BSplineIndependentSpinMagnitudes(_IndependentPair):
    _cls = BSplineSpinMagnitude


class BSplineIIDSpinTilts(_IIDPair):
    _cls = BSplineSpinTilt


class BSplineIndependentSpinTilts(_IndependentPair):
    _cls = BSplineSpinTilt


class BSplineEffectiveSpinDims:
    """separable.py:706-778."""

    def __init__(self, n_splines_e, n_splines_p, chieff, chip, chieff_inj, chip_inj, kwargs_e={}, kwargs_p={}, **kwargs):
        self.chi_eff_model = BSplineChiEffective(n_splines_e, chieff, chieff_inj, **kwargs_e, **kwargs)
        self.chi_p_model = BSplineChiPrecess(n_splines_p, chip, chip_inj, **kwargs_p, **kwargs)

    def __call__(self, ecoefs, pcoefs, pe_samples=True):
        return self.chi_eff_model(ecoefs, pe_samples=pe_samples) * self.chi_p_model(pcoefs, pe_samples=pe_samples)


class BSplinePrimaryBSplineRatio:
    """separable.py:446-530."""

    def __init__(self, n_splines_m, n_splines_q, m1, m1_inj, q, q_inj, mmax=100.0, m1min=3.0, m2min=3.0, kwargs_m={}, kwargs_q={}, **kwargs):
        self.primary_model = BSplineMass(n_splines_m, m1, m1_inj, mmin=m1min, mmax=mmax, **kwargs_m, **kwargs)
        self.ratio_model = BSplineRatio(n_splines_q, q, q_inj, qmin=m2min / mmax, **kwargs_q, **kwargs)

    def __call__(self, mcoefs, qcoefs, pe_samples=True):
        return self.ratio_model(qcoefs, pe_samples=pe_samples) * self.primary_model(mcoefs, pe_samples=pe_samples)


def _pairing_term(m1col, m2col, beta, pe_samples, cut01):
    """``(m2/m1)^beta`` (separable.py:608-613, 703); the IID form zeroes q outside [0, 1]."""

    def build(slots, group_base, cols):
        t = S.Term(S.TERM_LINEAR, [cols[0], cols[1]], slots=[slots[0]], cst=[0.0], feature=S.FEAT_LOG_RATIO, name="q^beta")
        cuts = [S.Cut(S.CUT_RATIO_RANGE, [cols[0], cols[1]], 0.0, 1.0)] if cut01 else []
        return [t], [], cuts

    return LazyWeight([_LazyTerm(("pairing", _uid_of(m1col), _uid_of(m2col), cut01), [m2col, m1col], [beta], build)], pe_samples)


class BSplineIIDComponentMasses:
    """separable.py:533-613."""

    def __init__(self, n_splines, m1, m2, m1_inj, m2_inj, mmin=2, mmax=100, **kwargs):
        self.primary_model = BSplineMass(n_splines, m1, m1_inj, mmin=mmin, mmax=mmax, **kwargs)
        self.secondary_model = BSplineMass(n_splines, m2, m2_inj, mmin=mmin, mmax=mmax, **kwargs)

    def __call__(self, coefs, beta=0, pe_samples=True):
        w = self.primary_model(coefs, pe_samples=pe_samples) * self.secondary_model(coefs, pe_samples=pe_samples)
        return w * _pairing_term(self.primary_model.column, self.secondary_model.column, beta, pe_samples, cut01=True)


class BSplineIndependentComponentMasses:
    """separable.py:616-703."""

    def __init__(self, n_splines1, n_splines2, m1, m2, m1_inj, m2_inj, mmin1=2, mmax1=100, mmin2=2, mmax2=100, kwargs1={}, kwargs2={}, **kwargs):
        self.primary_model = BSplineMass(n_splines1, m1, m1_inj, mmin=mmin1, mmax=mmax1, **kwargs1, **kwargs)
        self.secondary_model = BSplineMass(n_splines2, m2, m2_inj, mmin=mmin2, mmax=mmax2, **kwargs2, **kwargs)

    def __call__(self, pcoefs, scoefs, beta=0, pe_samples=True):
        w = self.primary_model(pcoefs, pe_samples=pe_samples) * self.secondary_model(scoefs, pe_samples=pe_samples)
        return w * _pairing_term(self.primary_model.column, self.secondary_model.column, beta, pe_samples, cut01=False)


# ================================================================================================
# parametric densities (gwinferno/models/parametric/parametric.py, gwinferno/distributions.py)
# Free functions of sample arrays, as in the reference; the sample set is inferred from the
# dimensionality of the arrays (2-D = PE samples, 1-D = injections).
# ================================================================================================
_COLUMN_CACHE = collections.OrderedDict()  # uid of the caller's array -> _Column, LRU
_COLUMN_CACHE_MAX = 64


def _col_of(arr):
    """One _Column per distinct sample array (keyed by its uid so the static plan can be cached)."""
    key = _uid_of(arr)
    col = _COLUMN_CACHE.get(key)
    if col is not None:
        _COLUMN_CACHE.move_to_end(key)
        return col
    a = _as_samples(arr)
    col = _Column(a if _is_pe(a) else None, None if _is_pe(a) else a)
    _COLUMN_CACHE[key] = col
    while len(_COLUMN_CACHE) > _COLUMN_CACHE_MAX:
        _COLUMN_CACHE.popitem(last=False)
    return col


def _powerlaw_term(x, alpha, lo, hi, tag):
    col = _col_of(x)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_POWERLAW, [cols[0]], slots=[slots[0]], cst=[float(lo), float(hi)], name=f"powerlaw[{cols[0]}]")], [], []

    return LazyWeight([_LazyTerm(("pl", _uid_of(x), float(lo), float(hi), tag), [col], [alpha], build)], _is_pe(x))


def _powerlaw_ratio_term(q, m1, beta, mmin):
    cq, cm = _col_of(q), _col_of(m1)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_POWERLAW_RATIO, [cols[0], cols[1]], slots=[slots[0]], cst=[float(mmin)], name="powerlaw[q|m1]")], [], []

    return LazyWeight([_LazyTerm(("plq", _uid_of(q), _uid_of(m1), float(mmin)), [cq, cm], [beta], build)], _is_pe(q))


def powerlaw_primary_ratio_pdf(m1, q, alpha, beta, mmin, mmax):
    """parametric.py:27-30."""
    return _powerlaw_ratio_term(q, m1, beta, mmin) * _powerlaw_term(m1, alpha, mmin, mmax, "m1")


def plpeak_primary_pdf(m1, alpha, mmin, mmax, mpp, sigpp, lam, delta=None, _delta_key=None):
    """parametric.py:49-53; with ``delta`` the power-law part carries the low-mass window
    ``smooth(delta, m1, mmin)`` exactly as the reference evaluates it (distributions.py:16-21)."""
    col = _col_of(m1)
    tapered = delta is not None

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_PLPEAK, [cols[0]], slots=list(slots[: 5 if tapered else 4]), cst=[float(mmin), float(mmax)], name="plpeak[m1]")], [], []

    params = [alpha, mpp, sigpp, lam] + ([delta] if tapered else [])
    keys = [None] * 4 + ([_delta_key] if tapered else [])
    return LazyWeight([_LazyTerm(("plpeak", _uid_of(m1), float(mmin), float(mmax), tapered), [col], params, build, param_keys=keys)], _is_pe(m1))


def _smooth_term(delta, x, x2, xmin, _key=None):
    """``smooth(delta, x [* x2], xmin)`` (distributions.py:16-21) as a lazy factor."""
    cx = [_col_of(x)] + ([_col_of(x2)] if x2 is not None else [])

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_SMOOTH, list(cols), slots=[slots[0]], cst=[float(xmin)], name="smooth")], [], []

    return LazyWeight([_LazyTerm(("smooth", _uid_of(x), _uid_of(x2), float(xmin)), cx, [delta], build, param_keys=[_key])], _is_pe(x))


def plpeak_primary_ratio_pdf(m1, q, alpha, beta, mmin, mmax, mpp, sigpp, lam, delta=None):
    """parametric.py:39-46."""
    dk = _share() if delta is not None else None  # ONE delta for both windows
    w = _powerlaw_ratio_term(q, m1, beta, mmin) * plpeak_primary_pdf(m1, alpha, mmin, mmax, mpp, sigpp, lam, delta=delta, _delta_key=dk)
    if delta is not None:
        w = w * _smooth_term(delta, q, m1, mmin, _key=dk)  # smooth(delta, q * m1, mmin)   (:46)
    return w


def beta_spin_magnitude(a, alpha, beta, amax=1, _keys=None):
    """parametric.py:63-64 -> distributions.py:146-162."""
    col = _col_of(a)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_BETA, [cols[0]], slots=list(slots[:2]), cst=[float(amax)], name=f"beta[{cols[0]}]")], [], []

    return LazyWeight([_LazyTerm(("beta", _uid_of(a), float(amax)), [col], [alpha, beta], build, param_keys=_keys)], _is_pe(a))


def iid_spin_magnitude(a1, a2, alpha_mag, beta_mag, amax=1):
    keys = [_share(), _share()]  # identically distributed: both components read the same two slots
    return beta_spin_magnitude(a1, alpha_mag, beta_mag, amax, _keys=keys) * beta_spin_magnitude(a2, alpha_mag, beta_mag, amax, _keys=keys)


def independent_spin_magnitude_beta_dist(a1, a2, alpha_mag1, beta_mag1, alpha_mag2, beta_mag2, amax1=1, amax2=1):
    return beta_spin_magnitude(a1, alpha_mag1, beta_mag1, amax1) * beta_spin_magnitude(a2, alpha_mag2, beta_mag2, amax2)


def mixture_isoalign_spin_tilt(ct, xi_tilt, sigma_tilt, _keys=None):
    """parametric.py:84-86."""
    col = _col_of(ct)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_ISOALIGN, [cols[0]], slots=list(slots[:2]), name=f"isoalign[{cols[0]}]")], [], []

    return LazyWeight([_LazyTerm(("isoalign", _uid_of(ct)), [col], [xi_tilt, sigma_tilt], build, param_keys=_keys)], _is_pe(ct))


def default_spin_tilt(ct1, ct2, xi_tilt, sigma_tilt):
    """parametric.py:97-102: ``(1 - xi)/4 + xi TN(ct1) TN(ct2)`` (both tilts aligned together)."""
    c1, c2 = _col_of(ct1), _col_of(ct2)

    def build(slots, group_base, cols):
        return [S.Term(S.TERM_ISOALIGN_PAIR, [cols[0], cols[1]], slots=list(slots[:2]), name=f"isoalign2[{cols[0]},{cols[1]}]")], [], []

    return LazyWeight([_LazyTerm(("isoalign2", _uid_of(ct1), _uid_of(ct2)), [c1, c2], [xi_tilt, sigma_tilt], build)], _is_pe(ct1))


def iid_spin_tilt(ct1, ct2, xi_tilt, sigma_tilt):
    keys = [_share(), _share()]
    return mixture_isoalign_spin_tilt(ct1, xi_tilt, sigma_tilt, _keys=keys) * mixture_isoalign_spin_tilt(ct2, xi_tilt, sigma_tilt, _keys=keys)


def independent_spin_tilt(ct1, ct2, xi_tilt_1, xi_tilt_2, sigma_tilt1, sigma_tilt2):
    return mixture_isoalign_spin_tilt(ct1, xi_tilt_1, sigma_tilt1) * mixture_isoalign_spin_tilt(ct2, xi_tilt_2, sigma_tilt2)


class BSplinePrimaryPowerlawRatio:
    """separable.py:295-365."""

    def __init__(self, n_splines, m1, m1_inj, mmin=2, mmax=100, **kwargs):
        self.primary_model = BSplineMass(n_splines, m1, m1_inj, mmin=mmin, mmax=mmax, **kwargs)

    def __call__(self, m1, q, beta, mmin, coefs, pe_samples=True):
        return self.primary_model(coefs, pe_samples=pe_samples) * _powerlaw_ratio_term(q, m1, beta, mmin)


class PLPeakPrimaryBSplineRatio:
    """separable.py:368-443."""

    def __init__(self, n_splines, q, q_inj, **kwargs):
        self.ratio_model = BSplineRatio(n_splines, q, q_inj, **kwargs)

    def __call__(self, m1, alpha, mmin, mmax, peak_mean, peak_sd, peak_frac, coefs, pe_samples=True):
        return plpeak_primary_pdf(m1, alpha, mmin, mmax, peak_mean, peak_sd, peak_frac) * self.ratio_model(coefs, pe_samples=pe_samples)


# ================================================================================================
# redshift models (parametric.py:112-145, spline_perturbation.py:304-372)
# ================================================================================================
class PowerlawRedshiftModel:
    def __init__(self, z_pe, z_inj, z_range=None):
        """``z_range=(zmin, zmax)`` (extension) overrides the data-derived range: a process that holds
        only a SHARD of the catalog must be given the range of the whole catalog."""
        self.column = _Column(z_pe, z_inj)
        z_pe = self.column.pe
        z_inj = self.column.inj
        if z_range is None:
            if is_device_array(z_pe) or is_device_array(z_inj):
                raise ValueError("device-resident redshift arrays: pass z_range=(zmin, zmax) (the data-derived range of parametric.py:114-115 needs the values on the host)")
            self.zmin = float(max(np.min(z_pe), np.min(z_inj)))  # parametric.py:114
            self.zmax = float(min(np.max(z_pe), np.max(z_inj)))  # parametric.py:115
        else:
            self.zmin, self.zmax = float(z_range[0]), float(z_range[1])
        self.zs = np.linspace(self.zmin, self.zmax, 1000)  # :116
        self.dVdz_ = Planck15.dVcdz(self.zs)  # :117
        with np.errstate(divide="ignore"):
            self._grid_log_w = np.log(S.trapezoid_weights(self.zs) * self.dVdz_)
        self._grid_log1pz = np.log(1.0 + self.zs)
        self._norm_cache = None

    def _which(self, z):
        if z is self.column.pe or (_is_pe(z) and tuple(z.shape) == tuple(self.column.pe.shape)):
            return True
        return False

    def _terms(self, slots, group_base, cols, spline=None):
        zc = cols[0]
        terms = [
            S.Term(S.TERM_STATIC, [zc], feature=S.FEAT_LOG_DVDZ, name="dVc/dz"),
            S.Term(
                S.TERM_LINEAR, [zc], slots=[slots[0]], cst=[-1.0], feature=S.FEAT_LOG1P, norm_group=group_base, grid_feat=self._grid_log1pz, name="(1+z)^(lamb-1)"
            ),
        ]
        groups = [S.NormGroup(self._grid_log_w, name="Z[redshift]")]
        # where(z <= zmax, ., 0)   (parametric.py:141-145, spline_perturbation.py:368-372)
        cuts = [S.Cut(S.CUT_RANGE, [zc], -np.inf, self.zmax)]
        return terms, groups, cuts

    def __call__(self, z, lamb):
        pe_samples = self._which(z)
        model = self

        def build(slots, group_base, cols):
            return model._terms(slots, group_base, cols)

        return LazyWeight([_LazyTerm((_uid_of(self), "plz"), [self.column], [lamb], build)], pe_samples)

    def log_prob(self, z, lamb):
        """parametric.py:129-136: the same density as ``__call__`` in the log domain (combine with
        ``+`` and ``- jnp.log(prior)``, evaluate with ``hierarchical_likelihood(..., log=True)``)."""
        w = self(z, lamb)
        return LazyWeight(w.terms, w.pe_samples, log_domain=True)

    # host-side value of the normaliser (a 1000-point sum; used for ``surveyed_hypervolume``)
    def normalization(self, lamb):
        lamb = float(np.asarray(lamb))
        return float(np.sum(np.exp(self._grid_log_w + (lamb - 1.0) * self._grid_log1pz)))


class PowerlawSplineRedshiftModel(PowerlawRedshiftModel):
    def __init__(self, n_splines, z_pe, z_inj, basis=LogXBSpline, z_range=None):
        super().__init__(z_pe=z_pe, z_inj=z_inj, z_range=z_range)
        if basis is not LogXBSpline:
            raise NotImplementedError("only the LogXBSpline perturbation basis is implemented")
        self.n_splines = int(n_splines)
        self.xi_range = (float(np.log(self.zmin)), float(np.log(self.zmax)))
        gxi = np.log(self.zs)
        inside = (gxi >= self.xi_range[0]) & (gxi <= self.xi_range[1])
        # LogX bases are ZERO outside the range (interpolation.py:175): mark with NaN
        self._grid_xi = np.where(inside, np.clip(gxi, *self.xi_range), np.nan)

    def __call__(self, z, lamb, cs):
        pe_samples = self._which(z)
        model = self

        def build(slots, group_base, cols):
            terms, groups, cuts = model._terms(slots, group_base, cols)
            t = S.Term(
                S.TERM_SPLINE,
                [cols[0]],
                slots=[slots[1]],
                n_splines=model.n_splines,
                logx=True,
                outside=S.OUTSIDE_ZERO,
                xrange=(model.zmin, model.zmax),
                norm_group=group_base,
                grid_xi=model._grid_xi,
                name="spline[log z]",
            )
            t.xi_range = model.xi_range
            return terms + [t], groups, cuts

        if np.shape(cs) != (self.n_splines,):
            raise ValueError(f"expected {self.n_splines} redshift coefficients, got shape {np.shape(cs)}")
        return LazyWeight([_LazyTerm((_uid_of(self), "plsz"), [self.column], [lamb, cs], build)], pe_samples)

    def normalization(self, lamb, cs):
        """Host-side ``Z(lamb, cs)`` (spline_perturbation.py:323-336): a 1000-point sum used for
        the ``surveyed_hypervolume`` argument; the likelihood path itself gets log Z on the GPU."""
        from .lowering import host_spline_on_grid  # tiny, grid-only helper

        lamb = float(np.asarray(lamb))
        s = host_spline_on_grid(self._grid_xi, self.xi_range, self.n_splines, np.asarray(cs, dtype=np.float64))
        return float(np.sum(np.exp(self._grid_log_w + (lamb - 1.0) * self._grid_log1pz + s)))
