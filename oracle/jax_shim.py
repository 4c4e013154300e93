"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (gwinferno_b200/).

NumPy-backed stand-in for the tiny part of ``jax`` / ``numpyro`` that the reference's hot-path
modules touch, so that the reference's OWN source files under ``/root/reference/gwinferno``
can be executed unmodified in fp64 in a container that has no jax/jaxlib/numpyro.

It is used by exactly two things:
  * ``tests/golden/make_golden.py`` (run in the build container, where /root/reference exists)
    to generate the committed golden vectors, and
  * ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on the
    GPU box) to pin ``oracle/popmodel.py`` against the reference's code.

API surface covered (from grep over the reference's hot-path files):
  jax.numpy (NumPy semantics + ``x.at[idx].set/get``), jax.scipy.integrate.trapezoid,
  jax.scipy.special.{erf,betaln,logsumexp}, jax.lax.fori_loop, jax.jit, jax.vmap,
  jax.tree_util.register_pytree_node_class, jax.random (empty), jax.Array (dummy);
  numpyro.* -> MagicMock.
"""

import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import scipy.integrate
import scipy.special

REFERENCE_ROOT = os.environ.get("GWI_REFERENCE_ROOT", "/root/reference")


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr = arr
        self._idx = idx

    def set(self, v):
        v = np.asarray(v)
        out = np.array(self._arr, dtype=np.result_type(self._arr, v), copy=True).view(ShimArray)
        out[self._idx] = v
        return out

    def get(self):
        return self._arr[self._idx]

    def add(self, v):
        v = np.asarray(v)
        out = np.array(self._arr, dtype=np.result_type(self._arr, v), copy=True).view(ShimArray)
        np.add.at(out, self._idx, v)
        return out


class ShimArray(np.ndarray):
    """ndarray with jax's functional ``.at[idx].set(v)`` update syntax."""

    @property
    def at(self):
        return _At(self)

    # jax arrays are immutable: augmented assignment REBINDS (and may promote the dtype, which
    # complex-step differentiation of the reference code relies on)
    def __imul__(self, o):
        return self * o

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __itruediv__(self, o):
        return self / o


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, ShimArray):
        return x.view(ShimArray)
    return x


def _wrapping(fn):
    def inner(*a, **k):
        out = fn(*a, **k)
        if isinstance(out, (list, tuple)):
            return type(out)(_wrap(o) for o in out)
        return _wrap(out)

    inner.__name__ = getattr(fn, "__name__", "fn")
    return inner


def _exp(x, *a, **k):
    """``exp`` that stays usable for complex-step differentiation: where a COMPLEX argument overflows
    (Re > 709.78) NumPy returns ``inf + nan j`` and everything downstream becomes NaN, although the real
    function is +inf there with every derivative lost to the same overflow.  Return ``inf + 0j`` at
    those points (real arguments are passed through untouched)."""
    if np.iscomplexobj(x):
        x = np.asarray(x)
        over = np.real(x) > 709.78
        if np.any(over):
            out = np.exp(np.where(over, 0.0, x), *a, **k)
            return np.where(over, complex(np.inf, 0.0), out)
    return np.exp(x, *a, **k)


def _build_jnp():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if name.startswith("_"):
            continue
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            setattr(jnp, name, _wrapping(obj))
        else:
            setattr(jnp, name, obj)
    jnp.exp = _wrapping(_exp)
    jnp.ndarray = np.ndarray
    jnp.linalg = np.linalg
    jnp.inf = np.inf
    jnp.pi = np.pi
    jnp.float64 = np.float64
    return jnp


def _jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def _fori_loop(lo, hi, body, init):
    val = init
    for i in range(lo, hi):
        val = body(i, val)
    return val


def _vmap(fn, in_axes=0, out_axes=0):
    def inner(*args):
        n = len(args[0])
        return _wrap(np.stack([fn(*[a[i] for a in args]) for i in range(n)]))

    return inner


def _betaln(a, b):
    # loggamma accepts complex arguments (complex-step differentiation of the reference code)
    lg = scipy.special.loggamma
    out = lg(a) + lg(b) - lg(a + b)
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        return out
    return np.real(out)


def install():
    """Register the shim modules in ``sys.modules`` (idempotent)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "_gwi_shim", False):
        return sys.modules["jax"]
    jax = types.ModuleType("jax")
    jax._gwi_shim = True
    jnp = _build_jnp()
    jax.numpy = jnp
    jax.jit = _jit
    jax.vmap = _vmap

    class Array:  # SciPy's array-API helper probes sys.modules["jax"].Array
        pass

    jax.Array = Array
    lax = types.ModuleType("jax.lax")
    lax.fori_loop = _fori_loop
    jax.lax = lax
    jscipy = types.ModuleType("jax.scipy")
    jint = types.ModuleType("jax.scipy.integrate")
    jint.trapezoid = _wrapping(scipy.integrate.trapezoid)
    jspec = types.ModuleType("jax.scipy.special")
    jspec.erf = scipy.special.erf
    jspec.betaln = _betaln
    jspec.logsumexp = scipy.special.logsumexp
    jspec.gammaln = scipy.special.gammaln
    jscipy.integrate = jint
    jscipy.special = jspec
    jax.scipy = jscipy
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.register_pytree_node_class = lambda c: c
    jax.tree_util = tree_util
    jax.random = types.ModuleType("jax.random")
    mods = {
        "jax": jax,
        "jax.numpy": jnp,
        "jax.lax": lax,
        "jax.scipy": jscipy,
        "jax.scipy.integrate": jint,
        "jax.scipy.special": jspec,
        "jax.tree_util": tree_util,
        "jax.random": jax.random,
    }
    sys.modules.update(mods)
    for name in (
        "numpyro",
        "numpyro.distributions",
        "numpyro.infer",
        "numpyro.optim",
        "numpyro.distributions.util",
        "numpyro.distributions.constraints",
    ):
        sys.modules.setdefault(name, mock.MagicMock())
    return jax


_LOADED = {}


def load_reference(root=None):
    """Load the reference's hot-path modules by file path (bypassing gwinferno/__init__.py,
    which imports h5py/arviz/...).  Returns a dict of module objects keyed by short name."""
    root = root or REFERENCE_ROOT
    if root in _LOADED:
        return _LOADED[root]
    if not os.path.isdir(os.path.join(root, "gwinferno")):
        raise FileNotFoundError(f"reference tree not found at {root}")
    install()
    pkg_root = os.path.join(root, "gwinferno")
    for pkg, sub in (
        ("gwinferno", ""),
        ("gwinferno.models", "models"),
        ("gwinferno.models.bsplines", "models/bsplines"),
        ("gwinferno.models.parametric", "models/parametric"),
        ("gwinferno.pipeline", "pipeline"),
    ):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(pkg_root, sub)]
        sys.modules[pkg] = m
    order = [
        ("cosmology", "gwinferno.cosmology", "cosmology.py"),
        ("distributions", "gwinferno.distributions", "distributions.py"),
        ("interpolation", "gwinferno.interpolation", "interpolation.py"),
        ("parametric", "gwinferno.models.parametric.parametric", "models/parametric/parametric.py"),
        ("single", "gwinferno.models.bsplines.single", "models/bsplines/single.py"),
        ("separable", "gwinferno.models.bsplines.separable", "models/bsplines/separable.py"),
        ("smoothing", "gwinferno.models.bsplines.smoothing", "models/bsplines/smoothing.py"),
        ("spline_perturbation", "gwinferno.models.spline_perturbation", "models/spline_perturbation.py"),
        ("parser", "gwinferno.pipeline.parser", "pipeline/parser.py"),
        ("analysis", "gwinferno.pipeline.analysis", "pipeline/analysis.py"),
    ]
    out = {}
    for short, modname, rel in order:
        spec = importlib.util.spec_from_file_location(modname, os.path.join(pkg_root, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        out[short] = mod
        parent, _, leaf = modname.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    _LOADED[root] = out
    return out
