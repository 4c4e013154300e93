"""Lower a pair of :class:`~gwinferno_b200.models.LazyWeight` products (PE samples, found
injections) to the static :class:`~gwinferno_b200.spec.ModelSpec` + column arrays that the C-ABI
consumes, and the per-call flat hyper-parameter vector ``Lambda``.

The static part depends only on WHICH model objects / sample arrays appear and on how parameter
objects are shared between terms (IID models pass one coefficient vector to two columns,
gwinferno/models/bsplines/separable.py:77-79), so it is cached by that structure; every later
call with new hyper-parameter values only rebuilds ``Lambda``.
"""

import numpy as np

from . import spec as S


def host_spline_on_grid(grid_xi, xi_range, n_splines, coefs):
    """Cubic B-spline ``sum_k B_k(xi) c_k`` on the <= 1500-point normalisation grid (host-side
    scalar glue for ``normalization()``; NaN grid entries are outside the basis range => 0)."""
    lo, hi = xi_range
    n_int = n_splines - 2
    dx = (hi - lo) / (n_int - 1)
    ok = np.isfinite(grid_xi)
    t = (np.where(ok, grid_xi, lo) - lo) / dx
    j = np.clip(np.floor(t), 0, n_int - 2).astype(np.int64)
    u = t - j
    w = np.stack([(1 - u) ** 3, 3 * u**3 - 6 * u**2 + 4, -3 * u**3 + 3 * u**2 + 3 * u + 1, u**3], axis=-1) / 6.0
    s = np.einsum("gk,gk->g", w, coefs[j[:, None] + np.arange(4)])
    return np.where(ok, s, 0.0)


def host_spline_design(grid_xi, xi_range, n_splines):
    """Dense ``[G, n_splines]`` design matrix of the cubic B-spline basis on a normalisation grid
    (host-side O(P) glue only: parameter maps and host normalisers; NaN grid entries => zero row)."""
    lo, hi = xi_range
    n_int = n_splines - 2
    dx = (hi - lo) / (n_int - 1)
    ok = np.isfinite(grid_xi)
    t = (np.where(ok, grid_xi, lo) - lo) / dx
    j = np.clip(np.floor(t), 0, n_int - 2).astype(np.int64)
    u = t - j
    w = np.stack([(1 - u) ** 3, 3 * u**3 - 6 * u**2 + 4, -3 * u**3 + 3 * u**2 + 3 * u + 1, u**3], axis=-1) / 6.0
    D = np.zeros((grid_xi.size, n_splines))
    rows = np.arange(grid_xi.size)
    for k in range(4):
        D[rows, j + k] = np.where(ok, w[:, k], 0.0)
    return D


class Lowered:
    """Static lowering result: spec, column arrays, and the recipe to flatten parameters."""

    def __init__(self, spec, pe_cols, inj_cols, param_layout, slot_of):
        self.spec = spec
        self.pe_cols = pe_cols  # name -> (E, S) float64
        self.inj_cols = inj_cols  # name -> (I,) float64
        self.param_layout = param_layout  # list of (slot, size) in order of first appearance
        # id(parameter object of the lowering call) -> Lambda offsets of every argument position it was
        # passed to (arrays shared by IID models: one; a scalar object passed twice: two)
        self.slot_of = slot_of

    def all_slots_for(self, obj):
        """Every Lambda index range a parameter object of the model calls occupies."""
        n = int(np.size(obj))
        return [slice(off, off + n) for off in self.slot_of[id(obj)]]

    def slots_for(self, obj):
        """Lambda index range of a parameter object that was passed to the model calls."""
        sl = self.all_slots_for(obj)
        if len(sl) != 1:
            raise ValueError("this scalar object was passed to several argument positions (each has its own Lambda slot): use all_slots_for() / LikelihoodResult.grad()")
        return sl[0]


_ZEROS = {}  # n_params -> a shared all-zero vector (read-only use: "no host-side normaliser")


def _walk(pe_w):
    """ONE pass over the PE-side lazy terms, cached on the weight object: the parameter-sharing pattern, the distinct
    parameter blocks ``(offset, value, ParamMap or None)`` in Lambda order, and whether any term carries a parameter map or
    a host-side normaliser (hierarchical_likelihood runs on every sampler step: ~60 us of repeated walks otherwise)."""
    w = getattr(pe_w, "_walk_cache", None)
    if w is not None:
        return w
    seen, pattern, blocks, off = {}, [], [], 0
    has_map = has_norm = False
    for t in pe_w.terms:
        maps = getattr(t, "maps", None) or {}
        if getattr(t, "host_norm", None) is not None:
            has_norm = True
        for i, (p, k) in enumerate(zip(t.params, t.param_keys)):
            m = maps.get(i)
            j = seen.get(k)
            if j is None:
                j = seen[k] = len(blocks)
                blocks.append((off, p, m))
                off += p.size
                if m is not None:
                    has_map = True
            elif blocks[j][2] is not m:
                raise ValueError("a parameter object shared between terms must use the same parameter map in all of them")
            pattern.append((j, p.size))
    w = (tuple(pattern), blocks, off, has_map, has_norm)
    try:
        pe_w._walk_cache = w
    except AttributeError:
        pass
    return w


def _structure(pe_w, inj_w):
    """Hashable description of the static structure + the parameter-sharing pattern."""
    if len(pe_w.terms) != len(inj_w.terms):
        raise ValueError("PE and injection weights must be built from the same sequence of model terms")
    keys = tuple(t.key for t in pe_w.terms) + tuple(t.key for t in inj_w.terms)
    return keys, _walk(pe_w)[0]


def object_slot_map(pe_w):
    """``{id(parameter object): [Lambda offsets]}`` of one lowering call (the objects are alive in the
    lazy terms for as long as the weights are)."""
    key_off, obj_offs, off = {}, {}, 0
    for t in pe_w.terms:
        for p, k, o in zip(t.params, t.param_keys, t._orig):
            if k not in key_off:
                key_off[k] = off
                off += p.size
            lst = obj_offs.setdefault(id(o), [])
            if key_off[k] not in lst:
                lst.append(key_off[k])
    return obj_offs


def _param_blocks(pe_w):
    """``[(offset, value, ParamMap or None)]`` per distinct parameter object, in Lambda order."""
    w = _walk(pe_w)
    return w[1], w[2]


def flatten_params(pe_w, n_params, param_layout=None):
    """Flat ``Lambda`` from the parameter values carried by the PE-side lazy terms (after their
    host-side parameter maps, if any)."""
    blocks, off = _param_blocks(pe_w)
    if off != n_params:
        raise ValueError("parameter structure changed between calls")
    Lam = np.zeros(n_params)
    for o, p, m in blocks:
        Lam[o : o + p.size] = p if m is None else m.fwd(p)
    return Lam


def pull_back(pe_w, grad):
    """Gradient (or Jacobian rows, last axis = Lambda) with respect to the kernel's Lambda -> with
    respect to the parameter values the caller passed (chain rule through the parameter maps)."""
    w = _walk(pe_w)
    blocks = w[1]
    if not w[3]:
        return grad
    out = np.array(grad, dtype=np.float64, copy=True)
    flat = out.reshape(-1, out.shape[-1])
    for o, p, m in blocks:
        if m is not None:
            for r in range(flat.shape[0]):
                flat[r, o : o + p.size] = m.vjp(p, flat[r, o : o + p.size])
    return out


def host_log_norm(pe_w, n_params):
    """``(log Z, dlog Z/dLambda[P])`` of the host-side normalisers: every sample's log-weight is
    lower by ``log Z`` than what the device model evaluates (the derivative is with respect to the
    parameter values the caller passed)."""
    if not _walk(pe_w)[4]:
        return 0.0, _ZEROS.get(n_params) if n_params in _ZEROS else _ZEROS.setdefault(n_params, np.zeros(n_params))
    slot, off = {}, 0
    for t in pe_w.terms:
        for p, k in zip(t.params, t.param_keys):
            if k not in slot:
                slot[k] = off
                off += p.size
    logZ, dlogZ = 0.0, np.zeros(n_params)
    for t in pe_w.terms:
        if getattr(t, "host_norm", None) is None:
            continue
        z, dz = t.host_norm(t.params)
        logZ += z
        for p, k, d in zip(t.params, t.param_keys, dz):
            dlogZ[slot[k] : slot[k] + p.size] += d
    return logZ, dlogZ


def lower(pe_w, inj_w):
    if pe_w.pe_samples is False or inj_w.pe_samples is True:
        raise ValueError("first argument must be the PE-sample weights, second the injection weights")
    if len(pe_w.terms) != len(inj_w.terms):
        raise ValueError("PE and injection weights must be built from the same sequence of model terms")
    # ---- parameter slots (shared objects share slots) -----------------------------------------
    slot_of = {}
    layout = []
    off = 0
    for t in pe_w.terms:
        for p, k in zip(t.params, t.param_keys):
            if k not in slot_of:
                slot_of[k] = off
                layout.append((off, p.size))
                off += p.size
    n_params = off
    # ---- columns ---------------------------------------------------------------------------------
    col_names = {}
    pe_cols, inj_cols = {}, {}

    def col_name(cpe, cinj):
        pe_arr = cpe.pe if cpe.pe is not None else cinj.pe
        inj_arr = cinj.inj if cinj.inj is not None else cpe.inj
        if pe_arr is None or inj_arr is None:
            raise ValueError("a term is missing its PE or injection sample array")
        key = (id(pe_arr), id(inj_arr))
        if key not in col_names:
            name = f"c{len(col_names)}"
            col_names[key] = name
            if len(pe_arr.shape) != 2 or len(inj_arr.shape) != 1:
                raise ValueError("PE columns must be (E, S) arrays and injection columns (I,) arrays")
            pe_cols[name] = pe_arr
            inj_cols[name] = inj_arr
        return col_names[key]

    terms, groups, cuts = [], [], []
    for tp, ti in zip(pe_w.terms, inj_w.terms):
        if len(tp.columns) != len(ti.columns):
            raise ValueError("PE / injection term mismatch")
        names = [col_name(a, b) for a, b in zip(tp.columns, ti.columns)]
        slots = [slot_of[k] for k in tp.param_keys]
        tt, gg, cc = tp.build(slots, len(groups), names)
        terms += tt
        groups += gg
        for c in cc:
            if not any(c.kind == o.kind and c.cols == o.cols and c.lo == o.lo and c.hi == o.hi for o in cuts):
                cuts.append(c)
    spec = S.ModelSpec(terms, groups, cuts, n_params)
    shapes = {tuple(v.shape) for v in pe_cols.values()}
    if len(shapes) != 1:
        raise ValueError(f"PE columns have inconsistent shapes: {shapes}")
    shapes = {tuple(v.shape) for v in inj_cols.values()}
    if len(shapes) != 1:
        raise ValueError(f"injection columns have inconsistent shapes: {shapes}")
    return Lowered(spec, pe_cols, inj_cols, layout, object_slot_map(pe_w))
