"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the plain-C oracle (oracle/c/gwi_oracle.c).

The C oracle takes the C-ABI's own descriptors (include/gwi.h); the structures are restated here so
that nothing from the product package is imported.  The model description is the duck-typed
``ModelSpec`` (only attributes are read), the columns are dictionaries of NumPy arrays."""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgwi_oracle.so")
_dp = C.POINTER(C.c_double)


class gwi_term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("feature", C.c_int32), ("outside", C.c_int32), ("logx", C.c_int32), ("col", C.c_int32 * 2), ("slot", C.c_int32 * 6),
                ("cst", C.c_double * 4), ("n_splines", C.c_int32), ("norm_group", C.c_int32), ("x_lo", C.c_double), ("x_hi", C.c_double),
                ("xi_lo", C.c_double), ("xi_hi", C.c_double), ("grid", _dp), ("knots", _dp), ("n_knots", C.c_int32), ("order", C.c_int32)]


class gwi_norm_group(C.Structure):
    _fields_ = [("n_grid", C.c_int32), ("log_w", _dp)]


class gwi_cut(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32 * 2), ("lo", C.c_double), ("hi", C.c_double)]


class gwi_model_desc(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("terms", C.POINTER(gwi_term)), ("n_groups", C.c_int32), ("groups", C.POINTER(gwi_norm_group)), ("n_cuts", C.c_int32),
                ("cuts", C.POINTER(gwi_cut)), ("n_params", C.c_int32), ("need_neff_grad", C.c_int32), ("chunk_steps", C.c_int32), ("n_deep", C.c_int32),
                ("batch_hint", C.c_int32), ("reserved_", C.c_int32)]


class gwi_catalog_desc(C.Structure):
    _fields_ = [("n_columns", C.c_int32), ("n_events", C.c_int32), ("pe_offsets", C.POINTER(C.c_int64)), ("pe_columns", C.POINTER(_dp)), ("n_inj", C.c_int64),
                ("inj_columns", C.POINTER(_dp)), ("total_inj", C.c_double), ("device", C.c_int32)]


_lib = None


def build():
    """Compile the C oracle with gcc (seconds)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.gwio_evaluate.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_dp)


class _Marshalled:
    """ctypes descriptors + every array they point to (kept alive)."""

    def __init__(self, spec, pe_cols, inj_cols, total_inj):
        self.keep = []
        names = list(pe_cols.keys())
        index = {n: i for i, n in enumerate(names)}
        nt = len(spec.terms)
        self.terms = (gwi_term * max(1, nt))()
        for i, t in enumerate(spec.terms):
            ct = self.terms[i]
            ct.kind, ct.feature, ct.outside, ct.logx = int(t.kind), int(t.feature), int(t.outside), int(bool(t.logx))
            cols = [index[c] for c in t.cols] + [-1, -1]
            ct.col[0], ct.col[1] = cols[0], cols[1]
            for k, s in enumerate((list(t.slots) + [-1] * 6)[:6]):
                ct.slot[k] = int(s)
            for k, v in enumerate((list(t.cst) + [0.0] * 4)[:4]):
                ct.cst[k] = float(v)
            ct.n_splines, ct.norm_group = int(t.n_splines), int(t.norm_group)
            ct.x_lo, ct.x_hi = float(t.xrange[0]), float(t.xrange[1])
            xi = t.xi_range if getattr(t, "xi_range", None) is not None else ((np.log(t.xrange[0]), np.log(t.xrange[1])) if t.logx and t.n_splines else t.xrange)
            ct.xi_lo, ct.xi_hi = float(xi[0]), float(xi[1])
            grid = t.grid_xi if t.grid_xi is not None else t.grid_feat
            if grid is not None and t.norm_group >= 0:
                g = np.ascontiguousarray(grid, dtype=np.float64)
                self.keep.append(g)
                ct.grid = _ptr(g)
            if getattr(t, "knots", None) is not None:
                kn = np.ascontiguousarray(t.knots, dtype=np.float64)
                self.keep.append(kn)
                ct.knots, ct.n_knots, ct.order = _ptr(kn), kn.size, int(t.order)
        ng = len(spec.groups)
        self.groups = (gwi_norm_group * max(1, ng))()
        for i, g in enumerate(spec.groups):
            lw = np.ascontiguousarray(g.log_w, dtype=np.float64)
            self.keep.append(lw)
            self.groups[i].n_grid, self.groups[i].log_w = lw.size, _ptr(lw)
        nc = len(spec.cuts)
        self.cuts = (gwi_cut * max(1, nc))()
        for i, c in enumerate(spec.cuts):
            cols = [index[x] for x in c.cols] + [-1, -1]
            self.cuts[i].kind, self.cuts[i].lo, self.cuts[i].hi = int(c.kind), float(c.lo), float(c.hi)
            self.cuts[i].col[0], self.cuts[i].col[1] = cols[0], cols[1]
        self.model = gwi_model_desc(nt, self.terms, ng, self.groups, nc, self.cuts, int(spec.n_params), 1, 0, -1)
        pe = [np.ascontiguousarray(pe_cols[n], dtype=np.float64) for n in names]
        inj = [np.ascontiguousarray(inj_cols[n], dtype=np.float64) for n in names]
        self.keep += pe + inj
        self.E, S = pe[0].shape
        self.offsets = np.arange(self.E + 1, dtype=np.int64) * S
        self.pe_ptrs = (_dp * len(names))(*[_ptr(a) for a in pe])
        self.inj_ptrs = (_dp * len(names))(*[_ptr(a) for a in inj])
        self.catalog = gwi_catalog_desc(len(names), self.E, self.offsets.ctypes.data_as(C.POINTER(C.c_int64)), self.pe_ptrs, int(inj[0].size), self.inj_ptrs,
                                        float(total_inj), 0)


def evaluate(spec, pe_cols, inj_cols, total_inj, Lam, want_jac=True, n_threads=1, want_neff_jac=True):
    """Same outputs as ``oracle.popmodel.evaluate`` (without ``dlogZ``), computed by the C oracle.
    ``want_neff_jac=False`` skips the N_eff Jacobians (not needed without ``marginalize_selection``)."""
    lib = load()
    m = _Marshalled(spec, pe_cols, inj_cols, total_inj)
    P, E = int(spec.n_params), m.E
    lam = np.ascontiguousarray(Lam, dtype=np.float64)
    out = dict(logBF=np.zeros(E), logNeff=np.zeros(E), log_mu=np.zeros(1), logNeff_inj=np.zeros(1), logZ=np.zeros(max(1, len(spec.groups))))
    jac = {}
    if want_jac:
        jac = dict(J_logBF=np.zeros((E, P)), J_log_mu=np.zeros(P))
        if want_neff_jac:
            jac.update(J_logNeff=np.zeros((E, P)), J_logNeff_inj=np.zeros(P))
    null = C.cast(None, _dp)
    jp = [_ptr(jac[k]) if k in jac else null for k in ("J_logBF", "J_logNeff", "J_log_mu", "J_logNeff_inj")]
    rc = lib.gwio_evaluate(C.byref(m.catalog), C.byref(m.model), _ptr(lam), int(n_threads), _ptr(out["logBF"]), _ptr(out["logNeff"]), _ptr(out["log_mu"]),
                           _ptr(out["logNeff_inj"]), *jp, _ptr(out["logZ"]))
    if rc != 0:
        raise RuntimeError("gwio_evaluate failed")
    out["log_mu"], out["logNeff_inj"] = float(out["log_mu"][0]), float(out["logNeff_inj"][0])
    with np.errstate(over="ignore", invalid="ignore"):  # analysis.py:86-88,131-135
        out["var"] = 1.0 / np.exp(out["logNeff"]) - 1.0 / (m.offsets[1] - m.offsets[0])
        out["var_inj"] = 1.0 / np.exp(out["logNeff_inj"]) - 1.0 / float(total_inj)
    out["logZ"] = out["logZ"][: len(spec.groups)]
    out.update(jac)
    return out
