"""CPU (no GPU): the host plan builder of libgwi.so.

The plan is read back through the gwi_debug_plan_* hooks and DECODED here in NumPy (test code
only): per sample, unpack the (piece, offset) words, rebuild the 4 tap weights, add the linear
terms and the static log-weight, and reduce per segment.  The result must equal the oracle on
the golden cases -- this pins masks/cuts, packing, sorting, lane layout, padding and the Monte
Carlo denominators without touching a GPU.  It is a checker of the static data, not a compute
path: the product has no CPU evaluation."""

import ctypes as C

import numpy as np
import pytest

from gwinferno_b200 import capi
from oracle import popmodel
from tests import cases

JMASK = np.uint64(63)
SPLINE_CASES = ["bspline_full", "bspline_iid", "bspline_indep_masses", "inference_test_bspline", "bspline_effspin", "bspline_symchieff"]


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    for name in capi.SYMBOLS:
        assert hasattr(lib, name), name
    assert lib.gwi_version() == capi.GWI_VERSION == 5
    # header and binding agree on the symbol list
    import os, re

    hdr = open(os.path.join(os.path.dirname(capi._HERE), "include", "gwi.h")).read()
    declared = set(re.findall(r"^(?:int|void|int64_t|double|const char\*)\s+(gwi_[a-z_]+)\s*\(", hdr, flags=re.M))
    assert declared == set(capi.SYMBOLS)


def test_struct_layouts_match_header_sizes():
    # sizes implied by include/gwi.h on LP64
    assert C.sizeof(capi.gwi_term) == 4 * 4 + 2 * 4 + 6 * 4 + 4 * 8 + 2 * 4 + 4 * 8 + 8 + 8 + 2 * 4
    assert C.sizeof(capi.gwi_cut) == 4 + 8 + 4 + 16 or C.sizeof(capi.gwi_cut) == 32
    assert C.sizeof(capi.gwi_like_opts) == 16
    assert C.sizeof(capi.gwi_model_info) == 6 * 8 + 10 * 4 + 2 * 4 + 5 * 8
    assert C.sizeof(capi.gwi_catalog_desc) == 56


def _decode(case, plan):
    dims = plan.read(0)
    n_cols, n_pad, n_chunks, n_seg, ns, n_kops = (int(x) for x in dims[:6])
    lw = int(dims[10])  # warps that share a chunk (CTA-cooperative geometry), 1 for the one-role kernel
    cols = np.ascontiguousarray(plan.read(1, dtype=np.uint64).reshape(n_pad // 64, n_cols, 64).transpose(1, 0, 2).reshape(n_cols, n_pad))
    chunks = plan.read(2).reshape(n_chunks, 4)
    segs = plan.read(3).reshape(n_seg, 4)
    dimt = plan.read(4).reshape(ns, 4)
    kops = plan.read(5).reshape(n_kops, 8)
    spec, Lam = case.low.spec, case.Lam
    x = np.zeros(n_pad)
    x += cols[n_cols - 1].view(np.float64)  # static log-weight (-inf on lane padding)
    for d in range(ns):
        term = spec.terms[int(dimt[d, 0])]
        rows = int(dimt[d, 1])
        J = (cols[d] & JMASK).astype(np.int64)
        u = cols[d].view(np.float64) + 0.5  # the word is w = u - 1/2 (piece index in its 6 low mantissa bits)
        assert J.max() <= rows - 1
        real = J < rows - 1
        Jc = np.where(real, J, 0)
        w = np.stack([(1 - u) ** 3, 3 * u**3 - 6 * u**2 + 4, -3 * u**3 + 3 * u**2 + 3 * u + 1, u**3], -1) / 6.0
        c = Lam[term.slots[0] : term.slots[0] + term.n_splines]
        f = np.einsum("nk,nk->n", w, c[Jc[:, None] + np.arange(4)])
        if term.kind == popmodel.TERM_SPLINE_LINEAR:  # the spline is the density: log of the cubic
            assert real.all()
            with np.errstate(divide="ignore", invalid="ignore"):
                f = np.where(f > 0, np.log(np.where(f > 0, f, 1.0)), -np.inf)
        x += np.where(real, f, 0.0)
    for k in range(n_kops):
        kind, c0, _, s0 = (int(v) for v in kops[k, :4])
        assert kind == 1, "only linear ops in the spline-family cases"
        off = np.array([kops[k, 7]], dtype=np.int64).view(np.float64)[0]
        x += (Lam[s0] + off) * cols[c0].view(np.float64)
    seg_of = np.full(n_pad, -1)
    for seg, first, steps, _ in chunks:
        seg_of[first : first + steps * 32 * lw] = seg
    assert (seg_of >= 0).all()
    return x, seg_of, segs, chunks, cols, dimt, lw


@pytest.mark.parametrize("name", SPLINE_CASES)
def test_plan_reproduces_oracle(name):
    case = cases.load_case(name)
    cat = capi.Catalog(case.low.pe_cols, case.low.inj_cols, case.total_inj)
    plan = capi.HostPlan(cat, case.low.spec, chunk_steps=8)
    x, seg_of, segs, chunks, cols, dimt, lw = _decode(case, plan)
    ev = popmodel.evaluate(case.low.spec, case.low.pe_cols, case.low.inj_cols, case.total_inj, case.Lam, want_jac=False)
    sumZ = np.sum(ev["logZ"])
    E = case.Nobs
    assert segs.shape[0] == E + 1
    for s in range(E + 1):
        xs = x[seg_of == s]
        n_total, n_valid = int(segs[s, 0]), int(segs[s, 1])
        assert np.isfinite(xs).sum() == n_valid
        m = xs.max()
        S1 = np.exp(xs - m).sum()
        S2 = np.exp(2 * (xs - m)).sum()
        if s == 0:
            logmean = m + np.log(S1) - np.log(case.total_inj) - sumZ
            logneff = 2 * np.log(S1) - np.log(S2 - S1**2 / case.total_inj)
            assert abs(logmean - ev["log_mu"]) < 1e-11
            assert abs(logneff - ev["logNeff_inj"]) < 1e-11
            assert n_total == case.low.inj_cols["c0"].size
        else:
            logmean = m + np.log(S1) - np.log(n_total) - sumZ
            logneff = 2 * np.log(S1) - np.log(S2)
            assert abs(logmean - ev["logBF"][s - 1]) < 1e-11
            assert abs(logneff - ev["logNeff"][s - 1]) < 1e-11
    # lane runs are sorted by the composite piece key (most pieces = most significant)
    ns = dimt.shape[0]
    key = np.zeros(cols.shape[1], dtype=np.int64)
    for d in range(ns):
        key = key * 64 + (cols[d] & JMASK).astype(np.int64)
    valid = np.isfinite(cols[-1].view(np.float64))
    for seg, first, steps, _ in chunks[:50]:
        blk = np.arange(first, first + steps * 32 * lw).reshape(lw, steps // 2, 32, 2)  # [warp][iter][lane][unroll]
        lane_major = blk.transpose(0, 2, 1, 3).reshape(32 * lw, steps)  # each lane's run in order
        kk = key[lane_major].reshape(-1)
        vv = valid[lane_major].reshape(-1)
        assert np.all(np.diff(kk[vv]) >= 0)
        assert not vv[np.argmin(vv) :].any() if not vv.all() else True  # padding only at the tail


@pytest.mark.parametrize("name", ["plpeak", "inference_test_parametric"])
def test_plan_parametric_counts(name):
    case = cases.load_case(name)
    cat = capi.Catalog(case.low.pe_cols, case.low.inj_cols, case.total_inj)
    plan = capi.HostPlan(cat, case.low.spec)
    segs = plan.read(3).reshape(-1, 4)
    _, valid_pe, _ = popmodel.log_weights(case.low.spec, case.low.pe_cols, case.Lam)
    _, valid_inj, _ = popmodel.log_weights(case.low.spec, case.low.inj_cols, case.Lam)
    assert segs[0, 1] == valid_inj.sum()
    assert np.array_equal(segs[1:, 1], valid_pe.sum(axis=1))


def test_bad_descriptions_are_rejected():
    case = cases.load_case("inference_test_bspline")
    cat = capi.Catalog(case.low.pe_cols, case.low.inj_cols, case.total_inj)
    import copy

    spec = copy.deepcopy(case.low.spec)
    spec.terms[0].n_splines = 3
    with pytest.raises(capi.GwiError):
        capi.HostPlan(cat, spec)
    spec = copy.deepcopy(case.low.spec)
    spec.n_params = 2  # coefficient slots out of range
    with pytest.raises(capi.GwiError):
        capi.HostPlan(cat, spec)
    with pytest.raises(capi.GwiError):
        capi.Catalog(case.low.pe_cols, case.low.inj_cols, total_inj=1.0)  # fewer than found


def test_model_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    case = cases.load_case("inference_test_bspline")
    cat = capi.Catalog(case.low.pe_cols, case.low.inj_cols, case.total_inj)
    with pytest.raises(capi.GwiError) as e:
        capi.Model(cat, case.low.spec)
    assert "no CPU fallback" in str(e.value)


class _Adhoc:
    pass


@pytest.mark.parametrize("seed", range(12))
def test_plan_reproduces_oracle_on_random_small_models(seed):
    """Randomised edge cases of the host plan builder: tiny and single-sample events, events whose
    samples are all masked, few or no valid injections, narrow supports, minimal spline counts, every
    basis flavour, several chunk lengths -- the decoded plan must give the oracle's per-segment sums."""
    from gwinferno_b200 import lowering, synthetic
    from gwinferno_b200 import models as M

    rng = np.random.default_rng(1000 + seed)
    E, S, I = int(rng.integers(1, 6)), int(rng.integers(1, 40)), int(rng.integers(40, 400))
    pe, inj, const = synthetic.make_catalog(E, S, I, cfg=200 + seed)
    mmin = float(rng.choice([3.0, 8.0, 20.0]))  # the larger values mask many samples (all of some events)
    mmax = float(rng.choice([60.0, 100.0]))
    nm, nq, na = int(rng.integers(4, 14)), int(rng.integers(4, 9)), int(rng.integers(4, 8))
    mass = M.BSplinePrimaryBSplineRatio(nm, nq, pe["mass_1"], inj["mass_1"], pe["mass_ratio"], inj["mass_ratio"], m1min=mmin, m2min=mmin, mmax=mmax)
    flavour = seed % 3
    if flavour == 0:
        spin = M.BSplineSpinMagnitude(na, pe["a_1"], inj["a_1"], normalize=True)
    elif flavour == 1:
        spin = M.BSplineChiPrecess(na, pe["a_1"], inj["a_1"], normalize=True)  # the spline is the density
    else:
        spin = M.BSplineSpinTilt(na, pe["cos_tilt_1"], inj["cos_tilt_1"], xrange=(-0.5, 0.7), normalize=True)  # narrow support
    zmod = M.PowerlawSplineRedshiftModel(int(rng.integers(4, 8)), pe["redshift"], inj["redshift"])
    p = dict(m=0.5 * rng.standard_normal(nm), q=0.5 * rng.standard_normal(nq), a=np.exp(0.4 * rng.standard_normal(na)), lamb=np.float64(rng.uniform(-1, 4)),
             z=0.4 * rng.standard_normal(zmod.n_splines))

    def w(d, pe_samples):
        return mass(p["m"], p["q"], pe_samples=pe_samples) * spin(p["a"], pe_samples=pe_samples) * zmod(d["redshift"], p["lamb"], p["z"]) / d["prior"]

    case = _Adhoc()
    case.low = lowering.lower(w(pe, True), w(inj, False))
    case.Lam = lowering.flatten_params(w(pe, True), case.low.spec.n_params)
    cat = capi.Catalog(case.low.pe_cols, case.low.inj_cols, const["total_inj"])
    plan = capi.HostPlan(cat, case.low.spec, chunk_steps=int(rng.choice([0, 4, 8, 32])))
    x, seg_of, segs, chunks, cols, dimt, _ = _decode(case, plan)
    ev = popmodel.evaluate(case.low.spec, case.low.pe_cols, case.low.inj_cols, const["total_inj"], case.Lam, want_jac=False)
    sumZ = np.sum(ev["logZ"])
    assert segs.shape[0] == E + 1
    with np.errstate(divide="ignore", invalid="ignore"):
        for s in range(E + 1):
            xs = x[seg_of == s]
            n_total, n_valid = int(segs[s, 0]), int(segs[s, 1])
            assert np.isfinite(xs).sum() == n_valid
            want = ev["log_mu"] if s == 0 else ev["logBF"][s - 1]
            if n_valid == 0:
                assert not np.isfinite(want)  # the oracle agrees that the segment has no weight
                continue
            m = xs.max()
            S1, S2 = np.exp(xs - m).sum(), np.exp(2 * (xs - m)).sum()
            denom = const["total_inj"] if s == 0 else n_total
            assert abs(m + np.log(S1) - np.log(denom) - sumZ - want) < 1e-11
            if s > 0:
                assert abs(2 * np.log(S1) - np.log(S2) - ev["logNeff"][s - 1]) < 1e-11


def test_explicit_knot_vectors_are_validated_by_the_library():
    """gwi_term.knots / order (interpolation.py:72-106): order 1..4, n_knots == n_splines + order, non-decreasing knots,
    at most 61 polynomial pieces inside the range; the piece tables are derived inside the library."""
    import copy

    from gwinferno_b200 import models as M
    from gwinferno_b200 import lowering, synthetic

    pe, inj, _ = synthetic.make_catalog(3, 40, 500, cfg=402)
    cat = None

    def plan(**kw):
        nonlocal cat
        a = M.BSplineSpinMagnitude(kw.pop("n", 8), pe["a_1"], inj["a_1"], normalize=True, **kw)
        c = np.zeros(a.n_splines)
        low = lowering.lower(a(c, pe_samples=True), a(c, pe_samples=False))
        cat = capi.Catalog(low.pe_cols, low.inj_cols, 2000.0)
        return low, cat

    low, cat = plan(degree=2)
    p = capi.HostPlan(cat, low.spec)
    dims = p.read(4).reshape(-1, 4)
    # quadratic, 8 bases on default knots: 6 spans inside [0, 1], the span that STARTS at 1 (half-open spans: x == 1 falls into
    # it, interpolation.py:143-146), and the dummy row
    assert dims[0, 1] == 6 + 1 + 1
    p.close()
    # a knot vector the library must refuse: decreasing
    bad = copy.deepcopy(low.spec)
    bad.terms[0].knots = bad.terms[0].knots[::-1].copy()
    with pytest.raises(capi.GwiError, match="non-decreasing"):
        capi.HostPlan(cat, bad)
    # wrong length / order
    bad = copy.deepcopy(low.spec)
    bad.terms[0].order = 5
    with pytest.raises(capi.GwiError, match="order 1..4"):
        capi.HostPlan(cat, bad)
    # too many pieces for the 6-bit piece index
    low, cat = plan(n=80, knots=np.linspace(-0.05, 1.05, 84))
    with pytest.raises(capi.GwiError, match="61 polynomial pieces"):
        capi.HostPlan(cat, low.spec)
    # the reference's default knots passed EXPLICITLY give the default plan's pieces (same words up to rounding of u)
    a_def = M.BSplineSpinTilt(9, pe["cos_tilt_1"], inj["cos_tilt_1"], normalize=True)
    dx = 2.0 / (9 - 2 - 1)
    a_exp = M.BSplineSpinTilt(9, pe["cos_tilt_1"], inj["cos_tilt_1"], normalize=True, knots=np.linspace(-1 - 3 * dx, 1 + 3 * dx, 13))
    c = np.linspace(-1, 1, 9)
    words = []
    for a in (a_def, a_exp):
        low = lowering.lower(a(c, pe_samples=True), a(c, pe_samples=False))
        cat = capi.Catalog(low.pe_cols, low.inj_cols, 2000.0)
        hp = capi.HostPlan(cat, low.spec)
        d = hp.read(0)
        w = hp.read(1, dtype=np.uint64).reshape(int(d[1]) // 64, int(d[0]), 64)[:, 0, :].ravel()
        words.append(w)
        hp.close()
    assert np.array_equal(words[0] & JMASK, words[1] & JMASK)
    assert np.max(np.abs(words[0].view(np.float64) - words[1].view(np.float64))) < 1e-12
