"""Deterministic synthetic catalogs of the BASELINE.json shapes.

There is no network for GWTC data, so benchmarks and parity tests run on seeded synthetic
catalogs with the reference's column schema (gwinferno/pipeline/utils.py:51-96):
PE columns ``(E, S)`` and found-injection columns ``(I,)`` named
``mass_1, mass_ratio, mass_2, a_1, a_2, cos_tilt_1, cos_tilt_2, redshift, prior``.

  * events: centres from a fiducial population, samples = centre + Gaussian scatter reflected
    into the support (so N_eff,i >> E);
  * injections: uniform over the support with ``prior`` = the analytic draw density (so
    N_eff,inj >> 4E); ``total_inj = 4 I``.
Seeds: ``1000*cfg + {0: PE, 1: injections, 2: Lambda}``.
"""

import numpy as np

from .cosmology import Planck15

MMIN, MMAX = 3.0, 100.0
ZLO, ZHI = 1e-3, 1.9

CONFIGS = {
    # cfg: (E, S, I)  -- BASELINE.json "configs"
    1: (69, 3000, 100_000),
    2: (70, 4000, 500_000),
    3: (300, 10_000, 100_000_000),
    5: (200, 8000, 20_000_000),
}


def _reflect(x, lo, hi):
    """Reflect values into [lo, hi] (keeps the scatter smooth at the boundaries)."""
    w = hi - lo
    y = np.mod(x - lo, 2.0 * w)
    y = np.where(y > w, 2.0 * w - y, y)
    return lo + y


INJ_BLOCK = 1 << 20


def _injection_block(seed, block, n):
    """Injections ``[block*INJ_BLOCK, block*INJ_BLOCK + n)``: every block has its own counter-based
    stream, so any index range can be generated independently (sharded runs never materialise the
    whole 1e8-injection set on one host)."""
    rng = np.random.default_rng([int(seed), int(block)])
    m1 = rng.uniform(MMIN, MMAX, n)
    qlo = MMIN / m1
    q = rng.uniform(qlo, 1.0, n)
    a1 = rng.uniform(0.0, 1.0, n)
    a2 = rng.uniform(0.0, 1.0, n)
    ct1 = rng.uniform(-1.0, 1.0, n)
    ct2 = rng.uniform(-1.0, 1.0, n)
    z = rng.uniform(ZLO, ZHI, n)
    # density of the draw in (m1, q, a1, a2, ct1, ct2, z)
    prior = 1.0 / (MMAX - MMIN) / (1.0 - qlo) / 4.0 / (ZHI - ZLO)
    return [m1, q, m1 * q, a1, a2, ct1, ct2, z, prior]


INJ_COLS = ["mass_1", "mass_ratio", "mass_2", "a_1", "a_2", "cos_tilt_1", "cos_tilt_2", "redshift", "prior"]


def make_injections(I, seed, start=0, stop=None):
    """Found injections ``[start, stop)`` of a set of ``I`` (uniform over the support, ``prior`` =
    the analytic draw density)."""
    stop = I if stop is None else stop
    out = [np.empty(stop - start) for _ in INJ_COLS]
    b0, b1 = start // INJ_BLOCK, (max(stop, start + 1) - 1) // INJ_BLOCK
    for b in range(b0, b1 + 1):
        lo, hi = b * INJ_BLOCK, min(I, (b + 1) * INJ_BLOCK)
        cols = _injection_block(seed, b, hi - lo)
        a, e = max(lo, start), min(hi, stop)
        if e <= a:
            continue
        for o, c in zip(out, cols):
            o[a - start : e - start] = c[a - lo : e - lo]
    return dict(zip(INJ_COLS, out))


# ---- the counter-based generator of libgwi's gwi_synth_injections (csrc/synth.cu), bit for bit ---------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on arrays of 32-bit words held in uint64 (so that 32 x 32 -> 64-bit products are exact)."""
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85, np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0), p1 & MASK, (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1), p0 & MASK
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def make_injections_philox(seed, first, count):
    """Injections ``[first, first + count)`` of the device generator (``gwi_synth_injections``): the same uniform draw over
    the support as :func:`make_injections`, but every injection takes its seven uniforms from Philox blocks keyed by its own
    GLOBAL index -- any range can be produced on any rank, on the device, without the rest."""
    i = np.arange(first, first + count, dtype=np.uint64)
    lo, hi = i & np.uint64(0xFFFFFFFF), i >> np.uint64(32)
    u = []
    for b in range(4):
        x, y, z, w = philox4x32_10(lo, hi, np.full(count, b, dtype=np.uint64), np.zeros(count, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for h, l in ((x, y), (z, w)):
            u.append((((h << np.uint64(32)) | l) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))
    m1 = MMIN + (MMAX - MMIN) * u[0]
    qlo = MMIN / m1
    q = qlo + (1.0 - qlo) * u[1]
    z = ZLO + (ZHI - ZLO) * u[6]
    prior = 1.0 / (MMAX - MMIN) / (1.0 - qlo) / 4.0 / (ZHI - ZLO)
    return dict(zip(INJ_COLS, [m1, q, m1 * q, u[2], u[3], -1.0 + 2.0 * u[4], -1.0 + 2.0 * u[5], z, prior]))


def make_events(E, S, seed):
    rng = np.random.default_rng(seed)
    # fiducial population for the event centres
    u = rng.uniform(size=E)
    a = -2.5 + 1.0
    m1c = (u * (90.0**a - 5.0**a) + 5.0**a) ** (1.0 / a)  # p(m1) ~ m1^-2.5 on [5, 90]
    qc = rng.uniform(0.4, 1.0, E)
    a1c = rng.beta(2.0, 5.0, E)
    a2c = rng.beta(2.0, 5.0, E)
    ct1c = rng.uniform(-1.0, 1.0, E)
    ct2c = rng.uniform(-1.0, 1.0, E)
    zg = np.linspace(0.01, 1.5, 4000)
    pz = Planck15.dVcdz(zg) * (1.0 + zg) ** 1.7
    cdf = np.cumsum(pz)
    cdf /= cdf[-1]
    zc = np.interp(rng.uniform(size=E), cdf, zg)

    def scat(c, sig, lo, hi):
        return _reflect(c[:, None] + sig * rng.standard_normal((E, S)), lo, hi)

    m1 = scat(m1c, 0.1 * m1c[:, None], MMIN + 1e-3, MMAX - 1e-3)
    q = scat(qc, 0.15, 0.0, 1.0)
    q = np.maximum(q, (MMIN + 1e-3) / m1)  # keep m2 >= mmin
    q = np.minimum(q, 1.0)
    a1 = scat(a1c, 0.15, 0.0, 1.0)
    a2 = scat(a2c, 0.15, 0.0, 1.0)
    ct1 = scat(ct1c, 0.4, -1.0, 1.0)
    ct2 = scat(ct2c, 0.4, -1.0, 1.0)
    z = scat(zc, 0.2 * zc[:, None], ZLO + 1e-3, ZHI - 1e-3)
    # sampling prior ~ the usual LVK PE prior shape: p(m1,q) ~ m1 (flat in component masses),
    # flat spins/tilts, p(z) ~ dVc/dz / (1+z); un-normalised constants do not matter for parity
    prior = m1 * Planck15.dVcdz(z) / (1.0 + z) * 1e-12
    return {
        "mass_1": m1,
        "mass_ratio": q,
        "mass_2": m1 * q,
        "a_1": a1,
        "a_2": a2,
        "cos_tilt_1": ct1,
        "cos_tilt_2": ct2,
        "redshift": z,
        "prior": prior,
    }


def make_catalog(E, S, I, cfg=0):
    """Return ``(pedict, injdict, constants)`` in the reference's schema."""
    pedict = make_events(E, S, 1000 * cfg + 0)
    injdict = make_injections(I, 1000 * cfg + 1)
    constants = {"total_inj": float(4 * I), "obs_time": 1.0, "nObs": E}
    return pedict, injdict, constants


def make_config(cfg, scale=1.0):
    """Catalog of BASELINE.json config ``cfg``; ``scale`` < 1 shrinks S and I for parity tests."""
    E, S, I = CONFIGS[cfg]
    S = max(8, int(round(S * scale)))
    I = max(64, int(round(I * scale)))
    return make_catalog(E, S, I, cfg=cfg)


def bspline_lambda(nsplines, seed, scale=1.0):
    """Test point for the B-spline model: coefficients ~ N(0, scale), c_z[0] = 0, lamb = 2.7."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, n in nsplines.items():
        out[k] = scale * rng.standard_normal(n)
    if "redshift" in out:
        out["redshift"][0] = 0.0
    out["lamb"] = 2.7
    return out


def effective_spins(d):
    """``(chi_eff, chi_p)`` of a sample dictionary (component spins -> effective spin and effective
    precession), the coordinates of ``BSplineEffectiveSpinDims`` (separable.py:706-778)."""
    q, a1, a2, c1, c2 = d["mass_ratio"], d["a_1"], d["a_2"], d["cos_tilt_1"], d["cos_tilt_2"]
    chi_eff = (a1 * c1 + q * a2 * c2) / (1.0 + q)
    s1 = np.sqrt(np.clip(1.0 - c1 * c1, 0.0, None))
    s2 = np.sqrt(np.clip(1.0 - c2 * c2, 0.0, None))
    chi_p = np.maximum(a1 * s1, q * (4.0 * q + 3.0) / (4.0 + 3.0 * q) * a2 * s2)
    return np.ascontiguousarray(chi_eff), np.ascontiguousarray(chi_p)
