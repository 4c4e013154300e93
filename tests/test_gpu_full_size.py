"""GPU parity at BASELINE.json's FULL sizes (-m gpu): the CUDA path through the C-ABI, with the plan
geometry the product picks by itself (no chunk_steps / n_deep override: 148 CTAs, guided slices, the
reduction tree of that size), against the plain-C oracle (oracle/c/gwi_oracle.c: straight from the raw
coordinates, no plan, no sorting) on the same seeded catalog and the same Lambda.

Tolerances (north_star): |d log L| <= 1e-10 |log L|, gradient <= 1e-8 of its largest component.
configs[1] (70 x 4000 + 5e5), configs[4] (200 x 8000 + 2e7, IID spins + IID component masses) and
configs[2] (300 x 10 000 + 1e8, the headline; ~25 GB of host memory, the oracle takes a few seconds on
the box's cores).  The same check runs on the host warp emulator in tests/test_full_size_emulated.py;
this file is the one that sees hardware ordering.
"""

import os

import numpy as np
import pytest

from gwinferno_b200 import lowering, workloads
from gwinferno_b200.likelihood import PopulationLikelihood
from oracle import c_oracle, popmodel

pytestmark = pytest.mark.gpu

VAL_RTOL = 1e-10
GRAD_RTOL = 1e-8


def _check(name, world=1, rank=0, min_neff_cut=True, need_neff_grad=False, **like_kw):
    c_oracle.build()
    pe, inj, const, z_range = workloads.shard_catalog(name, rank, world, scale=1.0, all_reduce_minmax=lambda lo, hi: (lo, hi))
    weights, params_fn = workloads.build_model(const["family"], pe, inj, z_range=z_range)
    low, lam, _ = workloads.lower_workload(weights, params_fn, pe, inj)
    E = low.pe_cols[next(iter(low.pe_cols))].shape[0]
    eng = PopulationLikelihood(low, const["total_inj"], need_neff_grad=need_neff_grad)
    info = eng.info()
    out = []
    for step in (0, 5):  # two different Lambda: the a-priori shift bound and the sums both move
        lam_s = lam if step == 0 else lowering.flatten_params(weights(pe, True, params_fn(step)), low.spec.n_params)
        log_l, grad, head = eng.loglike(lam_s, Nobs=E, min_neff_cut=min_neff_cut, **like_kw)
        ev = c_oracle.evaluate(low.spec, low.pe_cols, low.inj_cols, const["total_inj"], lam_s, want_jac=True, want_neff_jac=need_neff_grad,
                               n_threads=os.cpu_count() or 1)
        l_o, g_o, _ = popmodel.hierarchical_log_likelihood(ev, E, min_neff_cut=min_neff_cut, **like_kw)
        assert head["passed"] == 1.0 and head["status"] == 0.0
        el, eg = abs(log_l - l_o) / abs(l_o), np.max(np.abs(grad - g_o)) / np.max(np.abs(g_o))
        print(f"{name} world={world} rank={rank} step={step}: log L {log_l:.12f} (oracle {l_o:.12f}) rel {el:.2e}; gradient rel {eg:.2e}; "
              f"{info['n_chunks']} chunks, n_deep={info['n_deep']}")
        assert el <= VAL_RTOL
        assert eg <= GRAD_RTOL
        out.append((el, eg))
    eng.model.close()
    return info, out


def test_cfg2_full_size_vs_c_oracle():
    info, _ = _check("cfg2")
    assert info["n_valid_inj"] > 400_000


def test_cfg2_full_size_vs_c_oracle_marginalized_selection():
    _check("cfg2", need_neff_grad=True, marginalize_selection=True)


def test_cfg5_full_size_vs_c_oracle():
    info, _ = _check("cfg5")
    assert info["n_padded"] > 20_000_000


def test_cfg3_full_size_vs_c_oracle():
    info, _ = _check("cfg3")
    assert info["n_padded"] > 100_000_000


def test_cfg3_rank3_of_8_bucket_shard_vs_c_oracle():
    """One rank's shard of the 8-way run bench.py times (bucket sharding of the injections, events
    round-robin): the partition changes the plan geometry (shorter sorted runs), not the sums."""
    _check("cfg3", world=8, rank=3, min_neff_cut=False)
