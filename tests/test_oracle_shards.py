"""CPU: shard-wise oracle (partial_record / combine_records) == whole-catalog oracle, in process and
across two gloo ranks (the host-side logic of the multi-GPU path: injections split by index
range, whole events dealt round-robin, one all-gather of the partial records)."""

import os
import socket
import subprocess
import sys

import numpy as np

from oracle import popmodel
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(cols, rank, n, events):
    if events:
        return {k: v[rank::n] for k, v in cols.items()}
    I = next(iter(cols.values())).size
    return {k: v[rank * I // n : (rank + 1) * I // n] for k, v in cols.items()}


def test_partial_combine_equals_whole():
    for name in ("bspline_full_margsel", "plpeak", "bspline_iid"):
        c = cases.load_case(name)
        ev = popmodel.evaluate(c.low.spec, c.low.pe_cols, c.low.inj_cols, c.total_inj, c.Lam)
        l0, g0, _ = popmodel.hierarchical_log_likelihood(ev, c.Nobs, **c.like_kw)
        for n in (1, 2, 3):
            recs = [popmodel.partial_record(c.low.spec, _shard(c.low.pe_cols, r, n, True), _shard(c.low.inj_cols, r, n, False), c.Lam, want_g2=True) for r in range(n)]
            l1, g1, _ = popmodel.combine_records(c.low.spec, recs, c.total_inj, c.Lam, c.Nobs, **c.like_kw)
            assert abs(l1 - l0) <= 1e-12 * abs(l0)
            assert np.max(np.abs(g1 - g0)) <= 1e-10 * np.max(np.abs(g0))


def test_two_gloo_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_worker.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "GLOO_OK" in outs[0] and "GLOO_OK" in outs[1]
