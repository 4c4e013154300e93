"""Worker of tests/test_oracle_shards.py::test_two_gloo_ranks: each rank evaluates its shard with
the oracle, the records are all-gathered over gloo (the CPU stand-in for the NCCL all-gather of
bench.py), and every rank combines them identically."""

import os

import numpy as np
import torch
import torch.distributed as dist

from oracle import popmodel
from tests import cases

dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, n = dist.get_rank(), dist.get_world_size()
c = cases.load_case("bspline_full_margsel")
pe = {k: v[rank::n] for k, v in c.low.pe_cols.items()}
I = c.low.inj_cols["c0"].size
inj = {k: v[rank * I // n : (rank + 1) * I // n] for k, v in c.low.inj_cols.items()}
rec = torch.from_numpy(popmodel.partial_record(c.low.spec, pe, inj, c.Lam, want_g2=True))
allrec = [torch.empty_like(rec) for _ in range(n)]
dist.all_gather(allrec, rec)
log_l, grad, _ = popmodel.combine_records(c.low.spec, torch.stack(allrec).numpy(), c.total_inj, c.Lam, c.Nobs, **c.like_kw)
gold = float(c.out["log_l"])
G = cases.golden_jac_in_lambda_order(c, "log_l")
assert abs(log_l - gold) <= 1e-10 * abs(gold), (log_l, gold)
assert np.max(np.abs(grad - G)) <= 1e-8 * np.max(np.abs(G))
# every rank must hold the identical result
t = torch.tensor([log_l])
ts = [torch.empty_like(t) for _ in range(n)]
dist.all_gather(ts, t)
assert all(float(x) == float(ts[0]) for x in ts)
dist.destroy_process_group()
print("GLOO_OK")
