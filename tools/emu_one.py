import os, sys, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=True)
os.environ.setdefault("GWI_EMU_SMS", "4")
from tests import emu
emu.activate("")
import numpy as np
from tests import cases
from gwinferno_b200.likelihood import PopulationLikelihood
c = cases.load_case(sys.argv[1] if len(sys.argv) > 1 else "bspline_full")
eng = PopulationLikelihood(c.low, c.total_inj, need_neff_grad=False, chunk_steps=int(os.environ.get("CS", "8")))
print(eng.info())
log_l, grad, head = eng.loglike(c.Lam, Nobs=c.Nobs, **c.like_kw)
print(log_l, float(c.out["log_l"]), head)
G = cases.golden_jac_in_lambda_order(c, "log_l")
print("grad err", np.max(np.abs(grad - G)) / np.max(np.abs(G)))
