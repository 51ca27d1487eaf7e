"""Quick GPU check: sweep rate on C2 for forced CTA shapes."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi
from tests import workloads as W
w = W.c2()
for ept, warps in ((2, 8), (2, 4), (4, 4), (4, 8)):
    os.environ["EPS_FORCE_EPT"] = str(ept)
    os.environ["EPS_FORCE_WARPS"] = str(warps)
    ctx = cabi.Context(0)
    ctx.set_potentials(w["V"], w["s"])
    n_steps = ctx.curve_info(0).n_steps
    for nE in (65536, 69632, 148 * 1024, 148 * 2048, 1 << 20):
        for _ in range(2):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        ctx.sync(); ctx.stats_reset()
        for _ in range(4):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        st = ctx.stats()
        rate = n_steps * nE * 4 / (st.sweep_ms * 1e-3)
        print(f"ept{ept} w{warps}", nE, "ms %.3f" % (st.sweep_ms / 4), "steps/s %.4g" % rate, "pipe %.3f" % (rate * 4 / (148 * 64 * 1.965e9)), flush=True)
    ctx.close()
