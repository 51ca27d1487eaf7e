"""Short workload for ncu captures: C2 set-up, a few sweeps, one full level solve."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

nE = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
ctx = cabi.Context(0)
w = W.c2()
ctx.set_potentials(w["V"], w["s"])
for _ in range(3):
    ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
ctx.sync()
if "solve" in sys.argv:
    ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 4352, 1e-10, 8)
