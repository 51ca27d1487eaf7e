"""ncu target: the C3 sweep (1M-point tabulated curve, 4096 energies) through the scan path."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

ctx = cabi.Context(0)
w = W.c3()
ctx.set_potentials(w["V"], w["s"])
for _ in range(3):
    ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], nodes=False, tails=False)
ctx.sync()
