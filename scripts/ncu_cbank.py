"""ncu target: one large single-curve sweep through the constant-bank kernel (C2 table, 2^19 energies)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

ctx = cabi.Context(0)
w = W.c2()
ctx.set_potentials(w["V"], w["s"])
ctx.set_option(ctx.OPT_CBANK, 1)
for _ in range(2):
    ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1 << 19, nodes=False, tails=False)
ctx.sync()
