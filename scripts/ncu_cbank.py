"""ncu target: one single-curve sweep through the constant-bank kernel (C2 table; energies = argv[1], default 2^19)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

nE = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
ctx = cabi.Context(0)
w = W.c2()
ctx.set_potentials(w["V"], w["s"])
ctx.set_option(ctx.OPT_CBANK, 1)
for _ in range(2):
    ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
ctx.sync()
