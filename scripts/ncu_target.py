"""Short workload for ncu captures: one BASELINE workload's resident step, a few times.

    python scripts/ncu_target.py c2|c3|c4|c4s|c5|cooley [form] [reps]

c2: 65 536-energy coarse sweep + refinement of 17 levels (TMA ring kernel, flat rows);
c3: 4096 energies x 1M grid (scan path); c4: 4096 curves x (1024 coarse + packed refinement rows);
c5: 2^24 energies x 200k grid (constant-bank kernel, energy groups);
cooley: C2 and a 512-curve batch through EPS_SOLVE_COOLEY.  form = 0 (X form) / 1 (D form).
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
form = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = cabi.Context(0)
ctx.set_option(ctx.OPT_FORM, form)
if which == "c2":
    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    for _ in range(reps):
        ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 2228, 1e-10, 8)
elif which == "c3":
    w = W.c3()
    ctx.set_potentials(w["V"], w["s"])
    for _ in range(reps):
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], 4096, nodes=False, tails=False)
elif which in ("c4", "c4s"):  # c4s: one GPU's share of C4 on an 8-GPU box (512 curves)
    w = W.c4() if which == "c4" else W.c4(512, 10_000, 1024)
    ctx.set_potentials(w["V"], w["s"])
    import time
    for _ in range(reps):
        t0 = time.perf_counter()
        ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 16, 1e-10, 12)
        print("solve wall ms", (time.perf_counter() - t0) * 1e3)
elif which == "c5":
    import os
    if "EPS_CB_GROUP" in os.environ:  # tuning: energy-group size of the constant-bank sweep (resident waves; 0 = one group)
        ctx.set_option(ctx.OPT_CBANK_GROUP, int(os.environ["EPS_CB_GROUP"]))
    w = W.c5()  # the full 2^24 energies: the carried state's path through L2 / HBM depends on the size
    ctx.set_potentials(w["V"], w["s"])
    for _ in range(reps):
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1 << 24, nodes=False, tails=False)
    print("cbank launches per sweep", ctx.counter(ctx.CNT_CBANK_LAUNCHES) // reps)
elif which == "cooley":
    ctx.set_option(ctx.OPT_FORM, 1)
    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    for _ in range(reps):
        ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 1, 1e-10, 40, flags=ctx.SOLVE_COOLEY)
    w = W.c4(512, 10_000, 1024)
    ctx.set_potentials(w["V"], w["s"])
    for _ in range(reps):
        ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 1, 1e-10, 40, flags=ctx.SOLVE_COOLEY)
ctx.sync()
print("ncu_target done", which, form)
