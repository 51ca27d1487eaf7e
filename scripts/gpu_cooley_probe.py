"""Timing probe of the Cooley level search (run on the GPU box): coarse sweep alone vs coarse + Cooley."""
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge

ge.build()
from epseon_backend_b200 import cabi
from tests import workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "both"
ctx = cabi.Context(0)
ctx.set_option(ctx.OPT_FORM, 1)


def timed(fn, reps=5):
    fn()
    ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.timer_start()
        fn()
        best = min(best, ctx.timer_stop())
    return best


if which in ("c2", "both"):
    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    t_sweep = timed(lambda: ctx.sweep_uniform(w["E_lo"], w["E_hi"], 65536, nodes=False, tails=False))
    t_cool = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 1, 1e-10, 40, flags=ctx.SOLVE_COOLEY))
    t_cool4k = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 4096, 0, 16, 1, 1e-10, 40, flags=ctx.SOLVE_COOLEY))
    print(f"c2: coarse sweep {t_sweep:.3f} ms, coarse + cooley {t_cool:.3f} ms, 4096-coarse + cooley {t_cool4k:.3f} ms")
if which in ("c4", "both"):
    nC = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    w = W.c4(nC, 10_000, 1024)
    ctx.set_potentials(w["V"], w["s"])
    t_sweep = timed(lambda: ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1024, nodes=False, tails=False), 3)
    t_cool = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 1, 1e-10, 40, flags=ctx.SOLVE_COOLEY), 3)
    print(f"c4 ({nC} curves): coarse sweep {t_sweep:.3f} ms, coarse + cooley {t_cool:.3f} ms")
