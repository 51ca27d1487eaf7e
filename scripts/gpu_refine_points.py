"""Tuning: time to all levels against the refinement points per level per round (k-section needs
M / log2(M + 1) sweeps-worth of energies per bit, and small rounds are latency-bound)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

ctx = cabi.Context(0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ctx.sync()
    ms = []
    for _ in range(reps):
        ctx.l2_flush()
        ctx.timer_start()
        r = fn()
        ms.append(ctx.timer_stop())
    return float(np.mean(ms)), r


for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    ref = None
    for M in (4457, 2228, 1485, 1114, 557, 256):
        ms, r = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, M, 1e-10, 12))
        ref = r[0] if ref is None else ref
        st = ctx.stats()
        print(f"c2 form {form} M {M:5d}: {ms:7.3f} ms  max rel diff vs M=4457 {np.nanmax(np.abs(r[0] / ref - 1)):.2e}  "
              f"max width/E {np.nanmax(r[1] / np.abs(r[0])):.2e}", flush=True)
    w = W.c4()
    ctx.set_potentials(w["V"], w["s"])
    ref = None
    for M in (32, 16, 8):
        ms, r = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, M, 1e-10, 16), reps=3)
        ref = r[0] if ref is None else ref
        print(f"c4 form {form} M {M:5d}: {ms:7.3f} ms  max rel diff vs M=32 {np.nanmax(np.abs(r[0] / ref - 1)):.2e}  "
              f"max width/E {np.nanmax(r[1] / np.abs(r[0])):.2e}", flush=True)
    w = W.c4(512, 10_000, 1024)
    ctx.set_potentials(w["V"], w["s"])
    for M in (32, 16, 8):
        ms, r = timed(lambda: ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, M, 1e-10, 16), reps=5)
        print(f"c4/8 (512 curves) form {form} M {M:5d}: {ms:7.3f} ms", flush=True)
ctx.close()
