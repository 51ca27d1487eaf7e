"""Tuning: programmatic-dependent-launch trigger position of the constant-bank sweep (EPS_OPT_CBANK_PDL)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

w = W.c2()
ctx = cabi.Context(0)
ctx.set_potentials(w["V"], w["s"])
n_steps = ctx.curve_info(0).n_steps
ref = {}
for shape, pdl in ((0, 0), (4128, 0), (4128, 3), (4128, 4), (4128, 5), (4128, 7), (4128, 10), (2256, 4), (4256, 4)):
    ctx.set_option(ctx.OPT_CBANK, 1 if shape else 2)
    ctx.set_option(ctx.OPT_CBANK_PDL, pdl)
    if shape:
        ctx.set_option(ctx.OPT_CBANK_SHAPE, shape)
    for nE in (65536, 148 * 512, 148 * 1024, 1 << 20):
        for _ in range(2):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        ctx.sync()
        ctx.stats_reset()
        reps = 6
        for _ in range(reps):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        st = ctx.stats()
        rate = n_steps * nE * reps / (st.sweep_ms * 1e-3)
        same = ""
        if nE <= 148 * 512:
            n, m, x = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE)
            if shape == 0:
                ref[nE] = (n, m, x)
            else:
                r = ref[nE]
                same = "bits==tma" if (np.array_equal(n, r[0]) and np.array_equal(m.view(np.uint64), r[1].view(np.uint64))
                                       and np.array_equal(x, r[2])) else "MISMATCH"
        print(f"shape {shape:5d} pdl {pdl:2d} nE {nE:8d}  ms {st.sweep_ms / reps:8.3f}  steps/s {rate:.4g}  pipe {rate * 4 / (148 * 64 * 1.965e9):.3f} {same}",
              flush=True)
ctx.close()
