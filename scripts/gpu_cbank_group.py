"""Tuning: energy groups of the constant-bank sweep (EPS_OPT_CBANK_GROUP) on the C5 table: time per
sweep and node-count equality against the ungrouped order."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

nE = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
w = W.c5(nE=nE)
ctx = cabi.Context(0)
for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    ctx.set_potentials(w["V"], w["s"])
    n_steps = ctx.curve_info(0).n_steps
    ref = None
    for group, pdl in ((0, 1), (2, 1), (2, 3), (2, 4), (2, 6), (3, 1), (3, 4), (4, 1), (4, 4)):
        ctx.set_option(ctx.OPT_CBANK_GROUP, group)
        ctx.set_option(ctx.OPT_CBANK_PDL, pdl)
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        ctx.sync()
        ctx.stats_reset()
        reps = 3
        for _ in range(reps):
            ctx.l2_flush()
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        st = ctx.stats()
        rate = n_steps * nE * reps / (st.sweep_ms * 1e-3)
        n = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=True, tails=False)[0]
        if ref is None:
            ref = n
        print(f"form {form} group {group} pdl {pdl}: {st.sweep_ms / reps:9.3f} ms  steps/s {rate:.4g}  launches/sweep {st.kernel_launches // reps}"
              f"  nodes {'== ungrouped' if np.array_equal(n, ref) else 'MISMATCH'}", flush=True)
ctx.close()
