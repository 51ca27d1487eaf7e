"""Few energies on the C3 table (1M-point grid): sweep time against the segment count and the combine
kernel (EPS_OPT_SCAN_SEGMENTS / EPS_OPT_SCAN_COMBINE).  The serial combine capped the cut at 64 segments."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

w = W.c3()
ctx = cabi.Context(0)
for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    ctx.set_potentials(w["V"], w["s"])
    n_steps = ctx.curve_info(0).n_steps
    for nE in (256, 1024, 4096):
        ref = None
        for seg, comb in ((1, 0), (18, 1), (18, 2), (64, 1), (64, 2), (0, 0), (0, 1), (296, 2), (489, 2)):
            ctx.set_option(ctx.OPT_SCAN_SEGMENTS, seg)
            ctx.set_option(ctx.OPT_SCAN_COMBINE, comb)
            for _ in range(2):
                ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
            ctx.sync()
            ctx.stats_reset()
            f0 = ctx.counter(ctx.CNT_SCAN_FLAGGED)
            reps = 5
            for _ in range(reps):
                ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
            st = ctx.stats()
            flagged = (ctx.counter(ctx.CNT_SCAN_FLAGGED) - f0) // reps
            n = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=True, tails=False)[0]
            if ref is None:
                ref = n
            rate = n_steps * nE * reps / (st.sweep_ms * 1e-3)
            print(f"form {form} nE {nE:5d} segments {seg:3d} combine {comb}: {st.sweep_ms / reps:8.3f} ms  {rate:.4g} steps/s  "
                  f"flagged {flagged}  nodes {'== sequential' if np.array_equal(n, ref) else 'MISMATCH'}", flush=True)
ctx.close()
