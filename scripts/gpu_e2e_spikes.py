"""Diagnostic (run under torchrun): C2 end-to-end steps exactly as bench.py times them, with host wall
clocks per phase, to find what the rare long steps consist of."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

os.environ.setdefault("OMP_WAIT_POLICY", "passive")
comm = bench.Comm()
from epseon_backend_b200 import cabi  # noqa: E402

ctx = cabi.Context(comm.local)
use_sampler = len(sys.argv) > 1 and sys.argv[1] == "sampler"
if use_sampler:
    sampler = bench.ClockSampler(comm.local)
    sampler.start()
comm.attach(ctx)
for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    wl = bench.Workload("c2", ctx, comm)
    for _ in range(10):
        wl.gather(wl.resident())
    rows = []
    for i in range(150):
        ctx.l2_flush()
        comm.barrier(ctx)
        ctx.timer_start()
        t0 = time.perf_counter()
        ctx.set_potentials(wl.V, wl.s)
        t1 = time.perf_counter()
        res = wl.e2e_tail()
        t2 = time.perf_counter()
        wl.gather(res)
        t3 = time.perf_counter()
        ms = ctx.timer_stop()
        rows.append((ms, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    a = np.array(rows)
    print(f"rank {comm.rank} form {form} sampler {use_sampler}: event ms median {np.median(a[:,0]):.3f} max {a[:,0].max():.3f}; "
          f"slow steps (event, set_potentials, solve, gather wall ms): "
          + "; ".join(f"#{i}: {r[0]:.1f} {r[1]:.1f} {r[2]:.1f} {r[3]:.1f}" for i, r in enumerate(rows) if r[0] > 2 * np.median(a[:, 0])), flush=True)
comm.close()
ctx.close()
