"""racecheck probe: the shared-memory ring with STAGE REUSE (more tiles than stages) on every CTA shape."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

N = 24000  # ~11 tiles of 2048 steps: 4-stage rings wrap twice, 2-stage rings five times
V = np.stack([W.morse(5500.0, 2.2, 1.6, 0.4, 9.0, N), W.lj(4800.0, 2.4, 0.4, 9.0, N)])
s = W.scale(20.0, 20.0, W.grid_h(0.4, 9.0, N))
lo, hi = V.min(axis=1), V[:, -1] - 1.0
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if len(sys.argv) > 2:  # a library built with other flags (tuning): scripts/bin/libepseon_cuda_<variant>.so
    cabi.LIB_PATH = Path(sys.argv[2]).resolve()
with cabi.Context(0) as ctx:
    ctx.set_potentials(V, s)
    if which in ("all", "512"):
        ctx.sweep_uniform(lo, hi, 600, tails=False)                  # <4,4>: 4 stages
        print("512 done", flush=True)
    if which in ("all", "256"):
        ctx.sweep_uniform(lo, hi, 200, tails=False)                  # <2,4>: 4 stages
        print("256 done", flush=True)
    if which in ("all", "128"):
        ctx.solve_levels(lo, hi, 64, 0, 7, 16, 1e-6, 2)              # packed rows, 128-energy CTAs (2 x 2): 2 stages
        print("128 done", flush=True)
    if which in ("all", "scan"):
        ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 2)
        ctx.sweep_uniform(lo, hi, 100, tails=False)                  # <2,8> scan: 4 stages, 6 tiles per segment
        print("scan done", flush=True)
    ctx.sync()
