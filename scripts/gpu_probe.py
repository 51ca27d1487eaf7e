"""GPU tuning script: FP64 probe, sweep timings for EPT variants on C2 shapes (not a bench line)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

out = {}
w = W.c2()
exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
for ept, stride in ((1, 1), (2, 1), (1, 8), (2, 8), (1, 32), (2, 32)):
    os.environ["EPS_FORCE_EPT"] = str(ept)
    os.environ["EPS_FORCE_STRIDE"] = str(stride)
    ctx = cabi.Context(0)
    tf, ms = ctx.fp64_probe()
    print("probe", tf, ms, flush=True)
    ctx.set_potentials(w["V"], w["s"])
    n_steps = ctx.curve_info(0).n_steps
    for nE in (65536, 148 * 512, 148 * 1024, 148 * 2048, 1 << 20):
        for _ in range(2):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        ctx.sync()
        ctx.stats_reset()
        reps = 4
        for _ in range(reps):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        st = ctx.stats()
        ms = st.sweep_ms / reps
        rate = n_steps * nE / (ms * 1e-3)
        key = f"ept{ept}_s{stride}_nE{nE}"
        out[key] = dict(ms=ms, steps_per_s=rate, tflops6=6 * rate / 1e12, frac_of_probe=6 * rate / 1e12 / tf)
        print(key, out[key], flush=True)
    M = {1: 4352, 2: 4096}[ept]
    t = time.time()
    ctx.stats_reset()
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, M, 1e-10, 8)
    dt = time.time() - t
    st = ctx.stats()
    out[f"ept{ept}_s{stride}_solve"] = dict(wall_ms=dt * 1e3, sweep_ms=st.sweep_ms, launches=st.sweep_launches,
                                  steps=st.grid_steps, nb=int(nb[0]),
                                  max_rel_err=float(np.max(np.abs(lev[0] - exact) / exact)),
                                  max_width=float(wid.max()))
    print(out[f"ept{ept}_s{stride}_solve"], flush=True)
    ctx.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "probe.json").write_text(json.dumps(out, indent=1))
