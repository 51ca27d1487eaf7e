"""First-contact GPU script: FP64 probe, sweep timings on C1/C2 shapes (not a bench line)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

out = {}
ctx = cabi.Context(0)
p = cabi.device_props(0)
out["device"] = dict(name=p.name.decode(), sms=p.sm_count, clock_khz=p.clock_khz, cc=f"{p.cc_major}.{p.cc_minor}")
tf, ms = ctx.fp64_probe()
out["fp64_probe"] = dict(tflops=tf, ms=ms)
print(out, flush=True)

for name, w, nEs in (("c1", W.c1(), [1024]), ("c2", W.c2(), [65536, 75776, 151552, 1 << 20])):
    ctx.set_potentials(w["V"], w["s"])
    n_steps = ctx.curve_info(0).n_steps
    for nE in nEs:
        for _ in range(3):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        ctx.sync()
        ctx.stats_reset()
        reps = 5
        for _ in range(reps):
            ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, nodes=False, tails=False)
        st = ctx.stats()
        ms = st.sweep_ms / reps
        rate = n_steps * nE / (ms * 1e-3)
        out[f"{name}_nE{nE}"] = dict(ms=ms, steps_per_s=rate, tflops7=7 * rate / 1e12, frac_of_probe=7 * rate / 1e12 / tf)
        print(name, nE, out[f"{name}_nE{nE}"], flush=True)

w = W.c2()
ctx.set_potentials(w["V"], w["s"])
for M in (4352, 4448):
    t = time.time()
    ctx.stats_reset()
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, M, 1e-10, 8)
    dt = time.time() - t
    st = ctx.stats()
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    out[f"c2_solve_M{M}"] = dict(wall_ms=dt * 1e3, sweep_ms=st.sweep_ms, launches=st.sweep_launches,
                                 steps=st.grid_steps, nb=int(nb[0]),
                                 max_rel_err=float(np.max(np.abs(lev[0] - exact) / exact)),
                                 max_width=float(wid.max()))
    print(out[f"c2_solve_M{M}"], flush=True)
tf2, _ = ctx.fp64_probe()
out["fp64_probe_after"] = tf2
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "probe.json").write_text(json.dumps(out, indent=1))
