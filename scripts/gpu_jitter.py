"""Does polling NVML from a thread disturb short steps?  C2 solves, per-step CUDA-event time, with the
bench's clock sampler at several periods and without it."""
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

ctx = cabi.Context(0)
w = W.c2()
for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    ctx.set_potentials(w["V"], w["s"])
    for period in (None, 0.02, 0.1, 0.5):
        stop = [False]
        if period is not None:
            import pynvml as n
            n.nvmlInit()
            h = n.nvmlDeviceGetHandleByIndex(0)

            def loop():
                while not stop[0]:
                    n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                    n.nvmlDeviceGetCurrentClocksEventReasons(h)
                    n.nvmlDeviceGetPowerUsage(h)
                    time.sleep(period)
            threading.Thread(target=loop, daemon=True).start()
        ms = []
        for _ in range(300):
            ctx.l2_flush()
            ctx.sync()
            ctx.timer_start()
            ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 2228, 1e-10, 8)
            ms.append(ctx.timer_stop())
        stop[0] = True
        time.sleep(0.6)
        ms = np.array(ms)
        print(f"form {form} nvml period {period}: mean {ms.mean():.3f} median {np.median(ms):.3f} p99 {np.percentile(ms, 99):.3f} max {ms.max():.3f} ms", flush=True)
ctx.close()
