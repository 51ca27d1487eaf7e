"""C4 refinement with 256- vs 128-energy packed CTAs by per-device batch size (the 8-GPU shard is 512 curves)."""
import sys

sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge

ge.build()
from epseon_backend_b200 import cabi
from tests import workloads as W

ctx = cabi.Context(0)
for form in (0, 1):
    ctx.set_option(ctx.OPT_FORM, form)
    for nC in (512, 1024, 2048, 4096):
        w = W.c4(nC, 10_000, 1024)
        ctx.set_potentials(w["V"], w["s"])
        res = {}
        for opt, name in ((2, "256"), (1, "128"), (0, "auto")):
            ctx.set_option(ctx.OPT_PACK128, opt)
            best = 1e9
            for _ in range(4):
                ctx.timer_start()
                ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 32, 1e-10, 8)
                best = min(best, ctx.timer_stop())
            res[name] = best
        print(f"form {form} curves {nC}: " + "  ".join(f"{k} {v:.3f} ms" for k, v in res.items()), flush=True)
