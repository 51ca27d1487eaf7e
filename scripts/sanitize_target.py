"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
TMA sweep (tails, packed + flat refinement rows), scan path + fix-up, constant-bank chunks,
device prep, centrifugal expansion, spline evaluation, wavefunctions; round 2: the D form on every
route, the Cooley search kernel (box and open tail), a two-context group with its peer copies, the
stop flag; the block-level prefix combine, the constant-bank energy groups and both 128-energy CTA
shapes.  Sizes are tiny: the sanitizer slows kernels 10-100x."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from epseon_backend_b200 import cabi  # noqa: E402
from tests import workloads as W  # noqa: E402

N = 6000
h = W.grid_h(0.4, 9.0, N)
V = np.stack([W.morse(5500.0, 2.2, 1.6, 0.4, 9.0, N), W.lj(4800.0, 2.4, 0.4, 9.0, N)])
s = W.scale(20.0, 20.0, h)
with cabi.Context(0) as ctx:
    ctx.set_potentials(V, s)
    lo, hi = V.min(axis=1), V[:, -1] - 1.0
    ctx.sweep_uniform(lo, hi, 700)                                     # TMA kernel, tails
    lev, wid, nb = ctx.solve_levels(lo, hi, 512, 0, 5, 64, 1e-10, 6)   # packed refinement rows
    ctx.wavefunctions(lev, np.full(2, h))
    ctx.set_potentials_rot(V, s, 0.4, h, [0, 3])                      # centrifugal expansion + prep
    ctx.sweep_uniform(np.repeat(lo, 2), np.repeat(hi, 2), 64, tails=False)
    ctx.set_potentials(V[0], s)
    ctx.solve_levels(lo[0], hi[0], 1024, 0, 5, 300, 1e-10, 6)          # flat refinement rows (one curve)
    ctx.set_option(ctx.OPT_CBANK, 1)                                   # constant-bank chunks (2 launches)
    ctx.sweep_uniform(lo[0], hi[0], 1500)
    ctx.set_option(ctx.OPT_CBANK, 2)
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 3)                           # transfer-matrix scan + combine (+ fix-up)
    ctx.sweep_uniform(lo[0], hi[0], 300, tails=False)
    ctx.solve_levels(lo[0], hi[0], 256, 0, 3, 32, 1e-10, 6)
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
    rk = np.concatenate([np.linspace(0.4, 4.0, 40), np.linspace(4.5, 9.0, 8)])
    ctx.spline_resample(rk, W.morse(5500.0, 2.2, 1.6, 0.4, 9.0, 2)[0] + 5500.0 * (1 - np.exp(-1.6 * (rk - 2.2))) ** 2, 0.4, 9.0, 3000)
    # ---- round 2: accurate recurrence (D form) on every route + Cooley search + cancellation flag
    ctx.set_option(ctx.OPT_FORM, 1)
    ctx.set_potentials(V, s)
    ctx.sweep_uniform(lo, hi, 700)
    ctx.solve_levels(lo, hi, 512, 0, 5, 64, 1e-10, 6)
    ctx.solve_levels(lo, hi, 512, 0, 5, 1, 1e-12, 30, flags=ctx.SOLVE_COOLEY)
    ctx.solve_levels(lo, hi, 512, 0, 5, 1, 1e-12, 30, flags=ctx.SOLVE_COOLEY | ctx.SOLVE_OPEN_TAIL)
    ctx.set_potentials(V[0], s)
    ctx.set_option(ctx.OPT_CBANK, 1)
    ctx.sweep_uniform(lo[0], hi[0], 1500)
    ctx.set_option(ctx.OPT_CBANK, 2)
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 3)
    ctx.solve_levels(lo[0], hi[0], 256, 0, 3, 32, 1e-10, 6)
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
    ctx.request_stop()
    try:
        ctx.sweep_uniform(lo[0], hi[0], 600)
    except cabi.EpsError as e:
        assert e.code == cabi.EPS_ERR_CANCELLED
    ctx.reset_stop()
    ctx.sync()
# ---- round 2, later: block-level prefix combine, constant-bank energy groups (cta_base != 0), 128-energy CTA shapes
with cabi.Context(0) as ctx:
    for form in (0, 1):
        ctx.set_option(ctx.OPT_FORM, form)
        ctx.set_potentials(V[0], s)
        ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 3)
        ctx.set_option(ctx.OPT_SCAN_COMBINE, 2)                        # prefix kernel (16 lanes, 3 active)
        ctx.sweep_uniform(lo[0], hi[0], 300)
        ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
        ctx.set_option(ctx.OPT_SCAN_COMBINE, 0)
        ctx.set_option(ctx.OPT_CBANK, 1)
        ctx.set_option(ctx.OPT_CBANK_GROUP, 1)                         # 1184 CTAs per group: the second group starts at cta_base 1184
        ctx.sweep_uniform(lo[0], hi[0], 1184 * 512 + 700, tails=False)
        ctx.set_option(ctx.OPT_CBANK, 0)
        ctx.set_option(ctx.OPT_CBANK_GROUP, 2)
        ctx.solve_levels(lo[0], hi[0], 1024, 0, 5, 40, 1e-10, 8)       # flat rows in 128-energy CTAs (2 chains x 2 warps)
        ctx.set_potentials(V, s)
        ctx.solve_levels(lo, hi, 256, 0, 7, 16, 1e-10, 8)              # packed rows, 128-energy CTAs, one wave (2 x 2)
        Vm = np.ascontiguousarray(np.tile(V, (320, 1))[:, :1500])      # 640 short curves: more than one resident wave (1 x 4)
        ctx.set_potentials(Vm, s)
        ctx.solve_levels(Vm.min(axis=1), Vm[:, -1] - 1.0, 64, 0, 7, 16, 1e-9, 4)
    ctx.sync()
with cabi.Group([0, 0]) as g:                                          # two contexts, host threads, peer copies
    g.set_potentials(V, s, cabi.SHARD_CURVES)
    g.solve_levels(lo, hi, 256, 0, 3, 32, 1e-10, 6)
    g.set_potentials(V[0], s, cabi.SHARD_ENERGY)
    g.solve_levels(lo[0], hi[0], 257, 0, 3, 32, 1e-10, 6)
    g.sweep_uniform(lo[0], hi[0], 500)
print("sanitize_target done")
