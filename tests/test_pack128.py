"""128-energy packed CTAs (EPS_OPT_PACK128): refinement rows of many-curve batches in CTAs of one chain x
four warps, four resident per SM, chosen automatically when they balance the SMs better than the
256-energy ones (small per-device batches).  Same bits as every other shape, in both recurrences."""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.mark.parametrize("form", [0, 1])
@pytest.mark.parametrize("M,vmax", [(16, 7), (32, 7), (64, 2), (32, 10), (20, 4)])
def test_pack128_equals_oracle_and_other_shapes(gpu_ctx, oracle, oracle_d, form, M, vmax):
    orc = oracle_d if form else oracle
    ctx = gpu_ctx
    w = W.c4(nC=13, N=3000, nE=300)
    ctx.set_option(ctx.OPT_FORM, form)
    try:
        ctx.set_potentials(w["V"], w["s"])
        out = {}
        for opt in (2, 1):  # never / always
            ctx.set_option(ctx.OPT_PACK128, opt)
            out[opt] = ctx.solve_levels(w["E_lo"], w["E_hi"], 300, 0, vmax, M, 1e-11, 12)
        assert _same_bits(out[1][0], out[2][0]) and _same_bits(out[1][1], out[2][1]) and np.array_equal(out[1][2], out[2][2])
        for c in (0, 5, 12):
            T, *_ = orc.prep(w["V"][c], w["s"])
            lev_o, wid_o, nb_o, *_ = orc.solve_levels(T, w["s"], w["E_lo"][c], w["E_hi"][c], 300, 0, vmax, M, 1e-11, 12)
            assert _same_bits(out[1][0][c], lev_o) and _same_bits(out[1][1][c], wid_o) and out[1][2][c] == nb_o
    finally:
        ctx.set_option(ctx.OPT_PACK128, 0)
        ctx.set_option(ctx.OPT_FORM, 0)
