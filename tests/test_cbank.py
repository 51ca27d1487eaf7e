"""Constant-bank sweep kernel (EPS_OPT_CBANK): same bits as the oracle, for every CTA shape, sign
stride, ragged chunk boundary (3968-step chunks, 128-step renormalisation blocks) and row layout."""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cb_ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    ctx = cabi.Context(0)
    ctx.set_option(ctx.OPT_CBANK, 1)
    yield ctx
    ctx.close()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.mark.parametrize("shape", [2128, 2256, 4128, 4256])
def test_cbank_c1_all_shapes(oracle, cb_ctx, shape):
    w = W.c1()
    cb_ctx.set_option(cb_ctx.OPT_CBANK_SHAPE, shape)
    cb_ctx.set_potentials(w["V"], w["s"])
    F, *_ = oracle.prep(w["V"], w["s"])
    E = np.linspace(w["E_lo"], w["E_hi"], 1500)
    n_g, m_g, x_g = cb_ctx.sweep(E)
    assert cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES) == 3  # 9799 steps = 3 chunks
    n_o, m_o, x_o = oracle.sweep(F, w["s"], E)
    assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and np.array_equal(_bits(m_g[0]), _bits(m_o))


@pytest.mark.parametrize("N", [130, 3969, 3970, 4097, 7937, 7940, 8100, 16500])
@pytest.mark.parametrize("nE", [1, 257, 1030])
def test_cbank_ragged(oracle, cb_ctx, N, nE):
    """n_steps around multiples of the 3968-step chunk; t_max spans the three sign strides."""
    rng = np.random.default_rng(N + nE)
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    cb_ctx.set_potentials(V, s)
    F, *_ = oracle.prep(V, s)
    E = np.sort(rng.uniform(V.min(), min(V[-1], V.min() + 0.45 / s), nE))
    n_g, m_g, x_g = cb_ctx.sweep(E)
    n_o, m_o, x_o = oracle.sweep(F, s, E)
    assert cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES) >= 1
    assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and np.array_equal(_bits(m_g[0]), _bits(m_o))


def test_cbank_levels(oracle, cb_ctx):
    """Multi-row launches (one row per level) through the constant-bank kernel."""
    w = W.c1()
    cb_ctx.set_potentials(w["V"], w["s"])
    F, *_ = oracle.prep(w["V"], w["s"])
    lev_g, wid_g, nb = cb_ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 16, 600, 1e-13, 12)
    lev_o, wid_o, nb_o, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 600, 1e-13, 12)
    assert cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES) > 3
    assert nb[0] == nb_o == 17
    assert np.array_equal(_bits(lev_g[0]), _bits(lev_o)) and np.array_equal(_bits(wid_g[0]), _bits(wid_o))


def test_cbank_not_used_for_batches(cb_ctx):
    w = W.c4(nC=3, N=3000, nE=64)
    cb_ctx.set_potentials(w["V"], w["s"])
    cb_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 64, tails=False)
    assert cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES) == 0


@pytest.mark.parametrize("form", [0, 1])
def test_cbank_energy_groups_same_bits(oracle, oracle_d, cb_ctx, form):
    """EPS_OPT_CBANK_GROUP: the chunk launches of a sweep run group by group of resident waves (the carried
    state stays in L2).  Same node counts and tails whatever the group size -- groups of one wave put the
    second group at CTA 1184 -- and the oracle's bits on a sample."""
    N, nE = 4200, 1184 * 512 + 901  # two chunks; more than one wave of 512-energy CTAs, ragged last CTA
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    orc = oracle_d if form else oracle
    cb_ctx.set_option(cb_ctx.OPT_FORM, form)
    try:
        cb_ctx.set_potentials(V, s)
        lo, hi = float(V.min()), float(min(V[-1], V.min() + 0.4 / s))
        out = {}
        for group in (0, 1, 2):
            cb_ctx.set_option(cb_ctx.OPT_CBANK_GROUP, group)
            before = cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES)
            out[group] = cb_ctx.sweep_uniform(lo, hi, nE)
            chunks = -(-cb_ctx.curve_info(0).n_steps // 3968)
            assert cb_ctx.counter(cb_ctx.CNT_CBANK_LAUNCHES) - before == chunks * (2 if group == 1 else 1)
        for group in (1, 2):
            assert np.array_equal(out[group][0], out[0][0]) and np.array_equal(out[group][2], out[0][2])
            assert np.array_equal(_bits(out[group][1]), _bits(out[0][1]))
        T, *_ = orc.prep(V, s)
        idx = np.unique(np.concatenate([[0, nE - 1, 1184 * 512 - 1, 1184 * 512], np.random.default_rng(3).integers(0, nE, 2000)]))
        dE = (hi - lo) / (nE - 1)
        E = lo + idx.astype(np.float64) * dE
        n_o, m_o, x_o = orc.sweep(T, s, E)
        assert np.array_equal(out[1][0][0][idx], n_o) and np.array_equal(_bits(out[1][1][0][idx]), _bits(m_o))
    finally:
        cb_ctx.set_option(cb_ctx.OPT_CBANK_GROUP, 2)
        cb_ctx.set_option(cb_ctx.OPT_FORM, 0)
