"""Argument validation and degenerate inputs of the C ABI (include/epseon_cuda.h), called through raw
ctypes where the numpy wrapper would refuse first.  Error behaviour mirrors the reference's
conventions one level down: the C++ layer turns a non-zero status into std::runtime_error
(device_interface.hpp:40-43), the C ABI reports EPS_ERR_* + eps_last_error and never throws."""
import ctypes as C

import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu

INVALID, CUDA, RANGE, STATE = 1, 2, 3, 4


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.fixture()
def ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    c = cabi.Context(0)
    yield c
    c.close()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_null_and_size_arguments(ctx):
    from epseon_backend_b200 import cabi

    L, h = ctx.lib, ctx.h
    w = W.c1()
    V, s = np.ascontiguousarray(w["V"]), np.array([w["s"]])
    assert L.eps_set_potentials(h, None, C.c_uint32(1), C.c_uint32(V.size), _p(s)) == INVALID
    assert L.eps_set_potentials(h, _p(V), C.c_uint32(0), C.c_uint32(V.size), _p(s)) == INVALID
    assert L.eps_set_potentials(h, _p(V), C.c_uint32(1), C.c_uint32(2), _p(s)) == INVALID
    assert b"3 points" in L.eps_last_error(h)
    bad_s = np.array([0.0])
    assert L.eps_set_potentials(h, _p(V), C.c_uint32(1), C.c_uint32(V.size), _p(bad_s)) == INVALID
    assert L.eps_set_potentials(None, _p(V), C.c_uint32(1), C.c_uint32(V.size), _p(s)) == INVALID
    # nothing resident yet
    E = np.array([100.0])
    n = np.zeros(1, dtype=np.uint32)
    assert L.eps_sweep(h, _p(E), C.c_uint64(1), _p(n), None, None) == STATE
    ci = cabi.CurveInfo()
    assert L.eps_get_curve_info(h, C.c_uint32(0), C.byref(ci)) == INVALID
    # resident: bad energy arguments
    ctx.set_potentials(V, w["s"])
    assert L.eps_sweep(h, None, C.c_uint64(1), _p(n), None, None) == INVALID
    assert L.eps_sweep(h, _p(E), C.c_uint64(0), _p(n), None, None) == INVALID
    assert L.eps_sweep(h, _p(E), C.c_uint64(1 << 32), _p(n), None, None) == INVALID
    lo, hi = np.array([0.0]), np.array([1000.0])
    assert L.eps_sweep_uniform(h, _p(lo), _p(hi), C.c_uint64(0), _p(n), None, None) == INVALID
    assert L.eps_sweep_grid(h, _p(lo), _p(hi), C.c_uint32(0xFFFFFFFF), C.c_uint64(2), _p(n), None, None) == INVALID
    assert L.eps_get_curve_info(h, C.c_uint32(1), C.byref(ci)) == INVALID
    assert L.eps_get_curve_info(h, C.c_uint32(0), None) == INVALID
    for nan in (np.nan, np.inf):
        assert L.eps_sweep(h, _p(np.array([nan])), C.c_uint64(1), _p(n), None, None) == RANGE
    assert L.eps_set_option(h, C.c_int(99), C.c_int64(0)) == INVALID
    assert L.eps_get_counter(h, C.c_int(99), C.byref(C.c_uint64())) == INVALID


def test_solve_parameter_validation(ctx):
    from epseon_backend_b200 import cabi

    L, h = ctx.lib, ctx.h
    w = W.c1()
    ctx.set_potentials(w["V"], w["s"])
    lo, hi = np.array([w["E_lo"]]), np.array([w["E_hi"]])
    out = np.zeros(4)

    def solve(v_min=0, v_max=3, n_coarse=256, M=32, rounds=4, tol=1e-10, lo=lo, hi=hi, levels=out):
        p = cabi.SolveParams(v_min, v_max, n_coarse, M, rounds, 0, tol)
        return L.eps_solve_levels(h, C.byref(p), _p(lo), _p(hi), None if levels is None else _p(levels), None, None)

    assert solve() == 0
    assert solve(v_min=3, v_max=2) == INVALID
    assert solve(n_coarse=1) == INVALID
    assert solve(M=0) == INVALID
    assert solve(levels=None) == 0                              # levels may be NULL: results stay on the device (eps_mailbox_post_levels)
    assert solve(lo=hi, hi=lo) == INVALID                       # E_hi < E_lo
    assert solve(hi=np.array([1e9])) == RANGE                   # outside |s (E - V_min)| <= 0.5
    assert L.eps_solve_levels(h, None, _p(lo), _p(hi), _p(out), None, None) == INVALID


def test_degenerate_searches_match_the_oracle(oracle, ctx):
    w = W.c1()
    ctx.set_potentials(w["V"], w["s"])
    F, *_ = oracle.prep(w["V"], w["s"])
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    cases = [
        dict(E_lo=0.0, E_hi=exact[0] - 50.0, n_coarse=16, v=(0, 2)),            # no level in range: all NaN
        dict(E_lo=exact[2] + 10.0, E_hi=exact[5] - 10.0, n_coarse=33, v=(0, 6)),  # levels 3, 4 only
        dict(E_lo=0.0, E_hi=w["E_hi"], n_coarse=2, v=(0, 16)),                  # a 2-point coarse grid
        dict(E_lo=0.0, E_hi=w["E_hi"], n_coarse=300, v=(15, 20)),               # levels beyond the last bound one
        dict(E_lo=1234.5, E_hi=1234.5, n_coarse=4, v=(0, 1)),                   # empty interval
    ]
    for k in cases:
        for M, rounds in ((1, 60), (7, 3), (64, 0)):
            lev_g, wid_g, nb_g = ctx.solve_levels(k["E_lo"], k["E_hi"], k["n_coarse"], k["v"][0], k["v"][1], M, 1e-12, rounds)
            lev_o, wid_o, nb_o, *_ = oracle.solve_levels(F, w["s"], k["E_lo"], k["E_hi"], k["n_coarse"], k["v"][0], k["v"][1],
                                                         M, 1e-12, rounds)
            assert _same_bits(lev_g[0], lev_o) and _same_bits(wid_g[0], wid_o) and nb_g[0] == nb_o, (k, M, rounds)
    lev_g, _, _ = ctx.solve_levels(0.0, exact[0] - 50.0, 16, 0, 2, 8, 1e-12, 4)
    assert np.all(np.isnan(lev_g))
    lev_g, _, nb = ctx.solve_levels(0.0, w["E_hi"], 300, 15, 20, 8, 1e-12, 8)
    assert np.isfinite(lev_g[0][:2]).all() and np.isnan(lev_g[0][2:]).all() and nb[0] == 17


def test_single_energy_and_repeated_energies(oracle, ctx):
    w = W.c1()
    ctx.set_potentials(w["V"], w["s"])
    F, *_ = oracle.prep(w["V"], w["s"])
    for E in (np.array([12345.678]), np.full(100, 20000.0), np.array([w["E_hi"], 0.0, 5000.0, 0.0])):  # unsorted too
        n_g, m_g, x_g = ctx.sweep(E)
        n_o, m_o, x_o = oracle.sweep(F, w["s"], E)
        assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o)
    n_g, _, _ = ctx.sweep_uniform(5000.0, 5000.0, 1)  # one point: dE = 0/0 must not poison the energy
    n_o, _, _ = oracle.sweep(F, w["s"], np.array([5000.0]))
    assert n_g[0][0] == n_o[0]


def test_smallest_tables(oracle, ctx):
    """3-point and 4-point tables: the window rule leaves fewer than 2 steps -> EPS_ERR_RANGE, 5+ work."""
    from epseon_backend_b200 import cabi

    for n, ok in ((3, False), (4, False), (5, True), (6, True)):
        V = (np.arange(n, dtype=np.float64) - (n - 1) / 2.0) ** 2 * 1e-3
        try:
            ctx.set_potentials(V, 1.0)
            worked = True
        except cabi.EpsError as e:
            worked = False
            assert e.code == RANGE
        F = None
        try:
            F, i0, nst, vmin = oracle.prep(V, 1.0)
        except ValueError:
            pass
        assert worked == (F is not None), n
        if worked:
            ci = ctx.curve_info(0)
            assert (ci.i0, ci.n_steps) == (i0, nst)
            E = np.linspace(vmin, vmin + 0.4, 9)
            n_g, m_g, x_g = ctx.sweep(E)
            n_o, m_o, x_o = oracle.sweep(F, 1.0, E)
            assert np.array_equal(n_g[0], n_o) and _same_bits(m_g[0], m_o) and np.array_equal(x_g[0], x_o)
