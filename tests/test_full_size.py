"""BASELINE.json's configurations at (or near) their full sizes, checked through size-independent
properties -- the oracle would need minutes of CPU there:

* node counts are monotone non-decreasing in the trial energy and jump exactly at the located levels
  (Sturm count), the last count equals the number of bound levels;
* the three routes to the same numbers -- TMA ring kernel, constant-bank kernel, transfer-matrix
  scan -- return identical node counts, and slices of a global energy grid concatenate to the whole;
* two runs give identical bits (determinism checksum);
* a seeded sample of curves / energies is compared with the oracle bit for bit.
"""
import numpy as np
import pytest

from epseon_backend_b200 import multi
from tests import workloads as W

pytestmark = pytest.mark.gpu


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


def test_c2_full_size(oracle, ctx):
    """100k-point grid, 65 536 coarse energies, all 17 levels to 1e-10 (the headline workload)."""
    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    nodes, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], tails=False)
    nodes = nodes[0].astype(np.int64)
    assert np.all(np.diff(nodes) >= 0) and np.all(np.diff(nodes) <= 1)
    assert nodes[0] == 0 and nodes[-1] == 17
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], w["nE"], 0, 16, 4457, 1e-10, 8)
    lev2, _, _ = ctx.solve_levels(w["E_lo"], w["E_hi"], w["nE"], 0, 16, 4457, 1e-10, 8)
    assert _same_bits(lev, lev2) and nb[0] == 17
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    assert np.max(np.abs(lev[0] - exact) / exact) < 5e-8
    assert np.all(wid[0] <= 1e-10 * np.abs(lev[0]))
    # the count jumps exactly at the levels: #grid energies with nodes <= v  ==  #grid energies below E_v
    E = w["E_lo"] + np.arange(w["nE"]) * ((w["E_hi"] - w["E_lo"]) / (w["nE"] - 1))
    for v in range(17):
        assert np.count_nonzero(nodes <= v) == np.searchsorted(E, lev[0][v]), v
    # constant-bank kernel: same counts; a seeded sample of energies against the oracle
    ctx.set_option(ctx.OPT_CBANK, 1)
    nodes_cb, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], tails=False)
    ctx.set_option(ctx.OPT_CBANK, 0)
    assert np.array_equal(nodes_cb[0], nodes)
    F, *_ = oracle.prep(w["V"], w["s"])
    pick = np.sort(np.random.default_rng(5).choice(w["nE"], 64, replace=False))
    n_o, m_o, x_o = oracle.sweep(F, w["s"], E[pick])
    n_g, m_g, x_g = ctx.sweep(E[pick])
    assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o)
    assert np.array_equal(nodes[pick], n_o)


def test_c5_quarter_size_routes_and_slices(ctx):
    """200k-point grid, 2^22 energies over C5's range (a quarter of its 2^24): TMA == constant bank,
    8 slices of the global grid == the whole sweep, counts monotone."""
    nE = 1 << 22
    w = W.c5(nE=nE)
    ctx.set_potentials(w["V"], w["s"])
    ctx.set_option(ctx.OPT_CBANK, 2)
    tma, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, tails=False)
    ctx.set_option(ctx.OPT_CBANK, 1)
    cb, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, tails=False)
    ctx.set_option(ctx.OPT_CBANK, 0)
    assert np.array_equal(tma, cb)
    d = np.diff(tma[0].astype(np.int64))
    assert np.all(d >= 0) and tma[0][0] == 0 and tma[0][-1] == 17
    dE = multi.global_step(w["E_lo"], w["E_hi"], nE)
    parts = []
    for r in range(8):
        j0, n = multi.energy_shard(nE, 8, r)
        p, _, _ = ctx.sweep_grid(w["E_lo"], dE, j0, n, tails=False)
        parts.append(p[0][: n - 1] if r < 7 else p[0])  # neighbours share one point
    assert np.array_equal(np.concatenate(parts), tma[0])
    assert int(tma[0].astype(np.uint64).sum()) == int(cb[0].astype(np.uint64).sum())  # checksum of the run


def test_c5_full_size_sampled_against_oracle(oracle, ctx):
    """BASELINE configs[4] at its FULL size: 2^24 trial energies on the 200k-point grid (the auto
    policy selects the constant-bank kernel).  Node counts monotone from 0 to 17, a level-count jump
    of exactly one at each of the 17 analytic levels, and a seeded sample of 4096 of the 2^24
    energies -- plus the two grid points around every jump -- bit-identical to the oracle."""
    nE = 1 << 24
    w = W.c5(nE=nE)
    ctx.set_potentials(w["V"], w["s"])
    before = ctx.counter(ctx.CNT_CBANK_LAUNCHES)
    nodes, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, tails=False)
    assert ctx.counter(ctx.CNT_CBANK_LAUNCHES) > before
    n = nodes[0]
    d = np.diff(n.astype(np.int64))
    assert n[0] == 0 and n[-1] == 17 and np.all(d >= 0) and np.all(d <= 1)
    jumps = np.flatnonzero(d)  # last grid index below each level
    assert jumps.size == 17
    dE = multi.global_step(w["E_lo"], w["E_hi"], nE)
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    located = w["E_lo"] + (jumps + 0.5) * dE
    assert np.max(np.abs(located - exact) / exact) < 5e-8 + float(dE) / exact[0]
    pick = np.unique(np.concatenate([np.random.default_rng(24).integers(0, nE, 4096), jumps, jumps + 1, [0, nE - 1]]))
    E = w["E_lo"] + pick.astype(np.float64) * dE  # the device's operations: one multiplication, one addition
    F, *_ = oracle.prep(w["V"], w["s"])
    n_o, _, _ = oracle.sweep(F, w["s"], E, tails=False)
    assert np.array_equal(n[pick], n_o)


def test_c3_full_size_scan_equals_sequential(ctx):
    """1M-point tabulated curve x 4096 energies: transfer-matrix scan path == sequential march."""
    w = W.c3()
    ctx.set_potentials(w["V"], w["s"])
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 1)  # never
    seq, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], tails=False)
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)  # automatic: this shape selects the scan path
    before = ctx.counter(ctx.CNT_SCAN_LAUNCHES)
    scan, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], tails=False)
    assert ctx.counter(ctx.CNT_SCAN_LAUNCHES) > before
    assert np.array_equal(seq, scan)
    assert np.all(np.diff(seq[0].astype(np.int64)) >= 0) and seq[0][0] == 0


def test_c4_full_batch(oracle, ctx):
    """4096 perturbed Morse / LJ curves x 1024 coarse energies, levels 0..7: every level located and
    ordered, a seeded sample of curves bit-identical to the oracle, two runs identical."""
    w = W.c4()
    ctx.set_potentials(w["V"], w["s"])
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 64, 1e-10, 8)
    lev2, _, _ = ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 7, 64, 1e-10, 8)
    assert _same_bits(lev, lev2)
    assert lev.shape == (4096, 8) and np.all(np.isfinite(lev)) and np.all(np.diff(lev, axis=1) > 0)
    assert np.all(nb >= 8)
    for c in np.random.default_rng(3).choice(4096, 12, replace=False):
        F, *_ = oracle.prep(w["V"][c], w["s"])
        lev_o, *_ = oracle.solve_levels(F, w["s"], w["E_lo"][c], w["E_hi"][c], 1024, 0, 7, 64, 1e-10, 8)
        assert _same_bits(lev[c], lev_o), c
