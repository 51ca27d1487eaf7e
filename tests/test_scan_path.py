"""N4: the transfer-matrix (scan) path of the sweep against the sequential CPU oracle.

Contract (include/epseon_cuda.h EPS_OPT_SCAN_*):
  * node counts equal the oracle's, bit for bit (ill-conditioned energies are detected on the
    device and recomputed with the sequential kernel);
  * tails agree to rounding (median relative difference < 1e-10; larger only where the tail is
    itself the result of a cancellation, i.e. E close to an eigenvalue);
  * levels located through the scan path agree with the oracle's to <= 1e-9 relative, and are
    bit-identical with EPS_OPT_SCAN_EXACT = 1.
"""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture()
def scan_ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    ctx = cabi.Context(0)
    yield ctx
    ctx.close()


def _tail_ratio(m_g, x_g, m_o, x_o):
    return (m_g / m_o) * np.exp2((x_g - x_o).astype(np.float64))


@pytest.mark.parametrize("combine", [1, 2], ids=["serial", "prefix"])
@pytest.mark.parametrize("n_seg", [2, 3, 5])
def test_forced_scan_c1(oracle, scan_ctx, n_seg, combine):
    w = W.c1()
    scan_ctx.set_potentials(w["V"], w["s"])
    scan_ctx.set_option(scan_ctx.OPT_SCAN_SEGMENTS, n_seg)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_COMBINE, combine)
    F, *_ = oracle.prep(w["V"], w["s"])
    E = np.linspace(w["E_lo"], w["E_hi"], 1500)
    n_g, m_g, x_g = scan_ctx.sweep(E)
    n_o, m_o, x_o = oracle.sweep(F, w["s"], E)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1
    assert np.array_equal(n_g[0], n_o)
    r = _tail_ratio(m_g[0], x_g[0], m_o, x_o)
    # relative agreement degrades only where the tail itself is a cancellation (E close to a level)
    assert np.median(np.abs(r - 1.0)) < 1e-10 and np.abs(r - 1.0).max() < 1e-5


@pytest.mark.parametrize("combine", [1, 2], ids=["serial", "prefix"])
@pytest.mark.parametrize("N", [2300, 4097, 6200, 16500])
@pytest.mark.parametrize("n_seg,nE", [(2, 31), (4, 257), (64, 700)])
def test_forced_scan_ragged(oracle, scan_ctx, N, n_seg, nE, combine):
    """Segment counts above the tile count (clamped), ragged last tiles, ragged energy rows."""
    rng = np.random.default_rng(N + n_seg)
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    scan_ctx.set_potentials(V, s)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_SEGMENTS, n_seg)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_COMBINE, combine)
    F, *_ = oracle.prep(V, s)
    E = np.sort(rng.uniform(V.min(), min(V[-1], V.min() + 0.45 / s), nE))
    n_g, _, _ = scan_ctx.sweep(E, tails=False)
    n_o, _, _ = oracle.sweep(F, s, E, tails=False)
    assert np.array_equal(n_g[0], n_o)


def test_forced_scan_multi_curve_uniform(oracle, scan_ctx):
    w = W.c4(nC=7, N=9000, nE=300)
    scan_ctx.set_potentials(w["V"], w["s"])
    scan_ctx.set_option(scan_ctx.OPT_SCAN_SEGMENTS, 3)
    n_g, m_g, x_g = scan_ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"])
    for c in range(7):
        F, *_ = oracle.prep(w["V"][c], w["s"])
        dE = (w["E_hi"][c] - w["E_lo"][c]) / (w["nE"] - 1)
        n_o, m_o, x_o = oracle.sweep_uniform(F, w["s"], w["E_lo"][c], dE, 0, w["nE"])
        assert np.array_equal(n_g[c], n_o)


@pytest.mark.parametrize("combine", [1, 2], ids=["serial", "prefix"])
def test_flagged_energies_fall_back_to_sequential(oracle, scan_ctx, combine):
    """Energies sitting on eigenvalues to 1e-14.  For the upper levels a segment boundary (every
    2048 steps = 2 Angstrom here) lies inside the classically allowed region, so the decaying tail is
    a pure cancellation of the segment's two columns: the device flags those energies and the
    sequential kernel supplies the oracle's bits (nodes AND tails).  (Low levels are not flagged:
    their whole allowed region lies in segment 0, which the scan marches exactly like the
    sequential kernel.)"""
    w = W.c1()
    scan_ctx.set_potentials(w["V"], w["s"])
    scan_ctx.set_option(scan_ctx.OPT_SCAN_SEGMENTS, 5)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_COMBINE, combine)
    F, *_ = oracle.prep(w["V"], w["s"])
    lev, wid, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 96, 1e-15, 14)
    E = np.sort(np.concatenate([lev, lev - wid, lev + wid, np.linspace(100.0, 38000.0, 50)]))
    n_g, m_g, x_g = scan_ctx.sweep(E)
    n_o, m_o, x_o = oracle.sweep(F, w["s"], E)
    flagged = scan_ctx.counter(scan_ctx.CNT_SCAN_FLAGGED)
    assert flagged >= 3
    assert np.array_equal(n_g[0], n_o)
    top = np.isin(E, np.concatenate([lev[-2:], lev[-2:] - wid[-2:], lev[-2:] + wid[-2:]]))
    assert np.array_equal(m_g[0][top].view(np.uint64), m_o[top].view(np.uint64))
    assert np.array_equal(x_g[0][top], x_o[top])


def test_auto_scan_long_grid(oracle, scan_ctx):
    """Few energies on a long grid: the scan path is selected automatically."""
    N, nE = 150_000, 1024
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 12.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 12.0, N))
    scan_ctx.set_potentials(V, s)
    F, *_ = oracle.prep(V, s)
    n_g, _, _ = scan_ctx.sweep_uniform(0.0, W.H2["De"] - 1.0, nE, tails=False)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1
    n_o, _, _ = oracle.sweep_uniform(F, s, 0.0, (W.H2["De"] - 1.0) / (nE - 1), 0, nE, tails=False)
    assert np.array_equal(n_g[0], n_o)
    # with tails requested the sequential kernel runs (bit-exact tails)
    n_g, m_g, x_g = scan_ctx.sweep_uniform(0.0, W.H2["De"] - 1.0, 64)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1
    # many energies: sequential
    scan_ctx.sweep_uniform(0.0, W.H2["De"] - 1.0, 148 * 512, tails=False)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1


def test_levels_through_scan_path(oracle, scan_ctx):
    N = 150_000
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 12.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 12.0, N))
    scan_ctx.set_potentials(V, s)
    F, *_ = oracle.prep(V, s)
    args = (0.0, W.H2["De"] - 1.0, 1024, 0, 16, 128)
    lev_o, wid_o, nb_o, *_ = oracle.solve_levels(F, s, *args, 1e-10, 8)
    # default (EPS_OPT_SCAN_EXACT = 1): flagged energies are recomputed sequentially -> oracle's bits
    lev_x, wid_x, nb_x = scan_ctx.solve_levels(*args, 1e-10, 8)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1  # the coarse sweep; exact-mode refinement rounds march sequentially
    assert nb_x[0] == nb_o == 17
    assert np.array_equal(lev_x[0].view(np.uint64), lev_o.view(np.uint64))
    assert np.array_equal(wid_x[0].view(np.uint64), wid_o.view(np.uint64))
    # fast mode: no sequential fix-up; the scan's own rounding moves a level by at most the noise
    # floor of the FP64 recurrence (~2e-9 relative at N ~ 1e5)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_EXACT, 0)
    lev_g, wid_g, nb_g = scan_ctx.solve_levels(*args, 1e-10, 8)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) >= 3  # ... in the fast mode they take the scan path too
    assert nb_g[0] == 17
    assert np.abs(lev_g[0] / lev_o - 1.0).max() <= 5e-9


def test_c3_reduced_tabulated_curve(oracle, scan_ctx):
    """Reduced C3 (spline-resampled 'ab initio' table, 262 144 points, 1024 energies)."""
    w = W.c3(N=262_144, nE=1024)
    scan_ctx.set_potentials(w["V"], w["s"])
    F, *_ = oracle.prep(w["V"], w["s"])
    n_g, _, _ = scan_ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"], tails=False)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1
    dE = (w["E_hi"] - w["E_lo"]) / (w["nE"] - 1)
    n_o, _, _ = oracle.sweep_uniform(F, w["s"], w["E_lo"], dE, 0, w["nE"], tails=False)
    assert np.array_equal(n_g[0], n_o)
    assert n_o[-1] > n_o[0]


@pytest.mark.parametrize("n_seg", [8, 18, 37, 73])
def test_prefix_combine_many_segments(oracle, scan_ctx, n_seg):
    """The block-level parallel prefix over the 2x2 segment matrices (EPS_OPT_SCAN_COMBINE): node counts
    of the sequential oracle bit for bit, the same counts and (to rounding) tails as the serial combine,
    for segment counts that give the 16 lanes 1 .. 5 segments each (73 = one 2048-step tile per segment)."""
    N, nE = 150_000, 777
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 12.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 12.0, N))
    scan_ctx.set_potentials(V, s)
    scan_ctx.set_option(scan_ctx.OPT_SCAN_SEGMENTS, n_seg)
    F, *_ = oracle.prep(V, s)
    rng = np.random.default_rng(n_seg)
    E = np.sort(rng.uniform(0.0, W.H2["De"] - 1.0, nE))
    n_o, m_o, x_o = oracle.sweep(F, s, E)
    out = {}
    for combine in (1, 2):
        scan_ctx.set_option(scan_ctx.OPT_SCAN_COMBINE, combine)
        before = scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES)
        out[combine] = scan_ctx.sweep(E)
        assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == before + 1
        assert np.array_equal(out[combine][0][0], n_o)
        r = _tail_ratio(out[combine][1][0], out[combine][2][0], m_o, x_o)
        assert np.median(np.abs(r - 1.0)) < 1e-9 and np.abs(r - 1.0).max() < 1e-4
    r12 = _tail_ratio(out[1][1][0], out[1][2][0], out[2][1][0], out[2][2][0])
    assert np.median(np.abs(r12 - 1.0)) < 1e-11


def test_few_energies_long_grid_uses_many_segments(oracle, scan_ctx):
    """256 energies on a 300k grid: the automatic policy cuts the grid into one segment per tile (146,
    the serial combine stopped at 64) and chains them with the prefix kernel; energies placed ON
    eigenvalues are flagged and recomputed, so the counts are the oracle's."""
    N, nE = 300_000, 256
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 12.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 12.0, N))
    scan_ctx.set_potentials(V, s)
    F, *_ = oracle.prep(V, s)
    lev, wid, *_ = oracle.solve_levels(F, s, 0.0, W.H2["De"] - 1.0, 1024, 0, 16, 96, 1e-15, 14)
    E = np.sort(np.concatenate([lev, np.linspace(50.0, W.H2["De"] - 2.0, nE - lev.size)]))
    n_g, _, _ = scan_ctx.sweep(E, tails=False)
    assert scan_ctx.counter(scan_ctx.CNT_SCAN_LAUNCHES) == 1
    n_o, _, _ = oracle.sweep(F, s, E, tails=False)
    assert np.array_equal(n_g[0], n_o)
