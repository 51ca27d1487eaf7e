"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle.

Bar: node counts, tail mantissas/exponents and located
level energies BIT-EXACT against the oracle on the same float64 inputs (the
kernels use the same IEEE operations in the same order).  Accuracy against the
analytic Morse spectrum is asserted at the tolerance the discretisation allows.
"""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _check_sweep(oracle, ctx, V, s, E):
    ctx.set_potentials(V, s)
    ci = ctx.curve_info(0)
    AB, i0, n, vmin = oracle.prep(V, s)
    assert (ci.i0, ci.n_steps, ci.v_min) == (i0, n, vmin)
    n_g, m_g, x_g = ctx.sweep(E)
    n_o, m_o, x_o = oracle.sweep(AB, s, E)
    assert np.array_equal(n_g[0], n_o), "node counts differ"
    assert np.array_equal(x_g[0], x_o), "tail exponents differ"
    assert _same_bits(m_g[0], m_o), "tail mantissas differ"
    return n_o


def test_c1_sweep_bit_exact(oracle, gpu_ctx):
    w = W.c1()
    E = np.linspace(w["E_lo"], w["E_hi"], w["nE"])
    nodes = _check_sweep(oracle, gpu_ctx, w["V"], w["s"], E)
    assert nodes[0] == 0 and nodes[-1] == 17  # 17 bound levels of the H2-like curve


def test_c1_uniform_sweep_bit_exact(oracle, gpu_ctx):
    w = W.c1()
    gpu_ctx.set_potentials(w["V"], w["s"])
    AB, *_ = oracle.prep(w["V"], w["s"])
    nE = w["nE"]
    n_g, m_g, x_g = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE)
    dE = (w["E_hi"] - w["E_lo"]) / (nE - 1)
    n_o, m_o, x_o = oracle.sweep_uniform(AB, w["s"], w["E_lo"], dE, 0, nE)
    assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o)


@pytest.mark.parametrize("N", [4, 7, 130, 1000, 1024, 1025, 1027, 4097, 5000, 16500])
@pytest.mark.parametrize("nE", [1, 31, 257])
def test_ragged_sizes(oracle, gpu_ctx, N, nE):
    """Grid lengths around the tile (1024) and renormalisation (128) boundaries, ragged energy rows."""
    rng = np.random.default_rng(N * 1000 + nE)
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    if N < 100:  # a coarse grid needs a shallow well to stay inside the validity window
        V = V * 1e-4
    hi = min(V[-1], V.min() + 0.45 / s)
    E = np.sort(rng.uniform(V.min(), hi, nE))
    _check_sweep(oracle, gpu_ctx, V, s, E)


def test_reference_fixture_curve(oracle, gpu_ctx):
    """The reference's own fixture: MorsePotentialConfig(5500, 0.6, 10, 0, 10, 16500), Sr masses
    (python/test/test_device/test_gpu/test_libepseon_gpu.py:176-208).  V(0) ~ 9e8: the window rule
    must skip the wall."""
    N = 16500
    V = W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, N)
    s = W.scale(87.62, 87.62, W.grid_h(0.0, 10.0, N))
    E = np.linspace(0.0, 5499.9, 777)
    nodes = _check_sweep(oracle, gpu_ctx, V, s, E)
    assert gpu_ctx.curve_info(0).i0 > 1
    assert nodes[-1] == len(W.morse_levels(5500.0, 10.0, 87.62, 87.62))


def test_multi_curve_batch(oracle, gpu_ctx):
    w = W.c4(nC=9, N=3000, nE=300)
    gpu_ctx.set_potentials(w["V"], w["s"])
    n_g, m_g, x_g = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], w["nE"])
    for c in range(9):
        AB, *_ = oracle.prep(w["V"][c], w["s"])
        dE = (w["E_hi"][c] - w["E_lo"][c]) / (w["nE"] - 1)
        n_o, m_o, x_o = oracle.sweep_uniform(AB, w["s"], w["E_lo"][c], dE, 0, w["nE"])
        assert np.array_equal(n_g[c], n_o) and np.array_equal(x_g[c], x_o) and _same_bits(m_g[c], m_o)


def test_c1_levels_bit_exact_and_analytic(oracle, gpu_ctx):
    w = W.c1()
    gpu_ctx.set_potentials(w["V"], w["s"])
    AB, *_ = oracle.prep(w["V"], w["s"])
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    nlev = len(exact) + 2  # ask for two levels that do not exist
    lev_g, wid_g, nb_g = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, nlev - 1, 96, 1e-13, 12)
    lev_o, wid_o, nb_o, rounds, steps = oracle.solve_levels(AB, w["s"], w["E_lo"], w["E_hi"], 1024, 0,
                                                            nlev - 1, 96, 1e-13, 12)
    assert nb_g[0] == nb_o == len(exact)
    assert _same_bits(lev_g[0], lev_o), "level energies differ from the oracle"
    assert _same_bits(wid_g[0], wid_o)
    assert np.all(np.isnan(lev_g[0][len(exact):]))
    rel = np.abs(lev_g[0][: len(exact)] - exact) / exact
    assert rel.max() < 5e-8  # O(h^4) discretisation error at N = 10 000
    assert rel[:4].max() < 1e-9


def test_multi_curve_levels(oracle, gpu_ctx):
    w = W.c4(nC=5, N=4000, nE=256)
    gpu_ctx.set_potentials(w["V"], w["s"])
    lev_g, wid_g, nb_g = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 256, 2, 9, 64, 1e-12, 10)
    for c in range(5):
        AB, *_ = oracle.prep(w["V"][c], w["s"])
        lev_o, wid_o, nb_o, *_ = oracle.solve_levels(AB, w["s"], w["E_lo"][c], w["E_hi"][c], 256, 2, 9, 64,
                                                     1e-12, 10)
        assert nb_g[c] == nb_o
        assert _same_bits(lev_g[c], lev_o)


def test_range_error(gpu_ctx):
    from epseon_backend_b200.cabi import EpsError

    N = 500  # reference C++ fixture size (cpp/gpu/test/test_libgpu.cpp:41-48)
    V = W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, N)
    s = W.scale(87.62, 87.62, W.grid_h(0.0, 10.0, N))
    gpu_ctx.set_potentials(V, s)
    with pytest.raises(EpsError) as ei:
        gpu_ctx.sweep(np.array([1.0e9]))
    assert ei.value.code == 3


def test_stats_and_steps(gpu_ctx):
    w = W.c1()
    gpu_ctx.set_potentials(w["V"], w["s"])
    gpu_ctx.stats_reset()
    gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1000, tails=False)
    st = gpu_ctx.stats()
    assert st.sweep_launches == 1
    assert st.grid_steps == gpu_ctx.curve_info(0).n_steps * 1000
    assert st.sweep_ms > 0.0


@pytest.mark.parametrize("ept", [1, 2])
@pytest.mark.parametrize("stride", [1, 8, 32])
def test_kernel_variants_bit_exact(oracle, ept, stride, monkeypatch):
    """Every (energies-per-thread, sign-stride) kernel variant against the per-step oracle on C1
    (t_max = 9e-5: all strides are inside their validity bound)."""
    from epseon_backend_b200 import cabi

    monkeypatch.setenv("EPS_FORCE_EPT", str(ept))
    monkeypatch.setenv("EPS_FORCE_STRIDE", str(stride))
    w = W.c1()
    E = np.linspace(w["E_lo"], w["E_hi"], 1500)
    with cabi.Context(0) as ctx:
        _check_sweep(oracle, ctx, w["V"], w["s"], E)


@pytest.mark.parametrize("N,expect", [(6800, "32"), (6700, "8"), (1720, "8"), (1690, "1")])
def test_sign_stride_threshold(oracle, gpu_ctx, N, expect):
    """Grids right at the stride-selection thresholds (t_max = 2.0e-4 and 3.2e-3): the subsampled
    sign count must still equal the oracle's per-step count for a dense set of energies."""
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 10.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 10.0, N))
    t_max = s * (W.H2["De"] - 1.0)
    assert {"32": t_max <= 2.0e-4, "8": 2.0e-4 < t_max <= 3.2e-3, "1": t_max > 3.2e-3}[expect]
    rng = np.random.default_rng(N)
    E = np.sort(rng.uniform(0.0, W.H2["De"] - 1.0, 4096))
    _check_sweep(oracle, gpu_ctx, V, s, E)


def test_device_prep_rejects_bad_tables(gpu_ctx):
    """eps_set_potentials prepares on the device: non-finite values and degenerate windows still fail
    with EPS_ERR_RANGE, and a failed call leaves the context without resident curves."""
    from epseon_backend_b200.cabi import EpsError

    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, 3000)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, 3000))
    bad = V.copy()
    bad[1234] = np.nan
    with pytest.raises(EpsError) as ei:
        gpu_ctx.set_potentials(np.stack([V, bad]), s)
    assert ei.value.code == 3
    with pytest.raises(EpsError) as ei:
        gpu_ctx.sweep_uniform(0.0, 100.0, 8)
    assert ei.value.code == 4  # EPS_ERR_STATE: nothing resident
    spike = np.full(50, 1e12)
    spike[25] = 0.0  # a one-point well: the window has fewer than 2 steps
    with pytest.raises(EpsError) as ei:
        gpu_ctx.set_potentials(spike, 1.0)
    assert ei.value.code == 3


def test_device_prep_ties_and_plateaus(oracle, gpu_ctx):
    """Flat-bottomed and double-minimum tables: the first minimum wins, as in the oracle."""
    x = np.linspace(-2.0, 2.0, 5001)
    V = 4000.0 * (x * x - 1.0) ** 2  # symmetric double well: two equal minima
    V[2400:2600] = np.minimum(V[2400:2600], 3000.0)  # plateau on the barrier
    s = W.scale(15.0, 15.0, 4.0 / 5000)
    E = np.linspace(10.0, 3900.0, 300)
    _check_sweep(oracle, gpu_ctx, V, s, E)


@pytest.mark.parametrize("N", [65536, 65537, 100_000, 262_145, 1_000_000])
def test_chunked_prep_of_long_curves(oracle, gpu_ctx, N):
    """Few long curves are prepared with every curve cut into chunks over many CTAs
    (EPS_OPT_PREP_PARTS): window, table and sweeps identical to the one-CTA-per-curve kernel and to
    the oracle; first-minimum tie rule kept across chunk boundaries."""
    rmin, rmax = 0.2, 12.0
    h = W.grid_h(rmin, rmax, N)
    V = np.stack([W.morse(W.H2["De"], W.H2["re"], W.H2["a"], rmin, rmax, N),
                  W.lj(30000.0, 1.1, rmin, rmax, N)])
    # a flat bottom straddling chunk boundaries: many equal minima, the first index must win
    k = int(np.argmin(V[1]))
    V[1][k - 9000:k + 9000] = V[1][k]
    s = W.scale(W.H2["m0"], W.H2["m1"], h)
    info, nodes = [], []
    for mode in (0, 1):
        gpu_ctx.set_option(gpu_ctx.OPT_PREP_PARTS, mode)
        gpu_ctx.set_potentials(V, s)
        info.append([(ci.i0, ci.n_steps, ci.v_min, ci.v_last) for ci in (gpu_ctx.curve_info(0), gpu_ctx.curve_info(1))])
        lo = np.array([c[2] for c in info[-1]])
        n, m, x = gpu_ctx.sweep_uniform(lo, lo + 0.4 / s, 96)
        nodes.append((n, m, x))
    gpu_ctx.set_option(gpu_ctx.OPT_PREP_PARTS, 0)
    assert info[0] == info[1]
    assert np.array_equal(nodes[0][0], nodes[1][0]) and _same_bits(nodes[0][1], nodes[1][1]) and np.array_equal(nodes[0][2], nodes[1][2])
    for c in range(2):
        F, i0, nst, vmin = oracle.prep(V[c], s)
        assert info[0][c][:3] == (i0, nst, vmin)
        dE = (0.4 / s) / 95
        n_o, m_o, x_o = oracle.sweep_uniform(F, s, vmin, dE, 0, 96)
        assert np.array_equal(nodes[0][0][c], n_o) and _same_bits(nodes[0][1][c], m_o) and np.array_equal(nodes[0][2][c], x_o)


def test_chunked_prep_rejects_bad_tables(gpu_ctx):
    from epseon_backend_b200.cabi import EpsError

    N = 200_000
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    for poison in (np.nan, np.inf):
        bad = V.copy()
        bad[150_001] = poison
        with pytest.raises(EpsError) as ei:
            gpu_ctx.set_potentials(bad, s)
        assert ei.value.code == 3
    with pytest.raises(EpsError) as ei:
        gpu_ctx.set_potentials(np.full(N, np.nan), s)  # nothing finite at all
    assert ei.value.code == 3
    with pytest.raises(EpsError) as ei:
        gpu_ctx.set_potentials(np.full(3000, np.inf), s)  # same through the one-CTA kernel
    assert ei.value.code == 3
    spike = np.full(N, 1e12)
    spike[77_777] = 0.0
    with pytest.raises(EpsError) as ei:
        gpu_ctx.set_potentials(spike, 1.0)
    assert ei.value.code == 3


@pytest.mark.parametrize("nlev,M", [(8, 32), (8, 16), (3, 64), (5, 8), (2, 128), (16, 16), (8, 64)])
def test_packed_rows_in_small_ctas(oracle, gpu_ctx, nlev, M):
    """Many-curve refinement with few points per level: rows are packed into 256-energy CTAs when
    they would leave half of a 512-energy CTA idle ((8, 64) is the 512-energy control).  Levels and
    widths bit-identical to the oracle for every curve."""
    n = 3000
    rng = np.random.default_rng(nlev * 100 + M)
    V = np.stack([W.morse(5500.0 * (1 + 0.04 * rng.random()), 2.2, 1.6, 0.4, 10.0, n) if c % 2 == 0
                  else W.lj(5200.0 * (1 + 0.04 * rng.random()), 2.3, 0.4, 10.0, n) for c in range(5)])
    s = W.scale(20.0, 20.0, W.grid_h(0.4, 10.0, n))
    gpu_ctx.set_potentials(V, s)
    lo, hi = V.min(axis=1), V[:, -1] - 1.0
    lev_g, wid_g, nb_g = gpu_ctx.solve_levels(lo, hi, 700, 1, nlev, M, 1e-11, 12)
    for c in range(5):
        F, *_ = oracle.prep(V[c], s)
        lev_o, wid_o, nb_o, *_ = oracle.solve_levels(F, s, lo[c], hi[c], 700, 1, nlev, M, 1e-11, 12)
        assert _same_bits(lev_g[c], lev_o) and _same_bits(wid_g[c], wid_o) and nb_g[c] == nb_o, (c, nlev, M)
