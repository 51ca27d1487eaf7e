"""The accurate recurrence (EPS_OPT_FORM = 1, "D form") through the C ABI against
the oracle's statement of the same form: node counts, tails and levels BIT-EXACT on every route
(TMA ring kernel in both CTA shapes, packed and flat refinement rows, constant-bank kernel,
transfer-matrix scan), and agreement with the X form to the X form's noise floor."""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.fixture(scope="module")
def orc_d():
    from oracle import Oracle

    return Oracle(form=1)


@pytest.fixture()
def ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    c = cabi.Context(0)
    c.set_option(c.OPT_FORM, 1)
    yield c
    c.close()


def _check_sweep(orc_d, ctx, V, s, E):
    ctx.set_potentials(V, s)
    A, i0, n, vmin = orc_d.prep(V, s)
    ci = ctx.curve_info(0)
    assert (ci.i0, ci.n_steps, ci.v_min) == (i0, n, vmin)
    n_g, m_g, x_g = ctx.sweep(E)
    n_o, m_o, x_o = orc_d.sweep(A, s, E)
    assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o)
    return n_o


def test_c1_sweep_and_levels_bit_exact(orc_d, ctx):
    w = W.c1()
    E = np.linspace(w["E_lo"], w["E_hi"], w["nE"])
    nodes = _check_sweep(orc_d, ctx, w["V"], w["s"], E)
    assert nodes[0] == 0 and nodes[-1] == 17
    A, *_ = orc_d.prep(w["V"], w["s"])
    for n_coarse, M in ((512, 64), (1024, 256), (4096, 1000)):
        lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], n_coarse, 0, 16, M, 1e-13, 12)
        lev_o, wid_o, nb_o, *_ = orc_d.solve_levels(A, w["s"], w["E_lo"], w["E_hi"], n_coarse, 0, 16, M, 1e-13, 12)
        assert _same_bits(lev[0], lev_o) and _same_bits(wid[0], wid_o) and nb[0] == nb_o == 17


@pytest.mark.parametrize("N", [4, 130, 1025, 2049, 4097, 16500])
@pytest.mark.parametrize("nE", [1, 31, 257, 600])
def test_ragged_sizes(orc_d, ctx, N, nE):
    rng = np.random.default_rng(N * 1000 + nE)
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    if N < 100:
        V = V * 1e-4
    hi = min(V[-1], V.min() + 0.45 / s)
    E = np.sort(rng.uniform(V.min(), hi, nE))
    _check_sweep(orc_d, ctx, V, s, E)


@pytest.mark.parametrize("stride_case", ["coarse grid (per-step count)", "stride 8", "stride 32"])
def test_sign_strides(orc_d, ctx, stride_case):
    N = {"coarse grid (per-step count)": 600, "stride 8": 6000, "stride 32": 60000}[stride_case]
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, N))
    E = np.linspace(V.min() + 1.0, V[-1] - 1.0, 700)
    _check_sweep(orc_d, ctx, V, s, E)


def test_multi_curve_batch_packed_rows(orc_d, ctx):
    """Several curves: dense packed refinement rows in 512- and 256-energy CTAs."""
    w = W.c4(nC=24, N=3000, nE=300)
    ctx.set_potentials(w["V"], w["s"])
    for M in (16, 32, 64, 200):
        lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 300, 0, 7, M, 1e-11, 12)
        for c in range(24):
            A, *_ = orc_d.prep(w["V"][c], w["s"])
            lev_o, wid_o, nb_o, *_ = orc_d.solve_levels(A, w["s"], w["E_lo"][c], w["E_hi"][c], 300, 0, 7, M, 1e-11, 12)
            assert _same_bits(lev[c], lev_o) and _same_bits(wid[c], wid_o) and nb[c] == nb_o, (M, c)


def test_constant_bank_route(orc_d, ctx):
    w = W.c2(N=20_000)
    ctx.set_potentials(w["V"], w["s"])
    A, *_ = orc_d.prep(w["V"], w["s"])
    nE = 3000
    dE = (w["E_hi"] - w["E_lo"]) / (nE - 1)
    n_o, m_o, x_o = orc_d.sweep_uniform(A, w["s"], w["E_lo"], dE, 0, nE)
    for opt in (2, 1):  # never / always
        ctx.set_option(ctx.OPT_CBANK, opt)
        before = ctx.counter(ctx.CNT_CBANK_LAUNCHES)
        n_g, m_g, x_g = ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE)
        assert (ctx.counter(ctx.CNT_CBANK_LAUNCHES) > before) == (opt == 1)
        assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o)


def test_scan_route_nodes_bit_exact(orc_d, ctx):
    """Transfer matrices in the D form's own (Y, D) coordinates: node counts equal the sequential
    march's (flagged energies recomputed), tails agree to rounding."""
    N = 150_000
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 12.0, N)
    s = W.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 12.0, N))
    A, *_ = orc_d.prep(V, s)
    ctx.set_potentials(V, s)
    nE = 900
    E_lo, E_hi = 0.0, W.H2["De"] - 1.0
    dE = (E_hi - E_lo) / (nE - 1)
    n_o, m_o, x_o = orc_d.sweep_uniform(A, s, E_lo, dE, 0, nE)
    for seg in (2, 7, 19):
        ctx.set_option(ctx.OPT_SCAN_SEGMENTS, seg)
        before = ctx.counter(ctx.CNT_SCAN_LAUNCHES)
        n_g, m_g, x_g = ctx.sweep_uniform(E_lo, E_hi, nE)
        assert ctx.counter(ctx.CNT_SCAN_LAUNCHES) > before
        assert np.array_equal(n_g[0], n_o), seg
        ok = np.abs(m_o) > 0
        rel = np.abs(m_g[0][ok] * 2.0 ** (x_g[0][ok] - x_o[ok]).astype(np.float64) - m_o[ok]) / np.abs(m_o[ok])
        assert np.median(rel) < 1e-9
    ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
    lev, _, nb = ctx.solve_levels(E_lo, E_hi, 2048, 0, 16, 256, 1e-12, 10)  # auto: the scan path with the fix-up
    lev_o, _, nb_o, *_ = orc_d.solve_levels(A, s, E_lo, E_hi, 2048, 0, 16, 256, 1e-12, 10)
    assert _same_bits(lev[0], lev_o) and nb[0] == nb_o


def test_forms_agree_to_the_x_form_noise_floor(ctx, oracle):
    """Same curve, both recurrences: identical level counts, energies within the X form's floor."""
    w = W.c2(N=50_000)
    ctx.set_potentials(w["V"], w["s"])
    lev_d, _, nb_d = ctx.solve_levels(w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-13, 12)
    ctx.set_option(ctx.OPT_FORM, 0)
    ctx.set_potentials(w["V"], w["s"])
    lev_x, _, nb_x = ctx.solve_levels(w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-13, 12)
    assert nb_d[0] == nb_x[0] == 17
    rel = np.abs(lev_d[0] - lev_x[0]) / lev_x[0]
    assert rel.max() < 2e-9 and rel.max() > 1e-14  # they do differ: the X form carries rounding noise


def test_rotational_states_d_form(orc_d, ctx):
    N = 4000
    V = W.morse(5500.0, 2.2, 1.6, 0.5, 9.0, N)
    h = W.grid_h(0.5, 9.0, N)
    s = W.scale(20.0, 20.0, h)
    J = [0, 3, 11]
    ctx.set_potentials_rot(V, s, 0.5, h, J)
    for j, Jv in enumerate(J):
        VJ = orc_d.centrifugal(V, s, 0.5, h, Jv)
        A, *_ = orc_d.prep(VJ, s)
        ci = ctx.curve_info(j)
        lev, _, nb = ctx.solve_levels(np.full(3, ci.v_min), np.full(3, ci.v_last - 1.0), 512, 0, 5, 64, 1e-12, 10)
        lev_o, _, nb_o, *_ = orc_d.solve_levels(A, s, ci.v_min, ci.v_last - 1.0, 512, 0, 5, 64, 1e-12, 10)
        assert _same_bits(lev[j], lev_o) and nb[j] == nb_o
