"""Seeded randomised parity: random potentials (Morse, Lennard-Jones, harmonic, double well, rotating
Morse), masses, grid sizes and trial energies -- every sweep and level search through the C ABI must
reproduce the oracle bit for bit.  Catches what hand-picked cases miss: sign-stride selection at
arbitrary t_max, windows cut by walls on either side, rows that end inside a renormalisation block."""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _random_case(rng):
    kind = rng.choice(["morse", "lj", "harmonic", "double", "rot"])
    N = int(rng.choice([rng.integers(40, 300), rng.integers(300, 3000), rng.integers(3000, 9000)]))
    m0, m1 = rng.uniform(1.0, 90.0, 2)
    if kind in ("morse", "rot"):
        De, re, a = rng.uniform(300.0, 40000.0), rng.uniform(0.7, 3.0), rng.uniform(0.8, 3.0)
        rmin, rmax = re * rng.uniform(0.2, 0.6), re + rng.uniform(3.0, 12.0)
        V = W.morse(De, re, a, rmin, rmax, N)
        if kind == "rot":
            r = rmin + np.arange(N) * W.grid_h(rmin, rmax, N)
            J = int(rng.integers(1, 60))
            V = V + J * (J + 1) * (W.HBAR2_OVER_2 * (m0 + m1) / (m0 * m1)) / (r * r)
    elif kind == "lj":
        De, re = rng.uniform(50.0, 6000.0), rng.uniform(2.0, 4.5)
        rmin, rmax = re * rng.uniform(0.55, 0.8), re + rng.uniform(4.0, 15.0)
        V = W.lj(De, re, rmin, rmax, N)
    elif kind == "harmonic":
        rmin, rmax = -rng.uniform(1.0, 3.0), rng.uniform(1.0, 3.0)
        x = rmin + np.arange(N) * W.grid_h(rmin, rmax, N)
        V = rng.uniform(200.0, 30000.0) * x * x
    else:
        rmin, rmax = -2.2, 2.2
        x = rmin + np.arange(N) * W.grid_h(rmin, rmax, N)
        V = rng.uniform(500.0, 8000.0) * (x * x - 1.0) ** 2 + rng.uniform(-300.0, 300.0) * x
    s = W.scale(m0, m1, W.grid_h(rmin, rmax, N))
    return kind, np.ascontiguousarray(V), s


@pytest.mark.parametrize("seed", range(6))
def test_random_sweeps_bit_exact(oracle, gpu_ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    done = 0
    while done < 25:
        kind, V, s = _random_case(rng)
        try:
            F, i0, n, vmin = oracle.prep(V, s)
        except ValueError:
            continue  # window shorter than two steps: both sides refuse (covered in test_gpu_parity)
        span = 0.49 / s  # validity window |s (E - V_min)| <= 0.5
        top = min(vmin + span, float(V.max()))
        nE = int(rng.choice([1, rng.integers(2, 40), rng.integers(40, 700)]))
        gpu_ctx.set_potentials(V, s)
        ci = gpu_ctx.curve_info(0)
        assert (ci.i0, ci.n_steps) == (i0, n), (kind, V.size)
        if rng.random() < 0.5:
            E = rng.uniform(vmin - 0.2 * (top - vmin), top, nE)  # unsorted, some below the minimum
            E = np.clip(E, vmin - span, None)
            n_g, m_g, x_g = gpu_ctx.sweep(E)
            n_o, m_o, x_o = oracle.sweep(F, s, E)
        else:
            lo = vmin + rng.uniform(0.0, 0.3) * (top - vmin)
            hi = lo + rng.uniform(0.0, 1.0) * (top - lo)
            n_g, m_g, x_g = gpu_ctx.sweep_uniform(lo, hi, nE)
            dE = (hi - lo) / (nE - 1) if nE > 1 else 0.0
            n_o, m_o, x_o = oracle.sweep_uniform(F, s, lo, dE, 0, nE)
        assert np.array_equal(n_g[0], n_o), (seed, done, kind, V.size, nE)
        assert np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o), (seed, done, kind, V.size, nE)
        done += 1


@pytest.mark.parametrize("seed", range(4))
def test_random_level_searches_bit_exact(oracle, gpu_ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    done = 0
    while done < 10:
        kind, V, s = _random_case(rng)
        if V.size < 200:
            continue
        try:
            F, i0, n, vmin = oracle.prep(V, s)
        except ValueError:
            continue
        top = min(vmin + 0.49 / s, float(V.max()))
        n_coarse = int(rng.integers(2, 600))
        v_min = int(rng.integers(0, 4))
        v_max = v_min + int(rng.integers(0, 12))
        M = int(rng.choice([1, 2, 7, 33, 64, 200, 600]))
        rounds = int(rng.integers(0, 9))
        tol = float(rng.choice([1e-6, 1e-10, 1e-13]))
        gpu_ctx.set_potentials(V, s)
        lev_g, wid_g, nb_g = gpu_ctx.solve_levels(vmin, top, n_coarse, v_min, v_max, M, tol, rounds)
        lev_o, wid_o, nb_o, *_ = oracle.solve_levels(F, s, vmin, top, n_coarse, v_min, v_max, M, tol, rounds)
        key = (seed, done, kind, V.size, n_coarse, v_min, v_max, M, rounds, tol)
        assert _same_bits(lev_g[0], lev_o), key
        assert _same_bits(wid_g[0], wid_o) and nb_g[0] == nb_o, key
        done += 1


@pytest.mark.parametrize("seed", range(4))
def test_random_routes_agree(oracle, seed):
    """The other two routes on the same random inputs: constant-bank chunks (bits of nodes AND tails),
    transfer-matrix scan with 2..6 segments (node counts; flagged energies recomputed sequentially)."""
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    rng = np.random.default_rng(3000 + seed)
    with cabi.Context(0) as ctx:
        done = 0
        while done < 16:
            kind, V, s = _random_case(rng)
            if V.size < 300:
                continue
            try:
                F, i0, n, vmin = oracle.prep(V, s)
            except ValueError:
                continue
            top = min(vmin + 0.49 / s, float(V.max()))
            nE = int(rng.integers(1, 900))
            lo = vmin + rng.uniform(0.0, 0.2) * (top - vmin)
            dE = (top - lo) / (nE - 1) if nE > 1 else 0.0
            n_o, m_o, x_o = oracle.sweep_uniform(F, s, lo, dE, 0, nE)
            ctx.set_potentials(V, s)
            ctx.set_option(ctx.OPT_CBANK, 1)
            n_g, m_g, x_g = ctx.sweep_uniform(lo, top, nE)
            ctx.set_option(ctx.OPT_CBANK, 2)
            key = (seed, done, kind, V.size, nE)
            assert np.array_equal(n_g[0], n_o) and np.array_equal(x_g[0], x_o) and _same_bits(m_g[0], m_o), key
            ctx.set_option(ctx.OPT_SCAN_SEGMENTS, int(rng.integers(2, 7)))
            n_s, _, _ = ctx.sweep_uniform(lo, top, nE, tails=False)
            ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 1)
            assert np.array_equal(n_s[0], n_o), key
            ctx.set_option(ctx.OPT_SCAN_SEGMENTS, 0)
            ctx.set_option(ctx.OPT_CBANK, 0)
            done += 1
