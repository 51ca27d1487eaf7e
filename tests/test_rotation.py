"""Rotational states: V_J = V + J(J+1) (hbar^2/2mu) / r^2.

CPU part pins the oracle's statement (orc_centrifugal) with independent answers: the Pekeris
rovibrational formula of the Morse oscillator and scipy's tridiagonal eigen-solver on the
finite-difference Hamiltonian *with* the centrifugal term.  GPU part: eps_set_potentials_rot expands
the tables on the device; effective tables, node counts, tails and levels are bit-identical to the
oracle's on the same inputs.
"""
import numpy as np
import pytest

from tests import workloads as W

RMIN, RMAX, N = 0.2, 10.0, 10_000


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _h2():
    h = W.grid_h(RMIN, RMAX, N)
    return W.morse(W.H2["De"], W.H2["re"], W.H2["a"], RMIN, RMAX, N), W.scale(W.H2["m0"], W.H2["m1"], h), h


def _levels(oracle, V, s, vmax, E_hi):
    AB, *_ = oracle.prep(V, s)
    lev, *_ = oracle.solve_levels(AB, s, float(V.min()), E_hi, 2048, 0, vmax, 64, 1e-12, 12)
    return lev


# --------------------------------------------------------------------------- CPU pins


def test_j0_is_a_bit_copy(oracle):
    V, s, h = _h2()
    assert _same_bits(oracle.centrifugal(V, s, RMIN, h, 0), V)


def test_formula_against_numpy(oracle):
    V, s, h = _h2()
    r = RMIN + np.arange(N, dtype=np.float64) * h
    for J in (1, 2, 7, 30):
        cj = (float(J * (J + 1)) * (h * h)) / (12.0 * s)
        assert _same_bits(oracle.centrifugal(V, s, RMIN, h, J), V + cj / (r * r))
    # cj is J(J+1) * hbar^2/2mu
    mu = W.H2["m0"] * W.H2["m1"] / (W.H2["m0"] + W.H2["m1"])
    assert abs((h * h) / (12.0 * s) - W.HBAR2_OVER_2 / mu) < 1e-12 * W.HBAR2_OVER_2 / mu


def test_pekeris_rotational_constants(oracle):
    """E(v,J) - E(v,0) = B_v J(J+1) - D_e [J(J+1)]^2 with Pekeris' alpha_e and Kratzer's D_e.

    The formula is first order in (v + 1/2): measured deviation 3e-5 for v = 0, growing as
    (v + 1/2)^2 (the neglected gamma_e term; H2 is strongly anharmonic) to 0.8 % at v = 2."""
    V, s, h = _h2()
    mu = W.H2["m0"] * W.H2["m1"] / (W.H2["m0"] + W.H2["m1"])
    B = W.HBAR2_OVER_2 / mu
    a, De, re = W.H2["a"], W.H2["De"], W.H2["re"]
    we, wexe, Be = 2 * a * np.sqrt(De * B), a * a * B, B / re**2
    alpha = 6 * np.sqrt(wexe * Be**3) / we - 6 * Be**2 / we
    Dcen = 4 * Be**3 / we**2
    E0 = _levels(oracle, V, s, 2, De - 1.0)
    for J in (1, 2, 3):
        EJ = _levels(oracle, oracle.centrifugal(V, s, RMIN, h, J), s, 2, De - 1.0)
        x = J * (J + 1)
        for v in range(3):
            want = (Be - alpha * (v + 0.5)) * x - Dcen * x * x
            tol = 1e-4 if v == 0 else 1e-2
            assert abs((EJ[v] - E0[v]) - want) < tol * want, (J, v, EJ[v] - E0[v], want)


def test_levels_against_tridiagonal(oracle):
    """Independent discretisation (3-point FD Hamiltonian incl. the centrifugal term, O(h^2))."""
    from scipy.linalg import eigh_tridiagonal

    n = 4000
    h = W.grid_h(RMIN, 6.0, n)
    V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], RMIN, 6.0, n)
    s = W.scale(W.H2["m0"], W.H2["m1"], h)
    mu = W.H2["m0"] * W.H2["m1"] / (W.H2["m0"] + W.H2["m1"])
    B = W.HBAR2_OVER_2 / mu
    r = RMIN + np.arange(n) * h
    J = 4
    VJ = V + J * (J + 1) * B / r**2
    ev = eigh_tridiagonal(2 * B / h**2 + VJ[1:-1], -B / h**2 * np.ones(n - 3), select="i", select_range=(0, 3))[0]
    lev = _levels(oracle, oracle.centrifugal(V, s, RMIN, h, J), s, 3, W.H2["De"] - 1.0)
    # FD error of level v ~ (h^2/12) <psi''''> B: a few 1e-5 relative on this grid
    assert np.all(np.abs(lev - ev) < 1e-4 * ev), (lev, ev)


def test_origin_point_becomes_a_wall(oracle):
    """The reference fixture's grid starts at r = 0 (python/test/.../test_libepseon_gpu.py:183-190)."""
    n = 16500
    V = W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, n)
    h = W.grid_h(0.0, 10.0, n)
    s = W.scale(87.62, 87.62, h)
    VJ = oracle.centrifugal(V, s, 0.0, h, 3)
    assert VJ[0] == 1e300 and np.all(np.isfinite(VJ)) and np.all(VJ[1:] > V[1:])
    lev0 = _levels(oracle, V, s, 2, 5400.0)
    lev3 = _levels(oracle, VJ, s, 2, 5400.0)
    d = lev3 - lev0  # B_e = 16.86/43.81/0.36 = 1.07 cm^-1, J(J+1) = 12, falling with v
    assert np.all((d > 11.0) & (d < 13.0)) and np.all(np.diff(d) < 0)


# --------------------------------------------------------------------------- GPU parity


@pytest.mark.gpu
def test_gpu_rot_tables_sweeps_levels_bit_exact(oracle, gpu_ctx):
    n = 5000
    rmin = np.array([0.4, 0.5])
    h = np.array([W.grid_h(0.4, 9.0, n), W.grid_h(0.5, 10.0, n)])
    V = np.stack([W.morse(5500.0, 2.2, 1.6, 0.4, 9.0, n), W.lj(4800.0, 2.4, 0.5, 10.0, n)])
    s = np.array([W.scale(20.0, 20.0, h[0]), W.scale(18.0, 22.0, h[1])])
    Js = np.array([0, 1, 5, 40], dtype=np.uint32)
    gpu_ctx.set_potentials_rot(V, s, rmin, h, Js)
    assert gpu_ctx.n_curves == 8
    nE = 300
    E_lo, E_hi, tables = [], [], []
    for c in range(2):
        for J in Js:
            VJ = oracle.centrifugal(V[c], s[c], rmin[c], h[c], int(J))
            AB, i0, nst, vmin = oracle.prep(VJ, s[c])
            ci = gpu_ctx.curve_info(len(tables))
            assert (ci.i0, ci.n_steps, ci.scale) == (i0, nst, s[c])
            assert _same_bits([ci.v_min, ci.v_last], [vmin, VJ[-1]])  # the device table is the oracle's
            tables.append((AB, s[c]))
            E_lo.append(vmin)
            E_hi.append(VJ[-1] - 1.0)
    n_g, m_g, x_g = gpu_ctx.sweep_uniform(np.array(E_lo), np.array(E_hi), nE)
    lev_g, wid_g, nb_g = gpu_ctx.solve_levels(np.array(E_lo), np.array(E_hi), 512, 0, 5, 64, 1e-12, 10)
    for k, (AB, sk) in enumerate(tables):
        dE = (E_hi[k] - E_lo[k]) / (nE - 1)
        n_o, m_o, x_o = oracle.sweep_uniform(AB, sk, E_lo[k], dE, 0, nE)
        assert np.array_equal(n_g[k], n_o) and np.array_equal(x_g[k], x_o) and _same_bits(m_g[k], m_o), k
        lev_o, wid_o, nb_o, *_ = oracle.solve_levels(AB, sk, E_lo[k], E_hi[k], 512, 0, 5, 64, 1e-12, 10)
        assert _same_bits(lev_g[k], lev_o) and nb_g[k] == nb_o, k
    # rotational ladder: E(v, J) grows with J
    lev = lev_g.reshape(2, len(Js), 6)
    assert np.all(np.diff(lev[:, :, 0], axis=1) > 0)


@pytest.mark.gpu
def test_gpu_rot_j0_equals_plain_upload(oracle, gpu_ctx):
    w = W.c1()
    h = W.grid_h(0.2, 10.0, w["N"])
    gpu_ctx.set_potentials(w["V"], w["s"])
    n0, m0, x0 = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 257)
    gpu_ctx.set_potentials_rot(w["V"], w["s"], 0.2, h, [0])
    n1, m1, x1 = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 257)
    assert np.array_equal(n0, n1) and np.array_equal(x0, x1) and _same_bits(m0, m1)


@pytest.mark.gpu
def test_gpu_rot_origin_wall_and_errors(oracle, gpu_ctx):
    from epseon_backend_b200 import cabi

    n = 16500
    V = W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, n)
    h = W.grid_h(0.0, 10.0, n)
    s = W.scale(87.62, 87.62, h)
    gpu_ctx.set_potentials_rot(V, s, 0.0, h, [2, 0])
    for k, J in enumerate((2, 0)):
        VJ = oracle.centrifugal(V, s, 0.0, h, J)
        AB, i0, nst, vmin = oracle.prep(VJ, s)
        ci = gpu_ctx.curve_info(k)
        assert (ci.i0, ci.n_steps) == (i0, nst) and _same_bits([ci.v_min], [vmin])
    with pytest.raises(cabi.EpsError) as e:
        gpu_ctx.set_potentials_rot(V, s, 0.0, 0.0, [1])  # grid_step must be positive
    assert e.value.code == 1
    with pytest.raises(cabi.EpsError) as e:
        gpu_ctx.set_potentials_rot(V, s, 0.0, h, [1 << 26])
    assert e.value.code == 1


# --------------------------------------------------------------------------- committed golden vectors

import json  # noqa: E402
from pathlib import Path  # noqa: E402

GOLD_ROT = json.loads((Path(__file__).parent / "golden" / "rotation_mpmath.json").read_text())


@pytest.mark.parametrize("case", GOLD_ROT["rotation"], ids=lambda c: c["name"])
def test_mpmath_golden_rotating_curves(oracle, case):
    """tests/golden/rotation_mpmath.json (make_golden_rot.py): V_J from an independent numpy statement,
    node counts from a 60-digit replay of the recurrence on it."""
    V = np.array([float.fromhex(v) for v in case["V"]])
    s, h = float.fromhex(case["s"]), float.fromhex(case["h"])
    VJ = oracle.centrifugal(V, s, case["rmin"], h, case["J"])
    assert VJ[0] == float.fromhex(case["VJ_first"]) and VJ[V.size // 2] == float.fromhex(case["VJ_mid"])
    F, i0, n, _ = oracle.prep(VJ, s)
    assert (i0, n) == (case["i0"], case["n_steps"])
    E = np.array([float.fromhex(r["E"]) for r in case["rows"]])
    nodes, mant, expo = oracle.sweep(F, s, E)
    for k, r in enumerate(case["rows"]):
        assert nodes[k] == r["nodes"]
        ref = float.fromhex(r["tail_mant"]) * 2.0 ** (r["tail_exp"] - int(expo[k]))
        assert mant[k] == pytest.approx(ref, rel=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD_ROT["rotation"], ids=lambda c: c["name"])
def test_gpu_against_golden_rotating_curves(gpu_ctx, case):
    """The CUDA path against the same committed vectors (no oracle involved)."""
    V = np.array([float.fromhex(v) for v in case["V"]])
    s, h = float.fromhex(case["s"]), float.fromhex(case["h"])
    gpu_ctx.set_potentials_rot(V, s, case["rmin"], h, [case["J"]])
    ci = gpu_ctx.curve_info(0)
    assert (ci.i0, ci.n_steps) == (case["i0"], case["n_steps"])
    E = np.array([float.fromhex(r["E"]) for r in case["rows"]])
    nodes, mant, expo = gpu_ctx.sweep(E)
    for k, r in enumerate(case["rows"]):
        assert nodes[0][k] == r["nodes"]
        ref = float.fromhex(r["tail_mant"]) * 2.0 ** (r["tail_exp"] - int(expo[0][k]))
        assert mant[0][k] == pytest.approx(ref, rel=1e-7)


@pytest.mark.gpu
def test_gpu_rot_many_rows(oracle, gpu_ctx):
    """More (curve, J) rows than a grid's y extent allows (65 535): rows ride on grid.x."""
    n, nC, nJ = 64, 1100, 64
    rng = np.random.default_rng(9)
    x = np.linspace(0.5, 3.5, n)
    V = (rng.uniform(50.0, 150.0, (nC, 1)) * (x[None, :] - 2.0) ** 2).astype(np.float64)
    h = W.grid_h(0.5, 3.5, n)
    s = W.scale(10.0, 12.0, h)
    Js = np.arange(nJ, dtype=np.uint32)
    gpu_ctx.set_potentials_rot(V, s, 0.5, h, Js)
    assert gpu_ctx.n_curves == nC * nJ > 65535
    for c, j in ((0, 0), (0, 63), (517, 31), (1099, 63), (1023, 1)):
        VJ = oracle.centrifugal(V[c], s, 0.5, h, int(Js[j]))
        F, i0, nst, vmin = oracle.prep(VJ, s)
        ci = gpu_ctx.curve_info(c * nJ + j)
        assert (ci.i0, ci.n_steps) == (i0, nst) and _same_bits([ci.v_min, ci.v_last], [vmin, VJ[-1]])
