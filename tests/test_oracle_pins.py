"""Pin the CPU oracle with independent known answers (CPU only, no GPU).

The reference has no arithmetic for this path, hence no golden vectors of its
own; the oracle is pinned by
  (i)   the committed mpmath replay (tests/golden/numerov_mpmath.json, made by make_golden.py),
  (ii)  the analytic Morse spectrum incl. an h -> h/2 convergence-order check,
  (iii) the harmonic oscillator,
  (iv)  scipy's tridiagonal eigen-solver on the 3-point Hamiltonian (independent discretisation),
  (v)   structural properties: nodes(E) monotone, jumps exactly at the located levels.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from tests import workloads as W

GOLD = json.loads((Path(__file__).parent / "golden" / "numerov_mpmath.json").read_text())


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: c["name"])
def test_mpmath_golden(oracle, case):
    V = np.array([float.fromhex(v) for v in case["V"]])
    s = float.fromhex(case["s"])
    AB, i0, n, _ = oracle.prep(V, s)
    assert (i0, n) == (case["i0"], case["n_steps"])
    E = np.array([float.fromhex(r["E"]) for r in case["rows"]])
    nodes, mant, expo = oracle.sweep(AB, s, E)
    for k, r in enumerate(case["rows"]):
        assert nodes[k] == r["nodes"]
        ref = float.fromhex(r["tail_mant"]) * 2.0 ** (r["tail_exp"] - int(expo[k]))
        assert mant[k] == pytest.approx(ref, rel=1e-7)  # float64 rounding of u, f, g over n steps


def test_morse_analytic_c1(oracle):
    w = W.c1()
    AB, *_ = oracle.prep(w["V"], w["s"])
    exact = np.array(GOLD["morse_c1_levels"])
    lev, wid, nb, rounds, steps = oracle.solve_levels(AB, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 64, 1e-13, 12)
    assert nb == 17 == len(exact)
    rel = np.abs(lev - exact) / exact
    assert rel.max() < 5e-8 and rel[:4].max() < 1e-9


def test_morse_convergence_order(oracle):
    """Halving h must cut the error of the low levels by 2^4 (Numerov is O(h^4)); measured well above
    the h-independent floor that the hard wall at r_min = 0.2 puts on the upper levels."""
    errs = []
    exact = np.array(GOLD["morse_c1_levels"])
    for N in (1251, 2501):
        V = oracle.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 10.0, N)
        s = oracle.scale(W.H2["m0"], W.H2["m1"], W.grid_h(0.2, 10.0, N))
        AB, *_ = oracle.prep(V, s)
        lev, *_ = oracle.solve_levels(AB, s, 0.0, W.H2["De"] - 1.0, 512, 0, 4, 64, 1e-14, 12)
        errs.append(np.abs(lev - exact[:5]))
    ratio = errs[0] / errs[1]
    assert np.all((14.0 < ratio) & (ratio < 18.0)), ratio


def test_harmonic_oscillator(oracle):
    """V = k x^2 / 2 : E_v = hbar*omega (v + 1/2), hbar*omega = 2 sqrt(B k / 2), B = hbar^2/2mu."""
    N, m = 8001, 12.0
    x = np.linspace(-1.5, 1.5, N)
    k = 2.0e5
    V = 0.5 * k * x * x
    s = oracle.scale(m, m, 3.0 / (N - 1))
    B = W.HBAR2_OVER_2 / (m / 2.0)
    hw = 2.0 * np.sqrt(B * k / 2.0)
    AB, *_ = oracle.prep(V, s)
    lev, *_ = oracle.solve_levels(AB, s, 0.0, 12.0 * hw, 512, 0, 7, 64, 1e-13, 12)
    assert np.allclose(lev, hw * (np.arange(8) + 0.5), rtol=2e-9)


def test_tridiagonal_cross_check(oracle):
    from scipy.linalg import eigh_tridiagonal

    N = 4000
    V = oracle.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 8.0, N)
    h = W.grid_h(0.2, 8.0, N)
    s = oracle.scale(W.H2["m0"], W.H2["m1"], h)
    B = W.HBAR2_OVER_2 / (W.H2["m0"] / 2.0)
    d = V[1:-1] + 2.0 * B / h**2
    e = -B / h**2 * np.ones(N - 3)
    ev = eigh_tridiagonal(d, e, select="v", select_range=(0.0, W.H2["De"] - 1.0), eigvals_only=True)
    AB, *_ = oracle.prep(V, s)
    lev, _, nb, *_ = oracle.solve_levels(AB, s, 0.0, W.H2["De"] - 1.0, 512, 0, 16, 64, 1e-12, 12)
    assert nb == len(ev) == 17
    assert np.allclose(lev, ev, rtol=1e-4)  # the 3-point scheme is only O(h^2)


def test_nodes_monotone_and_jump_at_levels(oracle):
    w = W.c1()
    AB, *_ = oracle.prep(w["V"], w["s"])
    E = np.linspace(w["E_lo"], w["E_hi"], 4096)
    nodes, _, _ = oracle.sweep(AB, w["s"], E)
    assert np.all(np.diff(nodes.astype(np.int64)) >= 0)
    lev, wid, *_ = oracle.solve_levels(AB, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 64, 1e-13, 12)
    below, _, _ = oracle.sweep(AB, w["s"], lev - 4.0 * wid - 1e-9)
    above, _, _ = oracle.sweep(AB, w["s"], lev + 4.0 * wid + 1e-9)
    assert np.array_equal(below, np.arange(17)) and np.array_equal(above, np.arange(17) + 1)


def test_uniform_equals_explicit(oracle):
    w = W.c1()
    AB, *_ = oracle.prep(w["V"], w["s"])
    dE = (w["E_hi"] - w["E_lo"]) / 99
    E = w["E_lo"] + np.arange(100) * dE
    a = oracle.sweep(AB, w["s"], E)
    b = oracle.sweep_uniform(AB, w["s"], w["E_lo"], dE, 0, 100)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_omp_build_identical():
    from oracle import Oracle

    w = W.c1()
    a, b = Oracle(), Oracle(omp=True)
    AB, *_ = a.prep(w["V"], w["s"])
    E = np.linspace(w["E_lo"], w["E_hi"], 333)
    for x, y in zip(a.sweep(AB, w["s"], E), b.sweep(AB, w["s"], E)):
        assert np.array_equal(x, y)
    assert b.threads >= 1


def test_prep_window_skips_wall(oracle):
    """Reference fixture curve (test_libepseon_gpu.py:183-190): V(0) ~ 8.9e8 -> window starts later."""
    V = oracle.morse(5500.0, 0.6, 10.0, 0.0, 10.0, 16500)
    s = oracle.scale(87.62, 87.62, W.grid_h(0.0, 10.0, 16500))
    AB, i0, n, vmin = oracle.prep(V, s)
    assert i0 > 1 and i0 + n == 16499
    q = s * V
    assert np.all(q[i0:i0 + n] - q.min() <= 0.5) and q[i0 - 1] - q.min() > 0.5
