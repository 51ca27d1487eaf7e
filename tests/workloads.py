"""Synthetic inputs of BASELINE.md section 3 (C1..C5), shared by tests and bench.py.

Units cm^-1 / Angstrom / amu.  Potential tables are produced by the caller's
tabulator (oracle in tests, numpy here) -- the same float64 table is handed to
both the CUDA path and the oracle, so parity never depends on libm.
"""
from __future__ import annotations

import numpy as np

HBAR2_OVER_2 = 16.857629206  # amu * Angstrom^2 * cm^-1

# H2-like Morse curve used by C1/C2/C5
H2 = dict(De=38267.0, a=1.9426, re=0.7414, m0=1.00783, m1=1.00783)


def scale(m0: float, m1: float, h: float) -> float:
    """s = h^2 (2 mu / hbar^2) / 12, same operation order as the product and the oracle."""
    mu = (m0 * m1) / (m0 + m1)
    c = mu / HBAR2_OVER_2
    return ((h * h) * c) / 12.0


def grid_h(rmin: float, rmax: float, N: int) -> float:
    return (rmax - rmin) / float(N - 1)


def morse(De, re, a, rmin, rmax, N) -> np.ndarray:
    h = grid_h(rmin, rmax, N)
    r = rmin + np.arange(N, dtype=np.float64) * h
    t = 1.0 - np.exp(-a * (r - re))
    return (De * t) * t


def lj(De, re, rmin, rmax, N) -> np.ndarray:
    h = grid_h(rmin, rmax, N)
    r = rmin + np.arange(N, dtype=np.float64) * h
    x6 = (re / r) ** 6
    return De * ((x6 * x6 - 2.0 * x6) + 1.0)


def morse_levels(De, a, m0, m1) -> np.ndarray:
    """Analytic Morse spectrum E_v = 2a sqrt(De B)(v+1/2) - a^2 B (v+1/2)^2, B = hbar^2/(2 mu)."""
    mu = (m0 * m1) / (m0 + m1)
    B = HBAR2_OVER_2 / mu
    lam = np.sqrt(De / B) / a
    v = np.arange(int(np.floor(lam - 0.5)) + 1, dtype=np.float64)
    return 2 * a * np.sqrt(De * B) * (v + 0.5) - a * a * B * (v + 0.5) ** 2


def c1():
    """C1: H2-like Morse, N=10 000 on [0.2,10], 1024 energies uniform in [0, De-1]."""
    N, rmin, rmax = 10_000, 0.2, 10.0
    V = morse(H2["De"], H2["re"], H2["a"], rmin, rmax, N)
    return dict(V=V, s=scale(H2["m0"], H2["m1"], grid_h(rmin, rmax, N)), E_lo=0.0,
                E_hi=H2["De"] - 1.0, nE=1024, N=N)


def c2(N: int = 100_000, nE: int = 65_536):
    """C2: same curve on [0.2,12], N=100 000, 65 536 energies, all 17 bound levels."""
    rmin, rmax = 0.2, 12.0
    V = morse(H2["De"], H2["re"], H2["a"], rmin, rmax, N)
    return dict(V=V, s=scale(H2["m0"], H2["m1"], grid_h(rmin, rmax, N)), E_lo=0.0,
                E_hi=H2["De"] - 1.0, nE=nE, N=N)


def c3(N: int = 1_000_000, nE: int = 4096, knots: int = 64):
    """C3: 64-knot 'ab initio' table = Morse + seeded smooth 1% perturbation, natural cubic
    spline resampled to N points on [0.2, 20]."""
    from scipy.interpolate import CubicSpline

    rng = np.random.default_rng(1234)
    rmin, rmax = 0.2, 20.0
    rk = np.concatenate([np.linspace(rmin, 4.0, knots - 16), np.linspace(4.5, rmax, 16)])
    t = 1.0 - np.exp(-H2["a"] * (rk - H2["re"]))
    Vk = H2["De"] * t * t
    noise = rng.standard_normal(knots)
    noise = np.convolve(noise, np.ones(5) / 5.0, mode="same")  # low-pass
    Vk = Vk + 0.01 * H2["De"] * noise * np.exp(-((rk - 1.5) / 1.5) ** 2)
    cs = CubicSpline(rk, Vk, bc_type="natural")
    r = rmin + np.arange(N, dtype=np.float64) * grid_h(rmin, rmax, N)
    V = np.ascontiguousarray(cs(r))
    return dict(V=V, s=scale(H2["m0"], H2["m1"], grid_h(rmin, rmax, N)), E_lo=float(V.min()),
                E_hi=float(V[-1]) - 1.0, nE=nE, N=N)


def c4(nC: int = 4096, N: int = 10_000, nE: int = 1024):
    """C4: nC Morse/LJ curves with De, a, re jittered +-5% (default_rng(2024))."""
    rng = np.random.default_rng(2024)
    rmin, rmax = 0.4, 10.0
    De0, a0, re0, m = 5500.0, 1.6, 2.2, 20.0
    V = np.empty((nC, N), dtype=np.float64)
    for c in range(nC):
        j = 1.0 + 0.05 * (2.0 * rng.random(3) - 1.0)
        if c % 2 == 0:
            V[c] = morse(De0 * j[0], re0 * j[2], a0 * j[1], rmin, rmax, N)
        else:
            V[c] = lj(De0 * j[0], re0 * j[2], rmin, rmax, N)
    s = scale(m, m, grid_h(rmin, rmax, N))
    return dict(V=V, s=s, E_lo=V.min(axis=1), E_hi=V[:, -1] - 1.0, nE=nE, N=N)


def c5(N: int = 200_000, nE: int = 1 << 24):
    """C5: C1 curve on a 200k grid, 2^24 uniform energies."""
    rmin, rmax = 0.2, 10.0
    V = morse(H2["De"], H2["re"], H2["a"], rmin, rmax, N)
    return dict(V=V, s=scale(H2["m0"], H2["m1"], grid_h(rmin, rmax, N)), E_lo=0.0,
                E_hi=H2["De"] - 1.0, nE=nE, N=N)
