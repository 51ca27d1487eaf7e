"""N7 wavefunctions: oracle pins (CPU) and CUDA-vs-oracle parity (GPU, through the C ABI).

Tolerance: the CUDA path performs the oracle's operations in the oracle's order except for the
norm sum (tree reduction on the GPU, sequential on the CPU), so psi agrees to a few ulp of its
maximum: |psi_gpu - psi_cpu| <= 1e-12 * max|psi| is asserted.
"""
import numpy as np
import pytest

from tests import workloads as W


def _harmonic(N=8001, m=12.0, k=2.0e5, half=1.5):
    x = np.linspace(-half, half, N)
    V = 0.5 * k * x * x
    h = 2 * half / (N - 1)
    s = W.scale(m, m, h)
    B = W.HBAR2_OVER_2 / (m / 2.0)
    hw = 2.0 * np.sqrt(B * k / 2.0)
    alpha = np.sqrt(k / (2.0 * B))  # psi_0 ~ exp(-alpha x^2 / 2)
    return x, V, h, s, hw, alpha


def test_oracle_wavefunction_harmonic(oracle):
    """Analytic harmonic-oscillator eigenfunctions (Hermite functions)."""
    from scipy.special import eval_hermite, gammaln

    x, V, h, s, hw, alpha = _harmonic()
    F, i0, n, _ = oracle.prep(V, s)
    lev, *_ = oracle.solve_levels(F, s, 0.0, 12.0 * hw, 512, 0, 5, 64, 1e-13, 12)
    xs = x[i0:i0 + n]
    for v in range(6):
        psi, m = oracle.wavefunction(F, s, lev[v], h)
        assert 0 < m < n
        lognorm = 0.25 * np.log(alpha / np.pi) - 0.5 * (v * np.log(2.0) + gammaln(v + 1))
        ref = np.exp(lognorm - 0.5 * alpha * xs * xs) * eval_hermite(v, np.sqrt(alpha) * xs)
        ref *= np.sign(ref[np.argmax(np.abs(ref) > 1e-6)])  # first lobe positive
        assert np.abs(psi - ref).max() < 2e-7 * np.abs(ref).max()
        assert abs(h * np.sum(psi * psi) - 1.0) < 1e-13
        assert int(np.sum(np.signbit(psi[1:]) != np.signbit(psi[:-1]))) == v


def test_oracle_wavefunction_morse_properties(oracle):
    """C1: normalisation, node count = v, first lobe positive, mutual orthogonality, decaying tail."""
    w = W.c1()
    h = W.grid_h(0.2, 10.0, w["N"])
    F, i0, n, _ = oracle.prep(w["V"], w["s"])
    lev, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 96, 1e-13, 12)
    P = []
    for v, E in enumerate(lev):
        psi, m = oracle.wavefunction(F, w["s"], E, h)
        assert abs(h * np.sum(psi * psi) - 1.0) < 1e-13
        assert int(np.sum(np.signbit(psi[1:]) != np.signbit(psi[:-1]))) == v
        assert psi[np.argmax(np.abs(psi) > 1e-12)] > 0.0
        assert abs(psi[-1]) < 1e-6 * np.abs(psi).max()
        # smooth at the matching point: the kink of the discrete second difference stays small
        d2 = psi[m - 1] - 2 * psi[m] + psi[m + 1]
        d2n = psi[m] - 2 * psi[m + 1] + psi[m + 2]
        assert abs(d2 - d2n) < 1e-3 * np.abs(psi).max()
        P.append(psi)
    G = h * np.array(P) @ np.array(P).T
    assert np.abs(G - np.eye(len(lev))).max() < 1e-8


def test_oracle_wavefunction_grid_convergence(oracle):
    out = {}
    for N in (5001, 10001):
        h = W.grid_h(0.2, 10.0, N)
        V = W.morse(W.H2["De"], W.H2["re"], W.H2["a"], 0.2, 10.0, N)
        s = W.scale(W.H2["m0"], W.H2["m1"], h)
        F, i0, n, _ = oracle.prep(V, s)
        lev, *_ = oracle.solve_levels(F, s, 0.0, W.H2["De"] - 1.0, 1024, 0, 8, 96, 1e-13, 12)
        psi, _ = oracle.wavefunction(F, s, lev[8], h)
        full = np.zeros(N)
        full[i0:i0 + n] = psi
        out[N] = full
    assert np.abs(out[5001] - out[10001][::2]).max() < 5e-7


def test_oracle_wavefunction_below_table(oracle):
    w = W.c1()
    F, *_ = oracle.prep(w["V"], w["s"])
    psi, m = oracle.wavefunction(F, w["s"], -5.0, 1e-3)
    assert m == -1 and not psi.any()
    psi, m = oracle.wavefunction(F, w["s"], float("nan"), 1e-3)
    assert m == -1 and not psi.any()


# ------------------------------------------------------------------ GPU parity
def _gpu_vs_oracle(oracle, ctx, V, s, h, E):
    V = np.atleast_2d(V)
    ctx.set_potentials(V, s)
    E = np.atleast_2d(E)
    psi_g, mi = ctx.wavefunctions(E, h)
    for c in range(V.shape[0]):
        F, i0, n, _ = oracle.prep(V[c], s)
        for l in range(E.shape[1]):
            psi_o, m = oracle.wavefunction(F, s, E[c, l], h)
            full = np.zeros(V.shape[1])
            full[i0:i0 + n] = psi_o
            if m < 0:
                assert mi[c, l] == 0xFFFFFFFF and not psi_g[c, l].any()
                continue
            assert mi[c, l] == m + i0
            scale = np.abs(full).max()
            assert np.abs(psi_g[c, l] - full).max() <= 1e-12 * scale, (c, l)
    return psi_g


@pytest.mark.gpu
def test_gpu_wavefunctions_c1(oracle, gpu_ctx):
    w = W.c1()
    h = W.grid_h(0.2, 10.0, w["N"])
    F, *_ = oracle.prep(w["V"], w["s"])
    lev, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 96, 1e-13, 12)
    E = np.concatenate([lev, [np.nan, -3.0]])  # a skipped row and an energy below the table
    psi = _gpu_vs_oracle(oracle, gpu_ctx, w["V"], w["s"], h, E)
    G = h * psi[0, :17] @ psi[0, :17].T
    assert np.abs(G - np.eye(17)).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("N", [5, 130, 1023, 1025, 2049, 4100])
def test_gpu_wavefunctions_ragged(oracle, gpu_ctx, N):
    """Window lengths around the 1024-step tile and the 128-step renormalisation block."""
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, N)
    h = W.grid_h(1.0, 8.0, N)
    s = W.scale(20.0, 20.0, h)
    if N < 100:
        V = V * 1e-4
    hi = min(V[-1], V.min() + 0.45 / s)
    E = np.linspace(V.min() + 1e-3 * (hi - V.min()), hi, 7)
    _gpu_vs_oracle(oracle, gpu_ctx, V, s, h, E)


@pytest.mark.gpu
def test_gpu_wavefunctions_multi_curve_after_solve(oracle, gpu_ctx):
    w = W.c4(nC=4, N=3000, nE=256)
    h = W.grid_h(0.4, 10.0, 3000)
    gpu_ctx.set_potentials(w["V"], w["s"])
    lev, _, _ = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 256, 0, 5, 64, 1e-12, 10)
    psi = _gpu_vs_oracle(oracle, gpu_ctx, w["V"], w["s"], h, lev)
    for c in range(4):
        for v in range(6):
            p = psi[c, v]
            nz = p[np.abs(p) > 1e-9 * np.abs(p).max()]
            assert int(np.sum(np.signbit(nz[1:]) != np.signbit(nz[:-1]))) == v
            assert abs(h * np.sum(p * p) - 1.0) < 1e-12


@pytest.mark.gpu
def test_gpu_wavefunctions_reference_fixture(oracle, gpu_ctx):
    """Reference fixture curve (wall at r = 0 skipped by the window): psi is zero outside the window."""
    N = 16500
    V = W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, N)
    h = W.grid_h(0.0, 10.0, N)
    s = W.scale(87.62, 87.62, h)
    gpu_ctx.set_potentials(V, s)
    ci = gpu_ctx.curve_info(0)
    lev, _, nb = gpu_ctx.solve_levels(ci.v_min, ci.v_last - 0.1, 2048, 0, 11, 128, 1e-12, 10)
    psi = _gpu_vs_oracle(oracle, gpu_ctx, V, s, h, lev)
    assert not psi[0, :, : ci.i0].any()
