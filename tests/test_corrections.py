"""A-posteriori level corrections: Cooley's energy correction
written as the Rayleigh quotient of the Numerov pencil.  CPU: the oracle's statement is pinned by its
defining property -- E + dE converges quadratically to the discrete eigenvalue.  GPU: the C-ABI entry
against the oracle (tree reduction vs serial sum: tolerance, stated below)."""
import numpy as np
import pytest

from tests import workloads as W

H_C1 = W.grid_h(0.2, 10.0, 10_000)


def _c1_levels(oracle):
    w = W.c1()
    F, i0, n, vmin = oracle.prep(w["V"], w["s"])
    lev, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 2048, 0, 16, 64, 1e-13, 12)
    return w, F, lev


def test_correction_vanishes_at_a_located_level(oracle):
    w, F, lev = _c1_levels(oracle)
    for v in range(17):
        dE = oracle.level_correction(F, w["s"], lev[v], H_C1)
        assert abs(dE) < 1e-10 * lev[v], (v, dE)  # measured: a few 1e-12 relative (rounding noise)


def test_correction_converges_quadratically(oracle):
    w, F, lev = _c1_levels(oracle)
    for v in (0, 5, 11, 16):
        err = []
        for delta in (0.4, 0.2, 0.1):
            for sign in (1.0, -1.0):
                E = lev[v] + sign * delta
                err.append(abs(E + oracle.level_correction(F, w["s"], E, H_C1) - lev[v]))
        e04, e02, e01 = max(err[0:2]), max(err[2:4]), max(err[4:6])
        assert e04 < 0.01 * 0.4 and e01 < 0.002 * 0.1
        assert 3.0 < e04 / e02 < 5.0 and 3.0 < e02 / e01 < 5.0, (v, e04, e02, e01)
    # two correction steps from a 0.5 cm^-1 error reach the k-section answer to 1e-9 relative
    E = lev[7] + 0.5
    for _ in range(2):
        E += oracle.level_correction(F, w["s"], E, H_C1)
    assert abs(E - lev[7]) < 1e-9 * lev[7]


def test_no_matching_point_gives_nan(oracle):
    w, F, lev = _c1_levels(oracle)
    assert np.isnan(oracle.level_correction(F, w["s"], float(w["V"].min()) - 10.0, H_C1))


@pytest.mark.gpu
def test_gpu_level_corrections(oracle, gpu_ctx):
    w, F, lev = _c1_levels(oracle)
    gpu_ctx.set_potentials(w["V"], w["s"])
    E = np.concatenate([lev[:6], lev[6:12] + 0.3, lev[12:] - 0.05, [np.nan, float(w["V"].min()) - 10.0]])
    dE = gpu_ctx.level_corrections(E[None, :], H_C1)[0]
    assert np.isnan(dE[-2]) and np.isnan(dE[-1])
    for k in range(17):
        ref = oracle.level_correction(F, w["s"], E[k], H_C1)
        # psi agrees with the oracle to 1e-12 max|psi|; the second difference at the matching point
        # turns that into ~1e-9 relative on E at worst (measured far below)
        assert abs(dE[k] - ref) <= 1e-9 * abs(E[k]) + 1e-6 * abs(ref), (k, dE[k], ref)
    assert np.all(np.abs(E[6:12] + dE[6:12] - lev[6:12]) < 1e-3 * 0.3)  # 0.3 cm^-1 off -> 1-2e-4 after one step


@pytest.mark.gpu
def test_gpu_level_corrections_batch_after_solve(oracle, gpu_ctx):
    """Multi-curve batch: corrections of the levels a solve just located are below the bracket tolerance."""
    n = 6000
    V = np.stack([W.morse(5500.0, 2.2, 1.6, 0.4, 9.0, n), W.lj(4800.0, 2.4, 0.4, 9.0, n)])
    h = W.grid_h(0.4, 9.0, n)
    s = W.scale(20.0, 20.0, h)
    gpu_ctx.set_potentials(V, s)
    lev, wid, nb = gpu_ctx.solve_levels(V.min(axis=1), V[:, -1] - 1.0, 1024, 0, 9, 64, 1e-12, 10)
    dE = gpu_ctx.level_corrections(lev, h)
    assert dE.shape == lev.shape and np.all(np.abs(dE) < 1e-9 * np.abs(lev))


# --------------------------------------------------------------------------- committed golden vectors

import json  # noqa: E402
from pathlib import Path  # noqa: E402

GOLD_ROT = json.loads((Path(__file__).parent / "golden" / "rotation_mpmath.json").read_text())


@pytest.mark.parametrize("case", GOLD_ROT["corrections"], ids=lambda c: c["name"])
def test_mpmath_golden_corrections(oracle, case):
    """dE from a 60-digit evaluation of the matched solution and the Rayleigh quotient
    (tests/golden/make_golden_rot.py), for trial energies 0.05 .. 0.4 cm^-1 off a level."""
    V = np.array([float.fromhex(v) for v in case["V"]])
    s, h = float.fromhex(case["s"]), float.fromhex(case["h"])
    F, *_ = oracle.prep(V, s)
    for r in case["rows"]:
        E, ref = float.fromhex(r["E"]), float.fromhex(r["dE"])
        psi, m = oracle.wavefunction(F, s, E, h)
        assert m == r["m"]
        assert oracle.level_correction(F, s, E, h) == pytest.approx(ref, rel=1e-7, abs=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLD_ROT["corrections"], ids=lambda c: c["name"])
def test_gpu_against_golden_corrections(gpu_ctx, case):
    V = np.array([float.fromhex(v) for v in case["V"]])
    s, h = float.fromhex(case["s"]), float.fromhex(case["h"])
    gpu_ctx.set_potentials(V, s)
    E = np.array([float.fromhex(r["E"]) for r in case["rows"]])
    dE = gpu_ctx.level_corrections(E[None, :], h)[0]
    for k, r in enumerate(case["rows"]):
        assert dE[k] == pytest.approx(float.fromhex(r["dE"]), rel=1e-7, abs=1e-8 * abs(E[k]))
