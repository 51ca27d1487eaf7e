"""Natural cubic spline resampling of tabulated potentials (N1 for PotentialFileLoader).

Oracle pinned against scipy's CubicSpline(bc_type="natural"); the device evaluation
(eps_spline_resample) and the C++ loader (eps_spline_coefficients + host evaluation) must return
the oracle's bits."""
import numpy as np
import pytest

from tests import workloads as W


def _knots(seed=3, n=64):
    rng = np.random.default_rng(seed)
    rk = np.sort(np.concatenate([np.linspace(0.2, 4.0, n - 16), np.linspace(4.5, 20.0, 16)]))
    Vk = W.H2["De"] * (1 - np.exp(-W.H2["a"] * (rk - W.H2["re"]))) ** 2
    Vk = Vk + 300.0 * rng.standard_normal(rk.size) * np.exp(-((rk - 1.5) / 1.5) ** 2)
    return rk, Vk


@pytest.mark.parametrize("N", [2, 17, 1000, 100_001])
def test_oracle_spline_matches_scipy(oracle, N):
    from scipy.interpolate import CubicSpline

    rk, Vk = _knots()
    out = oracle.spline_resample(rk, Vk, 0.2, 20.0, N)
    r = 0.2 + np.arange(N) * (19.8 / (N - 1))
    ref = CubicSpline(rk, Vk, bc_type="natural")(r)
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    # interpolation property at the knots that fall on the grid, natural end conditions
    assert out[0] == pytest.approx(Vk[0], rel=1e-15)


def test_oracle_spline_three_knots_and_errors(oracle):
    out = oracle.spline_resample([0.0, 1.0, 3.0], [1.0, 0.0, 2.0], 0.0, 3.0, 7)
    assert out[0] == 1.0 and out[-1] == pytest.approx(2.0)
    with pytest.raises(ValueError):
        oracle.spline_resample([0.0, 1.0], [1.0, 0.0], 0.0, 1.0, 5)
    with pytest.raises(ValueError):
        oracle.spline_resample([0.0, 1.0, 1.0], [1.0, 0.0, 3.0], 0.0, 1.0, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2, 1000, 250_000])
def test_gpu_spline_bit_exact(oracle, gpu_ctx, N):
    rk, Vk = _knots(seed=N)
    out_g = gpu_ctx.spline_resample(rk, Vk, 0.2, 20.0, N)
    out_o = oracle.spline_resample(rk, Vk, 0.2, 20.0, N)
    assert np.array_equal(out_g.view(np.uint64), out_o.view(np.uint64))
    # extrapolation outside the knot range uses the end intervals, as in the oracle
    out_g = gpu_ctx.spline_resample(rk, Vk, 0.1, 21.0, 999)
    out_o = oracle.spline_resample(rk, Vk, 0.1, 21.0, 999)
    assert np.array_equal(out_g.view(np.uint64), out_o.view(np.uint64))


@pytest.mark.gpu
def test_gpu_spline_rejects_bad_knots(gpu_ctx):
    from epseon_backend_b200.cabi import EpsError

    with pytest.raises(EpsError) as ei:
        gpu_ctx.spline_resample(np.array([0.0, 2.0, 1.0]), np.array([1.0, 2.0, 3.0]), 0.0, 2.0, 10)
    assert ei.value.code == 1


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["dat", "npy"])
def test_file_loader_resamples_non_uniform_table(tmp_path, oracle, fmt, oracle_d):
    """PotentialFileLoader made real for ab initio tables: a 64-knot non-uniform 'r V' file (text, or
    a NumPy .npy array of shape (n, 2)), resampled by the C++ loader to 20 001 points, solved through
    the Python API == the oracle on the oracle's own resampling of the same knots."""
    from epseon_backend.device.gpu import _libepseon_gpu as gpu_mod

    rk, Vk = _knots(seed=11)
    path = tmp_path / f"curve.{fmt}"
    if fmt == "npy":
        np.save(path, np.column_stack([rk, Vk]))
    else:
        with open(path, "w") as f:
            f.write("# r [Angstrom]   V [cm^-1]\n")
            for a, b in zip(rk, Vk):
                f.write(f"{float(a)!r} {float(b)!r}\n")
    N = 20_001
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = interface.get_task_configurator("float64")
    cfg.set_hardware_config(potential_buffer_size=N, group_size=1024, allocation_block_size=1 << 20)
    cfg.set_potential_files([str(path)], point_count=N)
    cfg.set_vibwa_algorithm(mass_atom_0=W.H2["m0"], mass_atom_1=W.H2["m1"], integration_step=0.1,
                            min_distance_to_asymptote=1.0, min_level=0, max_level=9)
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    levels = np.array(handle.get_levels())[0]
    V = oracle.spline_resample(rk, Vk, rk[0], rk[-1], N)
    s = oracle.scale(W.H2["m0"], W.H2["m1"], (rk[-1] - rk[0]) / (N - 1))
    F, _, _, vmin = oracle_d.prep(V, s)
    n_coarse, M, rounds, tol = handle.get_search_parameters()  # one curve: the rounds are sized to the device (vibwa_run.hpp)
    assert (n_coarse, tol) == (1024, 1e-12) and M >= 256
    ref, *_ = oracle_d.solve_levels(F, s, vmin, V[-1] - 1.0, n_coarse, 0, 9, M, tol, rounds)
    assert np.array_equal(levels.view(np.uint64), ref.view(np.uint64))
