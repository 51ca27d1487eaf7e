"""The reference's Python surface (module ``_libepseon_gpu``) on the CUDA build.

GPU-marked tests restate python/test/test_device/test_gpu/test_libepseon_gpu.py (reference) case by
case -- same calls, same assertions -- and then check the results the reference cannot produce.
CPU tests cover what needs no device: import path, class list, constructor keywords, error types,
the ``format`` helper and ``_libepseon_cpu.greet``.
"""
import gc
import re
from pathlib import Path

import numpy as np
import pytest

from tests import workloads as W

COMPUTE_GROUP_AXES_COUNT = 3
ROOT = Path(__file__).resolve().parent.parent

REFERENCE_CLASSES = {
    "EpseonComputeContext", "ComputeDeviceInterface", "TaskConfiguratorFloat32", "TaskConfiguratorFloat64",
    "MorsePotentialConfig", "TaskHandleFloat32", "TaskHandleFloat64", "PhysicalDeviceInfo",
    "PhysicalDeviceProperties", "PhysicalDeviceLimits", "PhysicalDeviceSparseProperties",
    "PhysicalDeviceMemoryProperties", "MemoryHeap", "MemoryType",
}


@pytest.fixture(scope="module")
def gpu_mod():
    import __graft_entry__ as ge

    ge.build()
    import epseon_backend.device.gpu._libepseon_gpu as m

    return m


def _configure_task(m, configurator, max_level=0, point_count=16500):
    return (
        configurator.set_hardware_config(potential_buffer_size=16500, group_size=512,
                                         allocation_block_size=16 * 1024 * 1024)
        .set_morse_potential([
            m.MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6, well_width=10,
                                   min_r=0.0, max_r=10.0, point_count=point_count),
            m.MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6, well_width=10,
                                   min_r=0.0, max_r=10.0, point_count=point_count),
        ])
        .set_vibwa_algorithm(mass_atom_0=87.62, mass_atom_1=87.62, integration_step=0.1,
                             min_distance_to_asymptote=0.1, min_level=0, max_level=max_level)
    )


# ----------------------------------------------------------------------------- CPU
def test_import_path_and_class_list(gpu_mod):
    names = {n for n in dir(gpu_mod) if not n.startswith("_")}
    assert REFERENCE_CLASSES <= names
    import epseon_backend

    assert epseon_backend.__version__


def test_morse_config_keywords(gpu_mod):
    cfg = gpu_mod.MorsePotentialConfig(dissociation_energy=500, equilibrium_bond_distance=2.6, well_width=1.3,
                                       min_r=0.0, max_r=10.0, point_count=16500)
    assert repr(cfg)
    with pytest.raises(TypeError):
        gpu_mod.MorsePotentialConfig(1.0, 2.0)  # all six are required, as in the reference


def test_cpu_module_greet():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend.device.cpu._libepseon_cpu import greet

    assert greet() == "Hello, World from C++!"


@pytest.mark.parametrize(("value", "expect"), [(0, "0.000B"), (15, "15.000B"), (1024, "1.000KiB"),
                                               (256 * 1024**2, "256.000MiB"), (1024**11, "1073741824.000YiB")])
def test_format_helper(value, expect):
    from epseon_backend.format import convert_size_in_bytes_to_adaptive_unit

    assert convert_size_in_bytes_to_adaptive_unit(value) == expect


def test_create_without_gpu_raises_runtime_error(gpu_mod):
    from epseon_backend_b200 import cabi

    try:
        n = cabi.device_count()
    except cabi.EpsError:
        n = 0
    if n == 0:
        with pytest.raises(RuntimeError, match="Failed to create EpseonComputeContext."):
            gpu_mod.EpseonComputeContext.create()


# ----------------------------------------------------------------------------- GPU (reference cases)
@pytest.mark.gpu
class TestEpseonComputeContext:
    def test_create_epseon_compute_context(self, gpu_mod):
        assert repr(gpu_mod.EpseonComputeContext.create())

    def test_get_vulkan_version(self, gpu_mod):
        ctx = gpu_mod.EpseonComputeContext.create()
        assert re.match(r"\d+\.\d+\.\d+\.\d+", ctx.get_vulkan_version()) is not None

    def test_get_physical_device_info(self, gpu_mod):
        ctx = gpu_mod.EpseonComputeContext.create()
        infos = tuple(ctx.get_physical_device_info())
        assert len(infos) > 0
        for device in infos:
            p = device.device_properties
            assert isinstance(p.api_version, str) and isinstance(p.driver_version, str)
            assert isinstance(p.vendor_id, int) and isinstance(p.device_id, int)
            assert p.device_type in ("INTEGRATED_GPU", "DISCRETE_GPU", "VIRTUAL_GPU", "CPU", "OTHER")
            assert isinstance(p.device_name, str) and len(p.device_name) > 0
            assert isinstance(p.pipeline_cache_uuid, list) and len(p.pipeline_cache_uuid) == 16
            assert isinstance(p.limits.max_compute_shared_memory_size, int)
            assert isinstance(p.limits.max_compute_work_group_count, tuple)
            assert len(p.limits.max_compute_work_group_count) == COMPUTE_GROUP_AXES_COUNT
            assert isinstance(p.limits.max_compute_work_group_invocations, int)
            assert isinstance(p.limits.max_compute_work_group_size, int)
            assert hasattr(p, "sparse_properties")
            assert hasattr(device, "memory_properties")
            heaps = device.memory_properties.memory_heaps
            assert len(heaps) > 0 and heaps[0].size > 0 and "DEVICE_LOCAL" in heaps[0].flags
            for t in device.memory_properties.memory_types:
                assert isinstance(t.heap_index, int) and isinstance(t.flags, list)
        # device ids are unique per GPU and round-trip through get_device_interface
        ids = [d.device_properties.device_id for d in infos]
        assert len(set(ids)) == len(ids)
        for i in ids:
            assert id(ctx.get_device_interface(i))
        with pytest.raises(RuntimeError, match="Device not available."):
            ctx.get_device_interface(10_000)

    @pytest.mark.parametrize("precision", ["float32", "float64", "Float64"])
    def test_get_and_use_task_configurator(self, gpu_mod, precision):
        ctx = gpu_mod.EpseonComputeContext.create()
        interface = ctx.get_device_interface(0)
        cfg = _configure_task(gpu_mod, interface.get_task_configurator(precision))
        assert id(cfg) and cfg.is_configured()

    def test_unknown_precision(self, gpu_mod):
        interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
        with pytest.raises(ValueError, match='Invalid PrecisionType literal in string: "float80"'):
            interface.get_task_configurator("float80")

    def test_builder_identity(self, gpu_mod):
        interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
        a = interface.get_task_configurator("float32")
        b = a.set_hardware_config(potential_buffer_size=16500, group_size=4096, allocation_block_size=64 * 1024 * 1024)
        c = _configure_task(gpu_mod, b)
        assert id(a) == id(b) == id(c)

    def test_mismatched_point_count(self, gpu_mod):
        interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
        cfgr = interface.get_task_configurator("float64")
        mk = lambda n: gpu_mod.MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6,  # noqa: E731
                                                    well_width=10, min_r=0.0, max_r=10.0, point_count=n)
        with pytest.raises(RuntimeError, match="same point count"):
            cfgr.set_morse_potential([mk(100), mk(101)])

    def test_submit_unconfigured(self, gpu_mod):
        interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
        with pytest.raises(RuntimeError):
            interface.submit_task(interface.get_task_configurator("float64"))

    @pytest.mark.parametrize("precision", ["float32", "float64"])
    def test_submit_task(self, gpu_mod, precision):
        interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
        handle = interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator(precision)))
        assert isinstance(handle.get_status_message(), str)

    @pytest.mark.parametrize("precision", ["float32", "float64"])
    def test_submit_task_with_ctx_going_out_of_scope(self, gpu_mod, precision):
        ctx = gpu_mod.EpseonComputeContext.create()
        interface = ctx.get_device_interface(0)
        del ctx
        gc.collect(0), gc.collect(1), gc.collect(2)
        handle = interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator(precision)))
        assert isinstance(handle.get_status_message(), str)
        handle.wait()
        assert handle.is_done() and not handle.has_failed()

    def test_submit_and_wait_every_device(self, gpu_mod):
        ctx = gpu_mod.EpseonComputeContext.create()
        for info in ctx.get_physical_device_info():
            interface = ctx.get_device_interface(info.device_properties.device_id)
            handle = interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator("float32")))
            handle.wait()
            assert handle.is_done()


# ----------------------------------------------------------------------------- GPU (results: additive API)
@pytest.mark.gpu
def test_levels_of_reference_fixture(gpu_mod, oracle, oracle_d):
    """submit_task on the reference's fixture curve (two identical Sr2-like Morse curves, N = 16500):
    levels equal the oracle's bits (float64) and the analytic Morse spectrum."""
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = _configure_task(gpu_mod, interface.get_task_configurator("float64"), max_level=13)
    handle = interface.submit_task(cfg)
    handle.wait()
    assert handle.is_done() and not handle.has_failed(), handle.get_status_message()
    assert handle.get_status_message() == "done"
    levels = np.array(handle.get_levels())
    counts = handle.get_level_counts()
    exact = W.morse_levels(5500.0, 10.0, 87.62, 87.62)
    assert levels.shape == (2, 14) and counts == [len(exact)] * 2 == [12, 12]
    assert np.array_equal(levels[0], levels[1], equal_nan=True)
    assert np.all(np.isnan(levels[0][12:]))
    assert np.max(np.abs(levels[0][:12] - exact) / exact) < 2e-6
    # oracle with the parameters VibwaAlgorithm<FP>::run uses (vibwa_run.hpp): n_coarse = 1024, M = 256
    N = 16500
    V = oracle.morse(5500.0, 0.6, 10.0, 0.0, 10.0, N)
    s = oracle.scale(87.62, 87.62, W.grid_h(0.0, 10.0, N))
    F, _, _, vmin = oracle_d.prep(V, s)
    lev_o, *_ = oracle_d.solve_levels(F, s, vmin, V[-1] - 0.1, 1024, 0, 13, 256, 1e-12, 16)
    assert np.array_equal(levels[0].view(np.uint64), lev_o.view(np.uint64))


@pytest.mark.gpu
def test_levels_float32_task(gpu_mod):
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = _configure_task(gpu_mod, interface.get_task_configurator("float32"), max_level=3)
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    levels = np.array(handle.get_levels())
    exact = W.morse_levels(5500.0, 10.0, 87.62, 87.62)
    assert levels.shape == (2, 4)
    assert np.max(np.abs(levels[0] - exact[:4]) / exact[:4]) < 1e-5  # float32 curve + float32 output


@pytest.mark.gpu
def test_failure_is_reported_not_fatal(gpu_mod):
    """A task whose grid is too coarse for its energy range must fail cleanly (status + has_failed),
    where the reference would std::terminate on a worker exception (vibwa.hpp:352-354)."""
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfgr = interface.get_task_configurator("float64")
    cfgr.set_hardware_config(potential_buffer_size=64, group_size=16, allocation_block_size=1024)
    cfgr.set_morse_potential([gpu_mod.MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6,
                                                           well_width=10, min_r=0.0, max_r=10.0, point_count=64)])
    cfgr.set_vibwa_algorithm(mass_atom_0=87.62, mass_atom_1=87.62, integration_step=0.1,
                             min_distance_to_asymptote=0.1, min_level=0, max_level=2)
    handle = interface.submit_task(cfgr)
    handle.wait()
    assert handle.is_done() and handle.has_failed()
    assert handle.get_status_message().startswith("failed:")


@pytest.mark.gpu
def test_wavefunctions_through_python_api(gpu_mod, oracle):
    """Additive surface: set_wavefunction_output(True) -> TaskHandle.get_wavefunctions() returns the
    normalised wavefunctions [curve][level][point] of the levels the task located."""
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = _configure_task(gpu_mod, interface.get_task_configurator("float64"), max_level=5)
    assert cfg.set_wavefunction_output(True) is cfg
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    psi = handle.get_wavefunctions()
    levels = np.array(handle.get_levels())
    assert psi.shape == (2, 6, 16500) and psi.dtype == np.float64
    h = W.grid_h(0.0, 10.0, 16500)
    for v in range(6):
        p = psi[0, v]
        assert abs(h * np.sum(p * p) - 1.0) < 1e-12
        nz = p[np.abs(p) > 1e-9 * np.abs(p).max()]
        assert int(np.sum(np.signbit(nz[1:]) != np.signbit(nz[:-1]))) == v
    V = oracle.morse(5500.0, 0.6, 10.0, 0.0, 10.0, 16500)
    s = oracle.scale(87.62, 87.62, h)
    F, i0, n, _ = oracle.prep(V, s)
    ref, m = oracle.wavefunction(F, s, levels[0, 3], h)
    assert np.abs(psi[0, 3, i0:i0 + n] - ref).max() <= 1e-12 * np.abs(ref).max()
    # not requested -> empty array
    handle2 = interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator("float64")))
    handle2.wait()
    assert handle2.get_wavefunctions().size == 0


@pytest.mark.gpu
def test_rotational_states_through_python_api(gpu_mod, oracle, oracle_d):
    """Additive surface: set_rotational_states([J...]) -> rows [curve][J]; levels carry
    the bits of the oracle run on orc_centrifugal's table."""
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = _configure_task(gpu_mod, interface.get_task_configurator("float64"), max_level=3)
    with pytest.raises(RuntimeError):
        cfg.set_rotational_states([])
    with pytest.raises(RuntimeError):
        cfg.set_rotational_states([1 << 26])
    assert cfg.set_rotational_states([0, 2, 10]) is cfg
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    levels = np.array(handle.get_levels())
    assert levels.shape == (6, 4)  # 2 curves x 3 J
    assert np.array_equal(levels[:3], levels[3:])
    assert np.all(np.diff(levels[:3], axis=0) > 0)
    N = 16500
    h = W.grid_h(0.0, 10.0, N)
    V = oracle.morse(5500.0, 0.6, 10.0, 0.0, 10.0, N)
    s = oracle.scale(87.62, 87.62, h)
    for k, J in enumerate((0, 2, 10)):
        VJ = oracle.centrifugal(V, s, 0.0, h, J)
        F, _, _, vmin = oracle_d.prep(VJ, s)
        lev_o, *_ = oracle_d.solve_levels(F, s, vmin, VJ[-1] - 0.1, 1024, 0, 3, 256, 1e-12, 16)
        assert np.array_equal(levels[k].view(np.uint64), lev_o.view(np.uint64)), J


@pytest.mark.gpu
def test_potential_tables_through_python_api(gpu_mod, oracle, oracle_d):
    """Additive surface: curves handed over as a numpy array [curve][point]; float64 levels carry the
    bits of the oracle run on the same tables (Morse + Lennard-Jones in one batch, as in config 4)."""
    N, rmin, rmax = 12000, 0.4, 10.0
    V = np.stack([W.morse(5500.0, 2.2, 1.6, rmin, rmax, N), W.lj(5200.0, 2.3, rmin, rmax, N)])
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = interface.get_task_configurator("float64")
    cfg.set_hardware_config(potential_buffer_size=N, group_size=1024, allocation_block_size=1 << 20)
    assert cfg.set_potential_tables(V, rmin, rmax) is cfg and not cfg.is_configured()
    cfg.set_vibwa_algorithm(mass_atom_0=20.0, mass_atom_1=20.0, integration_step=0.1,
                            min_distance_to_asymptote=1.0, min_level=0, max_level=6)
    assert cfg.is_configured()
    with pytest.raises(RuntimeError):
        cfg.set_potential_tables(V[0], rmin, rmax)  # 1-D
    with pytest.raises(RuntimeError):
        cfg.set_potential_tables(V, rmax, rmin)
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    levels = np.array(handle.get_levels())
    s = oracle.scale(20.0, 20.0, W.grid_h(rmin, rmax, N))
    for c in range(2):
        F, _, _, vmin = oracle_d.prep(V[c], s)
        ref, *_ = oracle_d.solve_levels(F, s, vmin, V[c][-1] - 1.0, 1024, 0, 6, 256, 1e-12, 16)
        assert np.array_equal(levels[c].view(np.uint64), ref.view(np.uint64)), c


@pytest.mark.gpu
def test_concurrent_tasks_on_one_device(gpu_mod):
    """The reference documents concurrent tasks on one device as unsupported
    (python/epseon_backend/device/gpu/_libepseon_gpu.pyi:156-158).  Here every task owns its worker
    thread, CUDA stream and device buffers, so tasks submitted together all finish with the same bits."""
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    handles = [interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator("float64"), max_level=9))
               for _ in range(4)]
    for h in handles:
        h.wait()
        assert h.is_done() and not h.has_failed(), h.get_status_message()
    ref = np.array(handles[0].get_levels())
    assert np.all(np.isfinite(ref))
    for h in handles[1:]:
        assert np.array_equal(np.array(h.get_levels()).view(np.uint64), ref.view(np.uint64))


@pytest.mark.gpu
def test_example_script_runs():
    """examples/morse_levels.py: the reference example's call sequence plus the additive accessors."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(ROOT / "examples" / "morse_levels.py")], capture_output=True, text=True,
                       cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert "curve 1 J=1" in r.stdout and "rotational constant" in r.stdout


@pytest.mark.gpu
def test_batch_task_uses_fewer_points_per_round(gpu_mod, oracle, oracle_d):
    """Hundreds of curves in one task: the search runs with fewer points per level per round
    (get_search_parameters), and the levels still carry the oracle's bits for those parameters."""
    N, nC = 3000, 640
    rng = np.random.default_rng(12)
    De = 5500.0 * (1 + 0.05 * (2 * rng.random(nC) - 1))
    interface = gpu_mod.EpseonComputeContext.create().get_device_interface(0)
    cfg = interface.get_task_configurator("float64")
    cfg.set_hardware_config(potential_buffer_size=N, group_size=1024, allocation_block_size=1 << 20)
    cfg.set_morse_potential([gpu_mod.MorsePotentialConfig(dissociation_energy=float(d), equilibrium_bond_distance=2.2,
                                                          well_width=1.6, min_r=0.4, max_r=10.0, point_count=N) for d in De])
    cfg.set_vibwa_algorithm(mass_atom_0=20.0, mass_atom_1=20.0, integration_step=0.1,
                            min_distance_to_asymptote=1.0, min_level=0, max_level=7)
    handle = interface.submit_task(cfg)
    handle.wait()
    assert not handle.has_failed(), handle.get_status_message()
    n_coarse, M, rounds, tol = handle.get_search_parameters()
    assert (n_coarse, M, tol) == (1024, 16, 1e-12) and rounds >= 16
    levels = np.array(handle.get_levels())
    assert levels.shape == (nC, 8) and np.all(np.isfinite(levels))
    s = oracle.scale(20.0, 20.0, W.grid_h(0.4, 10.0, N))
    for c in (0, 1, 317, 639):
        V = oracle.morse(float(De[c]), 2.2, 1.6, 0.4, 10.0, N)
        F, _, _, vmin = oracle_d.prep(V, s)
        ref, *_ = oracle_d.solve_levels(F, s, vmin, V[-1] - 1.0, n_coarse, 0, 7, M, tol, rounds)
        assert np.array_equal(levels[c].view(np.uint64), ref.view(np.uint64)), c
    # a small task keeps the 256-point rounds
    small = interface.submit_task(_configure_task(gpu_mod, interface.get_task_configurator("float64"), max_level=3))
    small.wait()
    assert small.get_search_parameters()[:3] == (1024, 256, 16)
