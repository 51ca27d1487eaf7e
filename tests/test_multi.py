"""N > 1 host logic (epseon_backend_b200.multi): shard arithmetic on CPU, a world-size-2 gloo run,
and -- on the GPU box -- the energy-sharded solve through the C ABI against the single-call solve."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from epseon_backend_b200 import multi
from tests import workloads as W

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n,world", [(4096, 8), (7, 2), (3, 8), (17, 4), (1, 3)])
def test_curve_shard_tiles(n, world):
    parts = [multi.curve_shard(n, world, r) for r in range(world)]
    assert parts[0].start == 0 and parts[-1].stop == n
    assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
    sizes = [p.stop - p.start for p in parts]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_coarse,world", [(65536, 8), (1 << 24, 8), (10, 4), (3, 8), (2, 2)])
def test_energy_shard_covers_every_interval_once(n_coarse, world):
    seen = np.zeros(n_coarse - 1, dtype=np.int32)  # grid intervals [j, j+1]
    for r in range(world):
        j0, n = multi.energy_shard(n_coarse, world, r)
        if n:
            assert n >= 2 and j0 + n <= n_coarse
            seen[j0:j0 + n - 1] += 1
    assert np.all(seen == 1)


def test_merge_levels_rejects_double_ownership():
    a = np.array([[1.0, np.nan]])
    b = np.array([[np.nan, 2.0]])
    assert np.array_equal(multi.merge_levels([a, b]), [[1.0, 2.0]])
    with pytest.raises(RuntimeError):
        multi.merge_levels([a, a])


def test_energy_sharded_solve_single_process(oracle):
    """world = 1 path and the slice arithmetic against the oracle (no process group)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from tests.dist_worker import OracleSolver

    w = W.c1()
    solver = OracleSolver(w["V"], w["s"])
    ref = oracle.solve_levels(solver.F[0], w["s"], w["E_lo"], w["E_hi"], 513, 0, 16, 48, 1e-12, 10)
    # emulate 4 ranks by hand: slices of the global grid, merged
    dE = multi.global_step(w["E_lo"], w["E_hi"], 513)
    parts, lasts = [], []
    for r in range(4):
        j0, n = multi.energy_shard(513, 4, r)
        lev, wid, nl, nf = solver.solve_levels_grid(np.array([w["E_lo"]]), np.array([dE]), j0, n, 0, 16, 48, 1e-12, 10)
        parts.append(lev)
        lasts.append(nl)
    merged = multi.merge_levels(parts)
    assert np.array_equal(merged[0].view(np.uint64), ref[0].view(np.uint64))
    assert lasts[-1][0] == ref[2]


def test_world_size_2_gloo():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(ROOT / "tests" / "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK 2" in r.stdout


@pytest.mark.gpu
def test_gpu_energy_sliced_solve_equals_whole(gpu_ctx):
    """Through the C ABI: four slices of the global grid (eps_solve_levels_grid) merged == one
    eps_solve_levels call, bit for bit; eps_sweep_grid reproduces the global grid's node counts."""
    w = W.c1()
    gpu_ctx.set_potentials(w["V"], w["s"])
    whole, wid, nb = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 2049, 0, 17, 64, 1e-12, 10)
    dE = multi.global_step(w["E_lo"], w["E_hi"], 2049)
    parts = []
    for r in range(4):
        j0, n = multi.energy_shard(2049, 4, r)
        lev, _, nl, nf = gpu_ctx.solve_levels_grid(w["E_lo"], dE, j0, n, 0, 17, 64, 1e-12, 10)
        parts.append(lev)
    merged = multi.merge_levels(parts)
    assert np.array_equal(merged.view(np.uint64), whole.view(np.uint64))
    assert nl[0] == nb[0] == 17
    n_all, _, _ = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 2049, tails=False)
    j0, n = multi.energy_shard(2049, 4, 2)
    n_part, _, _ = gpu_ctx.sweep_grid(w["E_lo"], dE, j0, n, tails=False)
    assert np.array_equal(n_part[0], n_all[0][j0:j0 + n])


@pytest.mark.gpu
def test_in_process_fanout_over_reference_api():
    """One task per visible device through the reference's own classes (threads, no torchrun):
    the concatenated levels equal a single-device task over the whole batch, bit for bit."""
    import __graft_entry__ as ge

    ge.build()
    import epseon_backend.device.gpu._libepseon_gpu as m

    rng = np.random.default_rng(7)
    cfgs = [m.MorsePotentialConfig(dissociation_energy=5500.0 * (1 + 0.05 * rng.random()), equilibrium_bond_distance=2.2,
                                   well_width=1.6, min_r=0.4, max_r=10.0, point_count=8000) for _ in range(11)]
    hw = dict(potential_buffer_size=8000, group_size=1024, allocation_block_size=1 << 20)
    alg = dict(mass_atom_0=20.0, mass_atom_1=20.0, integration_step=0.1, min_distance_to_asymptote=1.0,
               min_level=0, max_level=4)
    ids = [d.device_properties.device_id for d in m.EpseonComputeContext.create().get_physical_device_info()]
    lev_all, cnt_all, handles = multi.solve_morse_batch_all_devices(m, cfgs, hw, alg)
    assert len(handles) == min(len(ids), len(cfgs)) and len(lev_all) == len(cfgs) == len(cnt_all)
    lev_one, cnt_one, _ = multi.solve_morse_batch_all_devices(m, cfgs, hw, alg, device_ids=ids[:1])
    assert np.array_equal(np.array(lev_all).view(np.uint64), np.array(lev_one).view(np.uint64))
    assert cnt_all == cnt_one
    assert np.all(np.diff(np.array(lev_all), axis=1) > 0)
