"""Multi-GPU fan-out of the hot path: one process per GPU, no data-path collective.

The path shards in two natural ways:

* by potential curve -- every rank owns a contiguous block of curves (``curve_shard``);
* by energy range    -- every rank sweeps a slice of ONE global uniform energy grid
  (``energy_shard``).  Slices of neighbouring ranks share one grid point, so each bracket
  ``[E_{j-1}, E_j]`` of the global grid lies in exactly one rank's slice; the slice is expressed
  as (E0, dE, j0) of the global grid (``eps_solve_levels_grid`` / ``eps_sweep_grid``), which
  reproduces the global energies bit for bit.  The merged result therefore equals the
  single-device result bit for bit.

Only the small results (level energies, a few node counts) travel: gathered with
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests).  The solver argument is
anything with ``solve_levels_grid`` / ``n_curves`` -- ``cabi.Context`` in production.
"""
from __future__ import annotations

import numpy as np


def curve_shard(n_curves: int, world: int, rank: int) -> slice:
    """Contiguous block of curves of ``rank``; the remainder goes to the first ranks."""
    base, rem = divmod(n_curves, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def energy_shard(n_coarse: int, world: int, rank: int) -> tuple[int, int]:
    """(j0, n_local): rank's slice j0 .. j0+n_local-1 of the global grid j = 0 .. n_coarse-1.

    The n_coarse-1 grid intervals are split contiguously; a rank owning intervals [a, b) sweeps the
    points a..b, i.e. shares point b with its right neighbour.  Ranks left without an interval get
    n_local = 0.
    """
    sl = curve_shard(n_coarse - 1, world, rank)
    n_int = sl.stop - sl.start
    return (sl.start, n_int + 1) if n_int > 0 else (sl.start, 0)


def global_step(E_lo, E_hi, n_coarse: int):
    """dE of the global grid, with the operations the device uses: (E_hi - E_lo) / (n_coarse - 1)."""
    lo = np.asarray(E_lo, dtype=np.float64)
    hi = np.asarray(E_hi, dtype=np.float64)
    return (hi - lo) / np.float64(n_coarse - 1)


def merge_levels(parts: list[np.ndarray]) -> np.ndarray:
    """Union of the ranks' level arrays [nC, nlev]: each level is finite on at most one rank."""
    out = np.full_like(parts[0], np.nan)
    taken = np.zeros(parts[0].shape, dtype=bool)
    for p in parts:
        ok = np.isfinite(p)
        if np.any(ok & taken):
            raise RuntimeError("a level was located by two ranks: energy slices overlap by more than one point")
        out[ok] = p[ok]
        taken |= ok
    return out


def _device_for(dist):
    import torch

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_array(dist, arr: np.ndarray) -> list[np.ndarray]:
    """all_gather of equally-shaped small numpy arrays (bit-preserving: moved as raw int64/uint8)."""
    import torch

    a = np.ascontiguousarray(arr)
    t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(_device_for(dist))
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in outs]


def solve_levels_energy_sharded(solver, dist, E_lo, E_hi, n_coarse: int, v_min: int, v_max: int,
                                refine_points: int, rel_tol: float = 1e-12, max_rounds: int = 8):
    """Locate levels v_min..v_max of the solver's resident curves with the coarse sweep split by
    energy range over the ranks.  Every rank returns (levels[nC, nlev], widths, n_below[nC]),
    bit-identical to ``solver.solve_levels(E_lo, E_hi, n_coarse, ...)`` on one device."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist is not None else (1, 0)
    nC = solver.n_curves
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(E_lo, dtype=np.float64), (nC,)))
    dE = np.ascontiguousarray(np.broadcast_to(global_step(E_lo, E_hi, n_coarse), (nC,)))
    nlev = v_max - v_min + 1
    j0, n_local = energy_shard(n_coarse, world, rank)
    if n_local >= 2:
        lev, wid, n_last, n_first = solver.solve_levels_grid(lo, dE, j0, n_local, v_min, v_max, refine_points,
                                                             rel_tol, max_rounds)
    else:  # more ranks than grid intervals
        lev = np.full((nC, nlev), np.nan)
        wid = np.full((nC, nlev), np.nan)
        n_last = np.zeros(nC, dtype=np.uint32)
    if dist is None:
        return lev, wid, n_last
    levs = all_gather_array(dist, lev)
    wids = all_gather_array(dist, wid)
    lasts = all_gather_array(dist, np.asarray(n_last, dtype=np.uint32))
    owners = [r for r in range(world) if energy_shard(n_coarse, world, r)[1] >= 2]
    return merge_levels(levs), merge_levels(wids), lasts[owners[-1]]


def solve_morse_batch_all_devices(gpu_mod, configurations, hardware: dict, algorithm: dict, precision: str = "float64",
                                  device_ids=None, rotational_states=None):
    """In-process fan-out over the reference's own API ("one
    ComputeDeviceInterface per GPU"): the list of ``MorsePotentialConfig`` is cut into contiguous
    blocks (``curve_shard``), one task per device is configured and submitted -- every
    ``TaskHandle`` owns a worker thread and its own CUDA context, so the devices run concurrently
    -- and the per-curve results are concatenated in the original order.

    ``hardware`` / ``algorithm`` are the keyword arguments of ``set_hardware_config`` /
    ``set_vibwa_algorithm``.  Returns (levels [curve][level], level_counts [curve], handles).
    """
    ctx = gpu_mod.EpseonComputeContext.create()
    if device_ids is None:
        device_ids = [d.device_properties.device_id for d in ctx.get_physical_device_info()]
    device_ids = list(device_ids)[: max(1, len(configurations))]
    handles = []
    for r, dev in enumerate(device_ids):
        sl = curve_shard(len(configurations), len(device_ids), r)
        if sl.stop == sl.start:
            continue
        interface = ctx.get_device_interface(dev)
        cfg = (interface.get_task_configurator(precision).set_hardware_config(**hardware)
               .set_morse_potential(list(configurations[sl])).set_vibwa_algorithm(**algorithm))
        if rotational_states is not None:
            cfg.set_rotational_states(list(rotational_states))
        handles.append(interface.submit_task(cfg))
    levels, counts = [], []
    for h in handles:
        h.wait()
        if h.has_failed():
            raise RuntimeError(h.get_status_message())
        levels.extend(h.get_levels())
        counts.extend(h.get_level_counts())
    return levels, counts, handles
