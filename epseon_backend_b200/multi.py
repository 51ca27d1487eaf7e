"""Multi-GPU fan-out of the hot path.  No torch, no NCCL: the path has no data-path collective.

The path shards in two natural ways:

* by potential curve -- every device owns a contiguous block of curves (``curve_shard``);
* by energy range    -- every device sweeps a slice of ONE global uniform energy grid
  (``energy_shard``).  Slices of neighbouring ranks share one grid point, so each bracket
  ``[E_{j-1}, E_j]`` of the global grid lies in exactly one slice; the slice is expressed as
  (E0, dE, j0) of the global grid (``eps_solve_levels_grid`` / ``eps_sweep_grid``), which
  reproduces the global energies bit for bit.  The merged result therefore equals the
  single-device result bit for bit.

Only the small results (level energies, a few node counts) travel.  Three carriers:

* one process, all devices: ``cabi.Group`` (``eps_group_*``: a host thread per device, levels
  gathered on the first device with ``cudaMemcpyPeerAsync``) -- ``solve_levels_all_devices``;
* one process per device (torchrun & co.): ``MailboxComm`` (``eps_mailbox_*``: rank 0's device
  buffer shared through a CUDA IPC handle, peer writes) -- ``solve_levels_energy_sharded``;
* the reference's own classes: one ``TaskHandle`` per device -- ``solve_morse_batch_all_devices``.

A "comm" is any object with ``world``, ``rank`` and ``gather(arr) -> list[np.ndarray] | None``
(the list at least on rank 0).  The CPU tests pass a gloo-backed one with the oracle as solver.
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


class LocalComm:
    """world = 1."""

    world, rank = 1, 0

    def gather(self, arr: np.ndarray):
        return [np.asarray(arr)]


class MailboxComm:
    """One process per GPU: small results gathered into rank 0's device memory with peer writes
    (``eps_mailbox_*``).  ``exchange(payload: bytes) -> list[bytes]`` is the launcher's rendezvous
    channel (any all-gather of 64 bytes: a torch.distributed object gather in ``bench.py``, a file,
    MPI ...); it is used once, to hand rank 0's CUDA IPC handle to the other ranks."""

    def __init__(self, ctx, world: int, rank: int, exchange, max_bytes: int = 1 << 20):
        from . import cabi

        self.ctx, self.world, self.rank, self.seq = ctx, world, rank, 0
        if rank == 0:
            self.box = cabi.Mailbox.create(ctx, world, max_bytes)
            handles = exchange(self.box.handle)
        else:
            handles = exchange(b"")
            self.box = cabi.Mailbox.open(ctx, handles[0], world, rank, max_bytes)

    def gather(self, arr: np.ndarray):
        a = np.ascontiguousarray(arr)
        self.seq += 1
        self.box.post(a, self.seq)
        if self.rank != 0:
            return None
        raw = self.box.collect(self.seq, a.nbytes)
        return [raw[r].view(a.dtype).reshape(a.shape).copy() for r in range(self.world)]

    def gather_levels(self, n_curves: int, n_levels: int):
        """The levels / widths of every rank's last ``solve_levels*``, device to device (no host
        staging on the senders) -> on rank 0: list of (levels[nC, nlev], widths[nC, nlev])."""
        self.seq += 1
        self.box.post_levels(self.seq)
        if self.rank != 0:
            return None
        n = n_curves * n_levels
        raw = self.box.collect(self.seq, 16 * n)
        out = []
        for r in range(self.world):
            d = raw[r].view(np.float64)
            out.append((d[:n].reshape(n_curves, n_levels).copy(), d[n:].reshape(n_curves, n_levels).copy()))
        return out

    def close(self):
        self.box.close()


def solve_levels_energy_sharded(solver, comm, E_lo, E_hi, n_coarse: int, v_min: int, v_max: int,
                                refine_points: int, rel_tol: float = 1e-12, max_rounds: int = 8):
    """Locate levels v_min..v_max of the solver's resident curves with the coarse sweep split by
    energy range over the ranks of ``comm``.  Returns (levels[nC, nlev], widths, n_below[nC]) --
    bit-identical to ``solver.solve_levels(E_lo, E_hi, n_coarse, ...)`` on one device -- wherever
    ``comm.gather`` returns the parts (rank 0 at least), None elsewhere.  ``solver`` is anything with
    ``solve_levels_grid`` / ``n_curves`` (``cabi.Context`` in production)."""
    comm = comm if comm is not None else LocalComm()
    world, rank = comm.world, comm.rank
    nC = solver.n_curves
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(E_lo, dtype=np.float64), (nC,)))
    dE = np.ascontiguousarray(np.broadcast_to(global_step(E_lo, E_hi, n_coarse), (nC,)))
    nlev = v_max - v_min + 1
    j0, n_local = energy_shard(n_coarse, world, rank)
    if n_local >= 2:
        lev, wid, n_last, _ = solver.solve_levels_grid(lo, dE, j0, n_local, v_min, v_max, refine_points, rel_tol, max_rounds)
    else:  # more ranks than grid intervals
        lev = np.full((nC, nlev), np.nan)
        wid = np.full((nC, nlev), np.nan)
        n_last = np.zeros(nC, dtype=np.uint32)
    packed = np.concatenate([lev.reshape(-1), wid.reshape(-1), np.asarray(n_last, dtype=np.float64)])
    parts = comm.gather(packed)
    if parts is None:
        return None
    n = nC * nlev
    levs = [p[:n].reshape(nC, nlev) for p in parts]
    wids = [p[n:2 * n].reshape(nC, nlev) for p in parts]
    owners = [r for r in range(world) if energy_shard(n_coarse, world, r)[1] >= 2]
    return merge_levels(levs), merge_levels(wids), parts[owners[-1]][2 * n:].astype(np.uint32)


def solve_levels_all_devices(V, scale, E_lo, E_hi, n_coarse: int, v_min: int, v_max: int, refine_points: int,
                             rel_tol: float = 1e-12, max_rounds: int = 8, shard: str = "curves", devices=None,
                             accurate: bool = False):
    """All visible devices from ONE process through ``eps_group_*`` (torch-free): the job is cut by
    curve (``shard="curves"``) or by energy range (``"energy"``, for few curves).  Returns
    (levels[nC, nlev], widths, n_below[nC], max-over-devices CUDA-event milliseconds); bit-identical to
    one device."""
    from . import cabi

    if devices is None:
        devices = list(range(cabi.device_count()))
    V = np.atleast_2d(V)
    if shard == "curves":
        devices = devices[: max(1, min(len(devices), V.shape[0]))]
    with cabi.Group(devices) as g:
        if accurate:
            g.set_option(cabi.Context.OPT_FORM, 1)
        g.set_potentials(V, scale, cabi.SHARD_CURVES if shard == "curves" else cabi.SHARD_ENERGY)
        lev, wid, nb = g.solve_levels(E_lo, E_hi, n_coarse, v_min, v_max, refine_points, rel_tol, max_rounds)
        return lev, wid, nb, g.last_ms()


def solve_morse_batch_all_devices(gpu_mod, configurations, hardware: dict, algorithm: dict, precision: str = "float64",
                                  device_ids=None, rotational_states=None, shard: str = "curves"):
    """In-process fan-out over the reference's own API ("one
    ComputeDeviceInterface per GPU"): one task per device is configured and submitted -- every
    ``TaskHandle`` owns a worker thread and its own CUDA context, so the devices run concurrently.

    ``shard="curves"``: the list of ``MorsePotentialConfig`` is cut into contiguous blocks
    (``curve_shard``) and the per-curve results are concatenated in the original order.
    ``shard="energy"``: every device gets ALL curves and a slice of the coarse energy grid
    (``set_energy_shard(rank, world)``, additive); each level is located by exactly one device and
    the per-device level arrays are merged.  Either way the result equals a single-device task bit
    for bit.

    ``hardware`` / ``algorithm`` are the keyword arguments of ``set_hardware_config`` /
    ``set_vibwa_algorithm``.  Returns (levels [curve][level], level_counts [curve], handles).
    """
    ctx = gpu_mod.EpseonComputeContext.create()
    if device_ids is None:
        device_ids = [d.device_properties.device_id for d in ctx.get_physical_device_info()]
    device_ids = list(device_ids)
    if shard == "curves":
        device_ids = device_ids[: max(1, len(configurations))]
    handles = []
    for r, dev in enumerate(device_ids):
        sl = curve_shard(len(configurations), len(device_ids), r) if shard == "curves" else slice(0, len(configurations))
        if sl.stop == sl.start:
            continue
        interface = ctx.get_device_interface(dev)
        cfg = (interface.get_task_configurator(precision).set_hardware_config(**hardware)
               .set_morse_potential(list(configurations[sl])).set_vibwa_algorithm(**algorithm))
        if rotational_states is not None:
            cfg.set_rotational_states(list(rotational_states))
        if shard == "energy":
            cfg.set_energy_shard(r, len(device_ids))
        handles.append(interface.submit_task(cfg))
    levels, counts = [], []
    for h in handles:
        h.wait()
        if h.has_failed():
            raise RuntimeError(h.get_status_message())
    if shard == "curves":
        for h in handles:
            levels.extend(h.get_levels())
            counts.extend(h.get_level_counts())
        return levels, counts, handles
    merged = merge_levels([np.array(h.get_levels(), dtype=np.float64) for h in handles])
    return merged.tolist(), list(handles[-1].get_level_counts()), handles
