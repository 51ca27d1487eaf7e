"""World-size-2 worker of tests/test_multi.py (CPU, gloo).  The CPU oracle stands in for the
device behind the same ``solve_levels_grid`` interface, so the host-side sharding / merging logic
of epseon_backend_b200.multi is exercised end to end without a GPU."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch.distributed as dist  # noqa: E402

from epseon_backend_b200 import multi  # noqa: E402
from oracle import Oracle  # noqa: E402
from tests import workloads as W  # noqa: E402


class GlooComm:
    """The `comm` protocol of epseon_backend_b200.multi (world, rank, gather) over torch.distributed /
    gloo -- test plumbing; the product's carriers are eps_group_* and eps_mailbox_* (CUDA)."""

    def __init__(self):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()

    def gather(self, arr):
        import torch

        a = np.ascontiguousarray(arr)
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy())
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        return [o.numpy().view(a.dtype).reshape(a.shape) for o in outs]


class OracleSolver:
    """cabi.Context look-alike over the oracle (tests only)."""

    def __init__(self, V, s):
        self.orc = Oracle()
        self.V = np.atleast_2d(V)
        self.s = s
        self.n_curves = self.V.shape[0]
        self.F = [self.orc.prep(v, s)[0] for v in self.V]

    def solve_levels_grid(self, E0, dE, j0, n_coarse, v_min, v_max, M, rel_tol, max_rounds):
        nlev = v_max - v_min + 1
        lev = np.empty((self.n_curves, nlev))
        wid = np.empty((self.n_curves, nlev))
        nl = np.empty(self.n_curves, dtype=np.uint32)
        nf = np.empty(self.n_curves, dtype=np.uint32)
        for c in range(self.n_curves):
            lev[c], wid[c], nl[c], nf[c], _, _ = self.orc.solve_levels_grid(self.F[c], self.s, E0[c], dE[c], j0, n_coarse,
                                                                          v_min, v_max, M, rel_tol, max_rounds)
        return lev, wid, nl, nf


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = GlooComm()
    # --- energy-range sharding of one solve == single-process solve, bit for bit
    w = W.c4(nC=3, N=3000, nE=257)
    solver = OracleSolver(w["V"], w["s"])
    for n_coarse in (257, 64, 2 * world, 3):
        lev, wid, nb = multi.solve_levels_energy_sharded(solver, comm, w["E_lo"], w["E_hi"], n_coarse, 0, 9, 32, 1e-12, 10)
        for c in range(3):
            ref = solver.orc.solve_levels(solver.F[c], w["s"], w["E_lo"][c], w["E_hi"][c], n_coarse, 0, 9, 32, 1e-12, 10)
            assert np.array_equal(lev[c].view(np.uint64), ref[0].view(np.uint64)), (rank, n_coarse, c, lev[c], ref[0])
            assert np.array_equal(wid[c].view(np.uint64), ref[1].view(np.uint64))
            assert nb[c] == ref[2]
    # --- curve sharding: the ranks' blocks tile the batch
    w = W.c4(nC=7, N=2000, nE=128)
    sl = multi.curve_shard(7, world, rank)
    mine = OracleSolver(w["V"][sl], w["s"])
    lo, hi = w["E_lo"][sl], w["E_hi"][sl]
    dE = multi.global_step(lo, hi, 128)
    lev, *_ = mine.solve_levels_grid(lo, dE, 0, 128, 0, 5, 32, 1e-12, 10)
    pad = np.full((4, 6), np.nan)  # equal shapes for the gather (7 curves over 2 ranks: 4 + 3)
    pad[: lev.shape[0]] = lev
    parts = comm.gather(pad)
    full = np.concatenate([parts[r][: (multi.curve_shard(7, world, r).stop - multi.curve_shard(7, world, r).start)]
                           for r in range(world)])
    whole = OracleSolver(w["V"], w["s"])
    ref, *_ = whole.solve_levels_grid(w["E_lo"], multi.global_step(w["E_lo"], w["E_hi"], 128), 0, 128, 0, 5, 32, 1e-12, 10)
    assert np.array_equal(full.view(np.uint64), ref.view(np.uint64))
    dist.barrier()
    if rank == 0:
        print("DIST_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
