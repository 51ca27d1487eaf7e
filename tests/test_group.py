"""The torch-free multi-device carriers of the C ABI: eps_group_* (one process, a host thread and a
context per device, levels gathered on the first device with peer copies) and eps_mailbox_* (one
process per device, rank 0's device buffer shared through a CUDA IPC handle).  The GPU test box has
one device: groups are built over (0, 0, 0) -- three contexts and three host threads on one GPU --
which exercises the same sharding, peer-copy and merge code; the driver's 8-GPU bench exercises the
real thing."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from epseon_backend_b200 import multi
from tests import workloads as W

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    return cabi


@pytest.mark.parametrize("n_dev", [1, 2, 3])
@pytest.mark.parametrize("form", [0, 1])
def test_group_curve_sharded_equals_one_device(cabi, gpu_ctx, n_dev, form):
    w = W.c4(nC=11, N=3000, nE=300)
    gpu_ctx.set_option(gpu_ctx.OPT_FORM, form)
    gpu_ctx.set_potentials(w["V"], w["s"])
    ref = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 300, 0, 7, 32, 1e-11, 12)
    ref_nodes, _, _ = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 300, tails=False)
    gpu_ctx.set_option(gpu_ctx.OPT_FORM, 0)
    with cabi.Group([0] * n_dev) as g:
        assert g.size == n_dev
        g.set_option(cabi.Context.OPT_FORM, form)
        g.set_potentials(w["V"], w["s"], cabi.SHARD_CURVES)
        lev, wid, nb = g.solve_levels(w["E_lo"], w["E_hi"], 300, 0, 7, 32, 1e-11, 12)
        assert _same_bits(lev, ref[0]) and _same_bits(wid, ref[1]) and np.array_equal(nb, ref[2])
        assert g.last_ms() > 0.0
        assert np.array_equal(g.sweep_uniform(w["E_lo"], w["E_hi"], 300), ref_nodes)


@pytest.mark.parametrize("n_dev", [2, 3, 5])
def test_group_energy_sharded_equals_one_device(cabi, gpu_ctx, n_dev):
    """One curve (C1) and a few curves: the slices of the global grid give the single-device bits."""
    for w, vmax in ((W.c1(), 17), (W.c4(nC=3, N=3000, nE=257), 9)):
        gpu_ctx.set_potentials(w["V"], w["s"])
        with cabi.Group([0] * n_dev) as g:
            g.set_potentials(w["V"], w["s"], cabi.SHARD_ENERGY)
            for n_coarse in (2049, 257, 4):
                ref = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], n_coarse, 0, vmax, 64, 1e-12, 10)
                lev, wid, nb = g.solve_levels(w["E_lo"], w["E_hi"], n_coarse, 0, vmax, 64, 1e-12, 10)
                assert _same_bits(lev, ref[0]) and _same_bits(wid, ref[1]) and np.array_equal(nb, ref[2]), (n_dev, n_coarse)
            ref_nodes, _, _ = gpu_ctx.sweep_uniform(w["E_lo"], w["E_hi"], 5000, tails=False)
            assert np.array_equal(g.sweep_uniform(w["E_lo"], w["E_hi"], 5000), ref_nodes)


def test_group_errors_are_reported_not_fatal(cabi):
    w = W.c1()
    with cabi.Group([0, 0]) as g:
        with pytest.raises(cabi.EpsError) as e:
            g.solve_levels(0.0, 1.0, 64, 0, 3, 16)
        assert e.value.code == 4  # EPS_ERR_STATE: no potentials yet
        g.set_potentials(w["V"], w["s"], cabi.SHARD_ENERGY)
        with pytest.raises(cabi.EpsError) as e:
            g.solve_levels(w["E_lo"], 1e9, 64, 0, 3, 16)  # far outside the validity window
        assert e.value.code == 3 and "device" in str(e.value)
        lev, _, nb = g.solve_levels(w["E_lo"], w["E_hi"], 512, 0, 16, 64, 1e-12, 10)  # the group still works
        assert nb[0] == 17 and np.all(np.isfinite(lev))
    with pytest.raises(cabi.EpsError):
        cabi.Group([0, 99])


def test_all_devices_helper(cabi, gpu_ctx):
    w = W.c4(nC=6, N=2500, nE=200)
    gpu_ctx.set_potentials(w["V"], w["s"])
    ref = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 200, 0, 5, 64, 1e-11, 10)
    for shard in ("curves", "energy"):
        lev, wid, nb, ms = multi.solve_levels_all_devices(w["V"], w["s"], w["E_lo"], w["E_hi"], 200, 0, 5, 64, 1e-11, 10, shard=shard)
        assert _same_bits(lev, ref[0]) and np.array_equal(nb, ref[2]) and ms > 0


RANK1 = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from epseon_backend_b200 import cabi, multi
from tests import workloads as W
handle = bytes.fromhex(sys.argv[1])
w = W.c1()
ctx = cabi.Context(0)
ctx.set_potentials(w["V"], w["s"])
comm = multi.MailboxComm(ctx, 2, 1, lambda payload: [handle, b""])
assert multi.solve_levels_energy_sharded(ctx, comm, w["E_lo"], w["E_hi"], 1025, 0, 16, 64, 1e-12, 10) is None
ctx.solve_levels(w["E_lo"] + 1.0, w["E_hi"], 512, 0, 16, 64, 1e-12, 10)   # rank 1's own problem
assert comm.gather_levels(1, 17) is None
comm.close(); ctx.close()
print("RANK1_OK")
"""


def test_mailbox_across_processes(cabi, gpu_ctx, tmp_path):
    """Two PROCESSES on device 0: the energy-sharded level search gathered through the IPC mailbox
    equals the single-process search, and the device-to-device level post arrives intact."""
    w = W.c1()
    gpu_ctx.set_potentials(w["V"], w["s"])
    whole = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 1025, 0, 16, 64, 1e-12, 10)
    other = gpu_ctx.solve_levels(w["E_lo"] + 1.0, w["E_hi"], 512, 0, 16, 64, 1e-12, 10)
    proc = {}

    def exchange(payload):  # rank 0's side of the rendezvous: start rank 1 with the handle on its command line
        proc["p"] = subprocess.Popen([sys.executable, "-c", RANK1.format(root=str(ROOT)), payload.hex()],
                                     stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(ROOT))
        return [payload, b""]

    comm = multi.MailboxComm(gpu_ctx, 2, 0, exchange)
    try:
        lev, wid, nb = multi.solve_levels_energy_sharded(gpu_ctx, comm, w["E_lo"], w["E_hi"], 1025, 0, 16, 64, 1e-12, 10)
        assert _same_bits(lev, whole[0]) and _same_bits(wid, whole[1]) and np.array_equal(nb, whole[2])
        mine = gpu_ctx.solve_levels(w["E_lo"], w["E_hi"], 256, 0, 16, 64, 1e-12, 10)
        parts = comm.gather_levels(1, 17)
        assert _same_bits(parts[0][0], mine[0]) and _same_bits(parts[0][1], mine[1])
        assert _same_bits(parts[1][0], other[0]) and _same_bits(parts[1][1], other[1])
        out, err = proc["p"].communicate(timeout=120)
        assert proc["p"].returncode == 0 and "RANK1_OK" in out, err[-3000:]
    finally:
        comm.close()
        if proc.get("p") and proc["p"].poll() is None:
            proc["p"].kill()


def test_energy_shard_through_reference_api(cabi):
    """C1-like single-curve problem fanned out by ENERGY RANGE over three tasks through the
    reference's own classes (set_energy_shard, additive): merged levels == the unsharded task."""
    import epseon_backend.device.gpu._libepseon_gpu as m

    cfgs = [m.MorsePotentialConfig(dissociation_energy=38267.0, equilibrium_bond_distance=0.7414, well_width=1.9426,
                                   min_r=0.2, max_r=10.0, point_count=10000),
            m.MorsePotentialConfig(dissociation_energy=30000.0, equilibrium_bond_distance=0.8, well_width=1.8,
                                   min_r=0.2, max_r=10.0, point_count=10000)]
    hw = dict(potential_buffer_size=10000, group_size=4096, allocation_block_size=1 << 20)
    alg = dict(mass_atom_0=1.00783, mass_atom_1=1.00783, integration_step=0.1, min_distance_to_asymptote=1.0,
               min_level=0, max_level=16)
    one, cnt_one, _ = multi.solve_morse_batch_all_devices(m, cfgs, hw, alg, device_ids=[0])
    for world in (2, 3):
        lev, cnt, handles = multi.solve_morse_batch_all_devices(m, cfgs, hw, alg, device_ids=[0] * world, shard="energy")
        assert len(handles) == world
        assert _same_bits(np.array(lev), np.array(one)) and list(cnt) == list(cnt_one)
