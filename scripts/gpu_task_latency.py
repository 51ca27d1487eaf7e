"""Wall-clock cost of one task through the drop-in Python API (submit_task .. wait), against the
device time of its level solve: what the per-task context set-up costs."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from epseon_backend.device.gpu._libepseon_gpu import EpseonComputeContext, MorsePotentialConfig  # noqa: E402

interface = EpseonComputeContext.create().get_device_interface(0)


def task(levels=13):
    cfg = (interface.get_task_configurator("float64")
           .set_hardware_config(potential_buffer_size=16500, group_size=512, allocation_block_size=1 << 24)
           .set_morse_potential([MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6, well_width=10,
                                                      min_r=0.0, max_r=10.0, point_count=16500)] * 2)
           .set_vibwa_algorithm(mass_atom_0=87.62, mass_atom_1=87.62, integration_step=0.1,
                                min_distance_to_asymptote=0.1, min_level=0, max_level=levels))
    t = time.perf_counter()
    h = interface.submit_task(cfg)
    h.wait()
    return (time.perf_counter() - t) * 1e3, h.get_device_milliseconds()


print("first task: wall %.2f ms, device solve %.2f ms" % task())
walls, devs = zip(*[task() for _ in range(20)])
print("next 20 tasks: wall median %.2f ms (min %.2f), device solve median %.2f ms" %
      (sorted(walls)[10], min(walls), sorted(devs)[10]))
