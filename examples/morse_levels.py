"""End-to-end use of the drop-in API: the call sequence of the reference's own example
(docs/examples/example.py: context -> device interface -> configurator -> submit_task -> wait),
followed by the additive result accessors of this build.

    python examples/morse_levels.py            # needs a B200 (sm_100-class GPU); there is no CPU fallback
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))  # run from a source checkout

from epseon_backend.device.gpu._libepseon_gpu import EpseonComputeContext, MorsePotentialConfig  # noqa: E402


def main() -> None:
    ctx = EpseonComputeContext.create()
    device_info = next(iter(ctx.get_physical_device_info()))
    interface = ctx.get_device_interface(device_info.device_properties.device_id)
    cfg = (
        interface.get_task_configurator("float64")
        .set_hardware_config(potential_buffer_size=16500, group_size=512, allocation_block_size=16 * 1024 * 1024)
        .set_morse_potential([
            # units: cm^-1, Angstrom, amu
            MorsePotentialConfig(dissociation_energy=500, equilibrium_bond_distance=2.6, well_width=1.3,
                                 min_r=0.0, max_r=10.0, point_count=16500),
            MorsePotentialConfig(dissociation_energy=5500.0, equilibrium_bond_distance=0.6, well_width=10,
                                 min_r=0.0, max_r=10.0, point_count=16500),
        ])
        .set_vibwa_algorithm(mass_atom_0=87.62, mass_atom_1=87.62, integration_step=0.1,
                             min_distance_to_asymptote=0.1, min_level=0, max_level=11)
        .set_rotational_states([0, 1])       # additive: J = 0 and J = 1 of every curve
        .set_wavefunction_output(True)       # additive: normalised wavefunctions of the located levels
    )
    handle = interface.submit_task(cfg)
    handle.wait()
    if handle.has_failed():
        raise SystemExit(handle.get_status_message())

    levels = np.array(handle.get_levels())            # [curve * 2 + J][level], NaN where not found
    counts = handle.get_level_counts()                # levels below the search ceiling per row
    psi = handle.get_wavefunctions()                  # [row][level][grid point]
    print(f"device: {device_info.device_properties.device_name}; solve took "
          f"{handle.get_device_milliseconds():.2f} ms on the device")
    for row, (lev, n) in enumerate(zip(levels, counts)):
        curve, J = divmod(row, 2)
        found = lev[np.isfinite(lev)]
        print(f"curve {curve} J={J}: {n} levels below the ceiling; E_0..E_{len(found) - 1} [cm^-1] =",
              np.array2string(found, precision=4, max_line_width=120))
    h = 10.0 / 16499
    print("norms h*sum(psi^2) of the first curve's levels:", np.round(h * np.sum(psi[0] ** 2, axis=1), 12)[:4], "...")
    b_rot = (levels[1][0] - levels[0][0]) / 2.0  # E(v=0, J=1) - E(v=0, J=0) = 2 B_0
    print(f"rotational constant of curve 0 from E(0,1) - E(0,0): B_0 = {b_rot:.5f} cm^-1 "
          f"(hbar^2 / (2 mu r_e^2) = {16.857629206 / 43.81 / 2.6**2:.5f})")


if __name__ == "__main__":
    main()
