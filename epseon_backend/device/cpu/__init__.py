"""``epseon_backend.device.cpu`` -> ``_libepseon_cpu`` (greet only, as in the reference)."""
import sys as _sys

from epseon_backend_b200.device.cpu import _libepseon_cpu

_sys.modules[__name__ + "._libepseon_cpu"] = _libepseon_cpu
