"""``epseon_backend.device.gpu`` -> the CUDA build of ``_libepseon_gpu``."""
import sys as _sys

from epseon_backend_b200.device.gpu import _libepseon_gpu

_sys.modules[__name__ + "._libepseon_gpu"] = _libepseon_gpu
