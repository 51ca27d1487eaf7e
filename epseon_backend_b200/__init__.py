"""epseon_backend_b200 -- B200-native (sm_100a) build of epseon_backend's Numerov hot path.

Layout
  csrc/     CUDA kernels + the C ABI (``include/epseon_cuda.h``) -> ``lib/libepseon_cuda.so``
  cpp/      C++ host mirror of the reference interface (``epseon::gpu::cpp``) + pybind11 module
  device/   Python packages holding the built ``_libepseon_gpu`` / ``_libepseon_cpu`` modules
  cabi.py   ctypes binding of the C ABI (tests / bench call the hot path through it)

There is no CPU fallback: without ``lib/libepseon_cuda.so`` (run ``__graft_entry__.build()``)
every entry point raises.
"""
__version__ = "0.1.0"
