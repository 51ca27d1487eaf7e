"""GPU compute module (``_libepseon_gpu``), CUDA / sm_100a build."""
