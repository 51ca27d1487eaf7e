"""Drop-in import shim: ``epseon_backend.*`` resolves to the B200-native build.

A user of the reference imports ``epseon_backend.device.gpu._libepseon_gpu``
(docs/examples/example.py:4-7 in the reference); this package keeps that path working and forwards
to ``epseon_backend_b200``.  Nothing here computes anything.
"""
__version__ = "0.1.0"
