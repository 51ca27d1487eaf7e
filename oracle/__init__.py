"""CPU oracle for the Numerov hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this package.  It is the build's own plain-C statement of the
numerical spec; the reference has no implementation of the
path (vibwa.hpp:605-637 allocates buffers only), so parity against the
reference is UNPINNED -- see ``numerov_oracle.c``.
"""
from .oracle import Oracle, build_oracle  # noqa: F401
