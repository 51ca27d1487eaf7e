"""pytest configuration: marker registration + shared fixtures."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def oracle_d():
    """The oracle's statement of the accurate D form (EPS_OPT_FORM = 1): what the drop-in
    VibwaAlgorithm<FP>::run uses."""
    from oracle import Oracle

    return Oracle(form=1)


@pytest.fixture(scope="session")
def gpu_ctx():
    """One C-ABI context on cuda:0; fails loudly (no fallback) when the library or GPU is missing."""
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    ctx = cabi.Context(0)
    yield ctx
    ctx.close()
