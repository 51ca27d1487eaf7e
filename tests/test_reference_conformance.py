"""The REFERENCE's own conformance suites, unmodified, against this build's boundary.

``oracle/ref_conformance.py`` compiles the reference's gtest sources where they lie under
``/root/reference`` (with the googletest the reference vendors) against
``epseon_backend_b200/cpp/include``; the binaries live in ``oracle/_ref/conformance`` (git-ignored,
travels to the GPU box).  CPU: the four typed configuration suites
(``cpp/gpu/test/task_configurator/*.cpp``).  GPU: ``test_libgpu.cpp``, ``test_compute_context.cpp``
and the reference's pytest file ``python/test/test_device/test_gpu/test_libepseon_gpu.py`` run as
it is (``--noconftest --import-mode=importlib``).  Every reference test must pass: no allow-list.
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_conformance as rc  # noqa: E402


def _binaries_ready() -> bool:
    if rc.available():
        import __graft_entry__ as ge

        ge.build()
    return all((rc.OUT / f"{n}.bin").exists() for n in rc.GTESTS)


pytestmark = pytest.mark.skipif(not _binaries_ready(), reason="reference tree absent and no prebuilt oracle/_ref/conformance")

CPU_SUITES = [n for n, (_, dev) in rc.GTESTS.items() if not dev]
GPU_SUITES = [n for n, (_, dev) in rc.GTESTS.items() if dev]
# typed gtests of the reference (float + double): tests per binary, counted from the sources
EXPECTED_COUNTS = {"test_hardware_config": 14, "test_potential_source": 24, "test_algorithm_confgu": 18,
                   "test_task_configurator": 18}


@pytest.mark.parametrize("name", CPU_SUITES)
def test_reference_config_gtests_unmodified(name):
    passed, failed, out = rc.run_gtest(name)
    assert not failed, out
    assert len(passed) == EXPECTED_COUNTS[name], out


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_SUITES)
def test_reference_device_gtests_unmodified(name):
    passed, failed, out = rc.run_gtest(name)
    assert not failed, out
    assert passed, out


@pytest.mark.gpu
def test_reference_pytest_file_unmodified():
    """python/test/test_device/test_gpu/test_libepseon_gpu.py of the reference, byte for byte."""
    f = rc.OUT / "test_libepseon_gpu.py"
    assert f.exists()
    res = subprocess.run([sys.executable, "-m", "pytest", str(f), "--noconftest", "--import-mode=importlib", "-q",
                          "-p", "no:cacheprovider", f"--rootdir={ROOT}"],
                         capture_output=True, text=True, cwd=str(ROOT), env={**__import__("os").environ, "PYTHONPATH": str(ROOT)})
    tail = res.stdout[-3000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout and "failed" not in res.stdout, tail
