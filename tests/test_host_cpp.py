"""C++ host mirror of the reference interface: compile and run the conformance programs.

test_config.cpp restates the reference's typed gtests for the configuration classes
(cpp/gpu/test/task_configurator/*.cpp) -- CPU only.  test_libgpu.cpp restates
cpp/gpu/test/test_libgpu.cpp + test_compute_context.cpp and needs a GPU.
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CPP = ROOT / "epseon_backend_b200" / "cpp"
LIB = ROOT / "epseon_backend_b200" / "lib"


def _build(name: str, link_cuda: bool) -> Path:
    import __graft_entry__ as ge

    ge.build()
    out = CPP / "test" / f"{name}.bin"
    srcs = [CPP / "test" / f"{name}.cpp"] + sorted((CPP / "source").glob("*.cpp"))
    cmd = ["/usr/bin/g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-I", str(CPP / "include"),
           "-I", str(ROOT / "include"), *map(str, srcs), "-o", str(out), f"-L{LIB}", "-lepseon_cuda",
           f"-Wl,-rpath,{LIB}"]
    subprocess.run(cmd, check=True)
    return out


def test_config_classes_cpu():
    exe = _build("test_config", link_cuda=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert res.stdout.startswith("OK")


@pytest.mark.gpu
def test_libgpu_device_integration():
    exe = _build("test_libgpu", link_cuda=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    sys.stdout.write(res.stdout)
    assert res.returncode == 0, res.stderr
    assert res.stdout.strip().endswith("OK")
