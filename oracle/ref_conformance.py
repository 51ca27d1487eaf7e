"""Build the REFERENCE's own conformance suites, unmodified, against this build's boundary.

Test infrastructure (like everything under ``oracle/``): nothing in the product imports this.

The reference holds no arithmetic for the hot path, but it does hold tests of
the drop-in boundary:

* ``cpp/gpu/test/task_configurator/*.cpp``  -- typed gtests of the configuration classes (CPU);
* ``cpp/gpu/test/test_libgpu.cpp``, ``test_compute_context.cpp`` -- device enumeration, task
  submission, ``startWorker / wait / isDone`` (need a device);
* ``python/test/test_device/test_gpu/test_libepseon_gpu.py`` -- the pybind11 module (needs a device).

The sources are compiled WHERE THEY LIE under ``/root/reference`` (never copied into the
repository's history) against ``epseon_backend_b200/cpp/include`` with the googletest the reference
vendors (``external/googletest``); the binaries -- and a verbatim copy of the pytest file, which
cannot be "compiled" -- go to ``oracle/_ref/conformance/``, a git-ignored directory that travels to
the GPU box with the snapshot exactly like the built ``.so`` files do (``/root/reference`` does not
exist there).  ``__graft_entry__.build()`` calls ``build()`` when the reference tree is present.
"""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "oracle" / "_ref" / "conformance"
CPP = ROOT / "epseon_backend_b200" / "cpp"
LIB = ROOT / "epseon_backend_b200" / "lib"
GXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"

# reference test sources (relative to /root/reference) -> binary name; needs_device
GTESTS = {
    "test_hardware_config": ("cpp/gpu/test/task_configurator/test_hardware_config.cpp", False),
    "test_potential_source": ("cpp/gpu/test/task_configurator/test_potential_source.cpp", False),
    "test_algorithm_confgu": ("cpp/gpu/test/task_configurator/test_algorithm_confgu.cpp", False),
    "test_task_configurator": ("cpp/gpu/test/task_configurator/test_task_configurator.cpp", False),
    "test_libgpu": ("cpp/gpu/test/test_libgpu.cpp", True),
    "test_compute_context": ("cpp/gpu/test/test_compute_context.cpp", True),
}
PYTEST_FILE = "python/test/test_device/test_gpu/test_libepseon_gpu.py"

def available() -> bool:
    return (REF / "external" / "googletest" / "googletest" / "src" / "gtest-all.cc").exists()


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).exists() and Path(s).stat().st_mtime > t for s in sources)


def _gtest_objects() -> list[Path]:
    gt = REF / "external" / "googletest" / "googletest"
    objs = []
    for name in ("gtest-all", "gtest_main"):
        obj = OUT / f"{name}.o"
        if not obj.exists():
            subprocess.run([GXX, "-O1", "-std=c++20", "-pthread", "-I", str(gt / "include"), "-I", str(gt),
                            "-c", str(gt / "src" / f"{name}.cc"), "-o", str(obj)], check=True)
        objs.append(obj)
    return objs


def build(force: bool = False) -> list[Path]:
    """Compile every reference gtest against this build's headers; copy the pytest file.  Needs
    libepseon_cuda.so (``epseon_backend_b200._build.build_cuda``) to link."""
    if not available():
        return sorted(OUT.glob("test_*.bin")) if OUT.exists() else []
    OUT.mkdir(parents=True, exist_ok=True)
    objs = _gtest_objects()
    gt = REF / "external" / "googletest" / "googletest"
    deps = sorted((CPP / "include").rglob("*.hpp")) + sorted((CPP / "source").rglob("*.cpp")) + [ROOT / "include" / "epseon_cuda.h"]
    host_srcs = sorted((CPP / "source").rglob("*.cpp"))
    out = []
    for name, (rel, _needs_device) in GTESTS.items():
        exe = OUT / f"{name}.bin"
        if force or _stale(exe, deps + [REF / rel]):
            subprocess.run([GXX, "-O1", "-std=c++20", "-pthread", "-ffp-contract=off",
                            "-I", str(CPP / "include"), "-I", str(ROOT / "include"), "-I", str(gt / "include"),
                            str(REF / rel), *map(str, host_srcs), *map(str, objs), "-o", str(exe),
                            f"-L{LIB}", "-lepseon_cuda", "-Wl,-rpath,$ORIGIN/../../../epseon_backend_b200/lib"],
                           check=True)
        out.append(exe)
    dst = OUT / "test_libepseon_gpu.py"
    if force or _stale(dst, [REF / PYTEST_FILE]):
        shutil.copyfile(REF / PYTEST_FILE, dst)  # verbatim; git-ignored build artefact, not repository source
    return out


def run_gtest(name: str) -> tuple[set[str], set[str], str]:
    """-> (passed, failed, raw output) of one prebuilt reference gtest binary."""
    exe = OUT / f"{name}.bin"
    res = subprocess.run([str(exe), "--gtest_color=no"], capture_output=True, text=True, cwd=str(OUT))
    passed, failed = set(), set()
    for line in res.stdout.splitlines():
        line = line.strip()
        if line.startswith("[       OK ]"):
            passed.add(line.split("]", 1)[1].split("(")[0].strip())
        elif line.startswith("[  FAILED  ]") and "listed below" not in line and "." in line:
            failed.add(line.split("]", 1)[1].split(",")[0].split("(")[0].strip())
    return passed, failed, res.stdout + res.stderr


if __name__ == "__main__":
    for p in build(force=True):
        print("built", p)
