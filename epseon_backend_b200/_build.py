"""In-tree build of the native pieces (nvcc / g++ directly; no cmake, no network)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB_DIR = PKG / "lib"
CUDA_LIB = LIB_DIR / "libepseon_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]
GXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -> lib/libepseon_cuda.so (sm_100a only)."""
    srcs = [PKG / "csrc" / "epseon_cuda.cu"]
    deps = srcs + sorted((PKG / "csrc").glob("*.cuh")) + [ROOT / "include" / "epseon_cuda.h"]
    if force or _stale(CUDA_LIB, deps):
        LIB_DIR.mkdir(exist_ok=True)
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", GXX, "-I", str(ROOT / "include")]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        cmd += ["-o", str(CUDA_LIB), *map(str, srcs)]
        subprocess.run(cmd, check=True)
    return CUDA_LIB


def ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def build_pybind(force: bool = False) -> list[Path]:
    """g++ -> device/gpu/_libepseon_gpu*.so and device/cpu/_libepseon_cpu*.so (pybind11)."""
    cpp = PKG / "cpp"
    if not (cpp / "python" / "api.cpp").exists():
        return []
    import pybind11

    inc = ["-I", str(cpp / "include"), "-I", str(ROOT / "include"), "-I", pybind11.get_include(),
           "-I", sysconfig.get_paths()["include"]]
    common = [GXX, "-O2", "-std=c++20", "-fPIC", "-shared", "-fvisibility=hidden",
              "-ffp-contract=off", "-pthread"]
    out = []
    gpu_so = PKG / "device" / "gpu" / f"_libepseon_gpu{ext_suffix()}"
    gpu_src = sorted((cpp / "source").rglob("*.cpp")) + [cpp / "python" / "api.cpp"]
    gpu_dep = gpu_src + sorted((cpp / "include").rglob("*.hpp")) + [ROOT / "include" / "epseon_cuda.h"]
    if force or _stale(gpu_so, gpu_dep):
        subprocess.run([*common, *inc, *map(str, gpu_src), "-o", str(gpu_so),
                        f"-L{LIB_DIR}", "-lepseon_cuda", "-Wl,-rpath,$ORIGIN/../../lib"], check=True)
    out.append(gpu_so)
    cpu_src = cpp / "python" / "libcpu.cpp"
    if cpu_src.exists():
        cpu_so = PKG / "device" / "cpu" / f"_libepseon_cpu{ext_suffix()}"
        if force or _stale(cpu_so, [cpu_src]):
            subprocess.run([*common, *inc, str(cpu_src), "-o", str(cpu_so)], check=True)
        out.append(cpu_so)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_pybind(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", CUDA_LIB)
