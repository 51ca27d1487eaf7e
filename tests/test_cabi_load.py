"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/epseon_cuda.h declares; with no GPU it fails loudly instead of falling back."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    return cabi.load()


def test_header_symbols_exported(lib):
    from epseon_backend_b200 import cabi

    header = (ROOT / "include" / "epseon_cuda.h").read_text()
    declared = sorted(set(re.findall(r"\b(eps_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/epseon_cuda.h but not exported"
    assert sorted(cabi.SYMBOLS) == declared
    assert lib.eps_abi_version() == int(re.search(r"EPS_ABI_VERSION (\d+)", header).group(1))


def test_struct_sizes_match_header(lib):
    """ctypes mirrors must match the C layout (compile a probe with gcc)."""
    import ctypes as C
    import subprocess
    import tempfile

    from epseon_backend_b200 import cabi

    src = ('#include "epseon_cuda.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n",'
           "sizeof(eps_device_props),sizeof(eps_curve_info),sizeof(eps_solve_params),sizeof(eps_stats));}")
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "p.c").write_text(src)
        subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), "-o", f"{d}/p", f"{d}/p.c"], check=True)
        sizes = list(map(int, subprocess.check_output([f"{d}/p"]).split()))
    assert sizes == [C.sizeof(cabi.DeviceProps), C.sizeof(cabi.CurveInfo), C.sizeof(cabi.SolveParams),
                     C.sizeof(cabi.Stats)]


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device the product path must raise (never route through the oracle)."""
    from epseon_backend_b200 import cabi

    try:
        n = cabi.device_count()
    except cabi.EpsError as e:
        assert e.code == 2
        n = 0
    if n == 0:
        with pytest.raises(cabi.EpsError):
            cabi.Context(0)


def test_product_does_not_reference_oracle():
    """The product may mention the oracle in comments, but never import, include, link or call it."""
    bad = [re.compile(p, re.M) for p in (r"^\s*(import|from)\s+oracle", r"liboracle",
                                         r'#include\s*[<"][^>"]*oracle', r"\borc_[a-z_]+\s*\(")]
    for p in (ROOT / "epseon_backend_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"}:
            txt = p.read_text()
            for rx in bad:
                assert not rx.search(txt), f"{p} references the oracle: {rx.pattern}"
