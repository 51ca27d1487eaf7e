"""Static checks of the built sm_100a code (cuobjdump on lib/libepseon_cuda.so; no GPU needed).

They guard the properties DESIGN.md section 4 relies on and that a harmless-looking source edit can
silently lose: the table travels by TMA bulk copies behind mbarriers, the constant-bank kernel feeds
the DADD from a uniform register, the hot kernels do not spill, and the statement order of
numerov_step still lets ptxas feed the three-register DFMA from the operand reuse cache (worth 6 % of
the FP64 pipe, section 4.4)."""
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "epseon_backend_b200" / "lib" / "libepseon_cuda.so"
sys.path.insert(0, str(ROOT / "scripts"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass():
    import __graft_entry__ as ge

    ge.build()
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for m in re.finditer(r"Function : (\S+)(.*?)(?=Function :|\Z)", txt, re.S):
        funcs[m.group(1)] = m.group(2)
    assert "arch = sm_100a" in txt or "sm_100a" in txt
    return funcs


def _one(funcs, *needles):
    hits = [k for k in funcs if all(n in k for n in needles)]
    assert hits, needles
    return funcs[hits[0]]


def test_tma_ring_and_mbarriers_in_the_sweep_kernel(sass):
    body = _one(sass, "numerov_sweep_kernelILi4ELi4ELi32ELb0ELb0E")
    assert "UBLKCP" in body          # cp.async.bulk global -> shared (TMA 1-D bulk copy)
    assert "SYNCS" in body           # mbarrier arrive / try_wait
    assert "LDS.128" in body         # two grid steps per warp-broadcast shared load
    assert not re.search(r"\b(STL|LDL)\b", body), "local-memory spill in the sweep kernel"


def test_constant_bank_kernel_uses_a_uniform_operand(sass):
    body = _one(sass, "numerov_cbank_kernelILi4ELi128ELi32ELb0E")
    assert re.search(r"LDCU(\.64|\.128)? UR\d+, c\[0x0\]\[UR\d+", body)   # table from the parameter bank
    assert re.search(r"DADD R\d+, -?R\d+(\.reuse)?, -?UR\d+", body)        # DADD with a uniform-register source
    assert not re.search(r"\b(STL|LDL)\b", body)


@pytest.mark.parametrize("kernel,min_share", [
    ("numerov_sweep_kernelILi4ELi4ELi32ELb0ELb0E", 0.55),   # C2 / C5-class sweeps (measured 107 of 164)
    ("numerov_sweep_kernelILi4ELi4ELi8ELb0ELb0E", 0.40),    # C4 coarse sweep
    ("numerov_sweep_kernelILi2ELi4ELi8ELb0ELb0E", 0.45),    # C4 refinement in 256-energy CTAs
    ("numerov_cbank_kernelILi4ELi128ELi32ELb0E", 0.40),     # C5
])
def test_three_register_dfma_fed_by_the_reuse_cache(kernel, min_share):
    from sass_reuse import analyze

    good, bad = analyze(str(LIB), kernel)
    assert good + bad > 0
    assert good / (good + bad) >= min_share, (kernel, good, bad)


def test_only_sm_100a_code_is_embedded():
    txt = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", txt))
    assert archs == {"sm_100a"}, archs
