"""INTEGRATION.md maps every entry point of include/epseon_cuda.h to the reference interface it replaces."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_every_export_is_documented():
    header = (ROOT / "include" / "epseon_cuda.h").read_text()
    exports = set(re.findall(r"\b(eps_[a-z0-9_]+)\s*\(", header))
    doc = (ROOT / "INTEGRATION.md").read_text()
    missing = sorted(e for e in exports if e not in doc)
    assert not missing, missing
    assert len(exports) >= 50


def test_cabi_symbol_list_matches_the_header():
    import sys

    sys.path.insert(0, str(ROOT))
    from epseon_backend_b200 import cabi

    header = (ROOT / "include" / "epseon_cuda.h").read_text()
    exports = set(re.findall(r"\b(eps_[a-z0-9_]+)\s*\(", header))
    assert exports == set(cabi.SYMBOLS), (sorted(exports - set(cabi.SYMBOLS)), sorted(set(cabi.SYMBOLS) - exports))
