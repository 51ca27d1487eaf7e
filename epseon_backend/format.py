"""Byte-size pretty printer (same behaviour as the reference's epseon_backend/format.py:21-38)."""
from __future__ import annotations

_UNITS = ("", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei", "Zi", "Yi")


def convert_size_in_bytes_to_adaptive_unit(value: int) -> str:
    """Return ``value`` bytes scaled to the largest binary prefix, three decimals."""
    if value < 0:
        raise ValueError(f"Values below 0 are not allowed, got {value}")
    idx = 0
    while idx < len(_UNITS) - 1 and value >= 1024 ** (idx + 1):
        idx += 1
    whole, rem = divmod(value * 1000, 1024**idx)
    if 2 * rem >= 1024**idx:
        whole += 1
    return f"{whole // 1000}.{whole % 1000:03d}{_UNITS[idx]}B"
