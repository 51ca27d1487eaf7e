"""Markdown scaling table from the committed bench lines profiles/r2_bench_n{1,2,4,8}.json."""
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
lines = {n: json.loads((P / f"r2_bench_n{n}.json").read_text()) for n in (1, 2, 4, 8) if (P / f"r2_bench_n{n}.json").exists()}
base = lines[1]


def rec(d, name, mode=None):
    r = d if name == "c5" else d["sub_records"][name]
    return r[mode] if mode else r


print("| record | " + " | ".join(f"{n} GPU{'s' if n > 1 else ''}" for n in lines) + " |")
print("|---|" + "---:|" * len(lines))
for name, mode, label in (("c5", None, "c5 strong (energy range)"), ("c5", "accurate_mode", "c5 accurate (D form)"),
                          ("c4", None, "c4 strong (by curve)"), ("c4", "accurate_mode", "c4 accurate"),
                          ("c2", None, "c2 weak (one curve per GPU): time to all levels"), ("c3", None, "c3 replicas")):
    cells = []
    for n, d in lines.items():
        r, r1 = rec(d, name, mode), rec(base, name, mode)
        sp = r["value"] / r1["value"]
        cells.append(f"{r['ms_per_step']:.2f} ms, e2e {r['e2e']['ms_per_step']:.2f} ms ({sp:.2f}x)")
    print(f"| {label} | " + " | ".join(cells) + " |")
