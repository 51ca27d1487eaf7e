#!/usr/bin/env python
"""Generator of tests/golden/accuracy_floor.json: eigenvalues of the DISCRETE Numerov problem of the
BASELINE grid sizes, solved in IEEE binary128 with the textbook recurrence (oracle/numerov_quad.c:
a division per step, bisection on the node count to 2^-100) -- independent of the product
recurrences in form, precision and root finder.  One entry per (config, table kind):

    kind 0: table F_k = (1 - q_k)/12   (the X form's discrete problem)
    kind 1: table A_k = 12 q_k         (the D form's; same problem up to the rounding of the table)

Run here (CPU, ~10 minutes for the 10^6-point grid):   python tests/golden/make_accuracy_golden.py
The table's SHA-256 is stored so that the test can tell a changed workload from a changed result.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import Oracle  # noqa: E402
from oracle.oracle import QuadReference  # noqa: E402
from tests import workloads as W  # noqa: E402

CONFIGS = {
    "c1": (lambda: W.c1(), list(range(17))),
    "c2": (lambda: W.c2(), list(range(17))),
    "c5": (lambda: W.c5(nE=1024), [0, 1, 2, 4, 8, 12, 16]),
    "c3": (lambda: W.c3(), [0, 1, 3, 8, 16]),
}


def main():
    q = QuadReference()
    out = {}
    for name, (make, pick) in CONFIGS.items():
        w = make()
        for kind in (0, 1):
            orc = Oracle(omp=True, form=kind)
            q.set_table_kind(kind)
            T, i0, n, vmin = orc.prep(w["V"], w["s"])
            lev, *_ = orc.solve_levels(T, w["s"], w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-13, 12)
            br = np.array([[lev[v] * (1 - 1e-6) - 1e-3, lev[v] * (1 + 1e-6) + 1e-3] for v in pick])
            hi, lo = np.empty(len(pick)), np.empty(len(pick))
            for i, v in enumerate(pick):
                h, l = q.levels(T, w["s"], v, v, br[i:i + 1])
                hi[i], lo[i] = h[0], l[0]
            rel = np.abs((lev[pick] - hi) - lo) / np.abs(hi)
            out[f"{name}/kind{kind}"] = {
                "n_steps": int(n), "table_sha256": hashlib.sha256(T.tobytes()).hexdigest(), "levels": pick,
                "E_hi": [float.hex(float(x)) for x in hi], "E_lo": [float.hex(float(x)) for x in lo],
                "oracle_rel_err_at_generation": [float(x) for x in rel],
            }
            print(name, kind, n, "max rel err of the FP64 oracle: %.2e" % rel.max(), flush=True)
    path = Path(__file__).resolve().parent / "accuracy_floor.json"
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
