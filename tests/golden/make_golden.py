"""Generate tests/golden/numerov_mpmath.json: a 60-digit mpmath replay of the Numerov recurrence.

The reference ships no golden vectors for this path (it has no implementation of it), so these
known answers are produced by an INDEPENDENT statement of the spec in
numpy + mpmath -- not by the oracle and not by the CUDA code:
  * preparation (window, F_k) in numpy float64, operation order as in the spec;
  * the recurrence itself in 60-digit arithmetic on those exact float64 coefficients (F_k, e/12), so the
    node counts are the mathematically exact ones for these inputs and the tails are correct to
    ~1e-50 (the float64 implementations must match nodes exactly and tails to rounding growth).
Also stores the analytic Morse spectrum of the C1 curve.
Run:  python tests/golden/make_golden.py   (about a minute)
"""
import json
import sys
from pathlib import Path

import mpmath as mp
import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from tests import workloads as W  # noqa: E402

mp.mp.dps = 60
T_MAX = 0.5


def prep(V, s):
    q = s * V
    m = int(np.argmin(q))
    thr = q[m] + T_MAX
    left = np.nonzero(q[:m] > thr)[0]
    right = np.nonzero(q[m + 1:] > thr)[0]
    ilo = int(left[-1]) + 1 if left.size else 0
    ihi = m + int(right[0]) if right.size else len(V) - 1
    i0 = max(ilo, 1)
    iend = min(ihi + 1, len(V) - 1)
    n = iend - i0
    qq = q[i0:i0 + n]
    return (1.0 - qq) / 12.0, i0, n


def replay(F, s, E):
    """X form of the recurrence in 60-digit arithmetic."""
    ep = (np.float64(s) * np.float64(E)) / np.float64(12.0)
    ep_mp = mp.mpf(float(ep))
    X, S = mp.mpf(1), mp.mpf(0)
    nodes = 0
    for k in range(len(F)):
        fp = mp.mpf(float(F[k])) + ep_mp
        Q = 10 * X + S
        Xn = X - fp * Q
        S = fp * X
        if (Xn < 0) != (X < 0):
            nodes += 1
        X = Xn
    man, ex = mp.frexp(X)  # X = man * 2^ex, 0.5 <= |man| < 1
    return nodes, float(man * 2), int(ex - 1)  # mantissa in [1,2)


def main():
    cases = []
    specs = [
        dict(name="h2_morse_N2000", V=W.morse(38267.0, 0.7414, 1.9426, 0.2, 6.0, 2000),
             s=W.scale(1.00783, 1.00783, W.grid_h(0.2, 6.0, 2000)),
             E=[10.0, 2166.0, 2200.0, 6309.0, 6400.0, 13839.5, 20000.0, 30000.0, 36000.0, 38000.0, 38266.0]),
        dict(name="sr2_fixture_N1500", V=W.morse(5500.0, 0.6, 10.0, 0.0, 10.0, 1500),
             s=W.scale(87.62, 87.62, W.grid_h(0.0, 10.0, 1500)),
             E=[1.0, 300.0, 1000.0, 2500.0, 4000.0, 5400.0]),
        dict(name="lj_N1300", V=W.lj(800.0, 3.0, 2.2, 12.0, 1300),
             s=W.scale(40.0, 40.0, W.grid_h(2.2, 12.0, 1300)),
             E=[5.0, 100.0, 400.0, 700.0, 790.0]),
    ]
    for sp in specs:
        F, i0, n = prep(sp["V"], sp["s"])
        rows = []
        for E in sp["E"]:
            nodes, man, ex = replay(F, sp["s"], E)
            rows.append(dict(E=float(E).hex(), nodes=nodes, tail_mant=man.hex(), tail_exp=ex))
            print(sp["name"], E, nodes, man, ex, flush=True)
        cases.append(dict(name=sp["name"], s=float(sp["s"]).hex(), i0=i0, n_steps=n,
                          V=[float(v).hex() for v in sp["V"]], rows=rows))

    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    out = dict(comment="mpmath 60-digit replay of the Numerov recurrence; see make_golden.py",
               cases=cases, morse_c1_levels=[float(x) for x in exact])
    Path(__file__).with_name("numerov_mpmath.json").write_text(json.dumps(out))
    print("wrote", Path(__file__).with_name("numerov_mpmath.json"))


if __name__ == "__main__":  # prep / replay are also imported by make_golden_rot.py
    main()
