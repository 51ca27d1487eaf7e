"""Generate tests/golden/rotation_mpmath.json: 60-digit mpmath known answers for the two pieces added
next to the path (DESIGN.md sections 3.7 and 3.8), from an INDEPENDENT statement of the spec in
numpy + mpmath (not the oracle, not the CUDA code):

  * rotating curves: V_J = V + J(J+1) (h^2 / 12 s) / r^2 formed in numpy float64 with the spec's
    operation order, then the Numerov recurrence replayed in 60-digit arithmetic on those exact
    float64 coefficients -> exact node counts, tails to ~1e-50;
  * level corrections: the matched outward/inward solution and the Rayleigh quotient of the Numerov
    pencil in 60-digit arithmetic -> dE for trial energies a few tenths of a cm^-1 off a level.

Run:  python tests/golden/make_golden_rot.py   (about a minute)
"""
import json
import sys
from pathlib import Path

import mpmath as mp
import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from tests import workloads as W  # noqa: E402
from tests.golden.make_golden import prep, replay  # noqa: E402  (same independent prep / replay)

mp.mp.dps = 60


def centrifugal(V, s, rmin, h, J):
    r = rmin + np.arange(V.size, dtype=np.float64) * h
    cj = (np.float64(J * (J + 1)) * (h * h)) / (12.0 * s)
    return np.minimum(V + cj / (r * r), 1e300) if J else V.copy()


def correction(F, s, E):
    """dE of DESIGN.md section 3.8 in 60-digit arithmetic on the float64 coefficients."""
    n = len(F)
    ep = mp.mpf(float((np.float64(s) * np.float64(E)) / np.float64(12.0)))
    fp = [mp.mpf(float(f)) + ep for f in F]
    twelfth = mp.mpf(float(np.float64(1.0) / np.float64(12.0)))
    m = max(k for k in range(n) if fp[k] > twelfth)
    m = min(max(m, 1), n - 2)
    psi = [mp.mpf(0)] * n
    u_prev, u = mp.mpf(0), fp[0]
    for k in range(m + 1):                       # outward, k = 0..m
        psi[k] = u / fp[k]
        u_prev, u = u, (1 / fp[k] - 10) * u - u_prev
    u_prev, u = mp.mpf(0), fp[n - 1]
    inward = {}
    for k in range(n - 1, m - 1, -1):            # inward, k = n-1..m
        inward[k] = u / fp[k]
        u_prev, u = u, (1 / fp[k] - 10) * u - u_prev
    rho = psi[m] / inward[m]
    for k in range(m + 1, n):
        psi[k] = inward[k] * rho
    uu = [fp[k] * psi[k] for k in range(n)]
    r_m = (uu[m + 1] - uu[m]) - (uu[m] - uu[m - 1]) - (1 - 12 * fp[m]) * psi[m]
    D = mp.mpf(0)
    for k in range(n):
        left = psi[k - 1] if k > 0 else 0
        right = psi[k + 1] if k + 1 < n else 0
        D += psi[k] * (left + 10 * psi[k] + right)
    D /= 12
    return float(-(psi[m] * r_m / D) / mp.mpf(float(s))), m


def main():
    out = dict(comment="mpmath 60-digit known answers for rotating curves and level corrections; see make_golden_rot.py",
               rotation=[], corrections=[])
    N, rmin, rmax = 2000, 0.2, 6.0
    h = W.grid_h(rmin, rmax, N)
    V = W.morse(38267.0, 0.7414, 1.9426, rmin, rmax, N)
    s = W.scale(1.00783, 1.00783, h)
    for J in (1, 7, 25):
        VJ = centrifugal(V, s, rmin, h, J)
        F, i0, n = prep(VJ, s)
        rows = []
        for E in (float(VJ.min()) + 50.0, 5000.0, 15000.0, 30000.0, 38000.0):
            nodes, man, ex = replay(F, s, E)
            rows.append(dict(E=float(E).hex(), nodes=nodes, tail_mant=man.hex(), tail_exp=ex))
            print("J", J, E, nodes, man, ex, flush=True)
        out["rotation"].append(dict(name=f"h2_morse_N2000_J{J}", J=J, rmin=rmin, h=float(h).hex(), s=float(s).hex(),
                                    i0=i0, n_steps=n, V=[float(v).hex() for v in V],
                                    VJ_first=float(VJ[0]).hex(), VJ_mid=float(VJ[N // 2]).hex(), rows=rows))
    F, i0, n = prep(V, s)
    exact = W.morse_levels(W.H2["De"], W.H2["a"], W.H2["m0"], W.H2["m1"])
    rows = []
    for v, off in ((0, 0.3), (3, -0.2), (9, 0.4), (14, -0.05)):
        E = float(exact[v]) + off
        dE, m = correction(F, s, E)
        rows.append(dict(E=float(E).hex(), dE=float(dE).hex(), m=m))
        print("corr", v, off, dE, m, flush=True)
    out["corrections"].append(dict(name="h2_morse_N2000", h=float(h).hex(), s=float(s).hex(),
                                   V=[float(v).hex() for v in V], rows=rows))
    Path(__file__).with_name("rotation_mpmath.json").write_text(json.dumps(out))
    print("wrote", Path(__file__).with_name("rotation_mpmath.json"))


if __name__ == "__main__":
    main()
