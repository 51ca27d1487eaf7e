"""ctypes wrapper around oracle/numerov_oracle.c (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent


def _cpu_tag() -> str:
    """-march=native objects must not travel between hosts: key the build dir by CPU flags."""
    import hashlib

    flags = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = line
                break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_BUILD = _HERE / "_build" / _cpu_tag()

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build_oracle(force: bool = False) -> None:
    """Compile the oracle with gcc (a few seconds)."""
    srcs = [_HERE / "numerov_oracle.c", _HERE / "numerov_quad.c", _HERE / "Makefile"]
    libs = [_BUILD / "liboracle.so", _BUILD / "liboracle_omp.so", _BUILD / "libquad.so"]
    newest = max(p.stat().st_mtime for p in srcs)
    if not force and all(p.exists() and p.stat().st_mtime >= newest for p in libs):
        return
    subprocess.run(["make", "-s", "-C", str(_HERE), f"OUT={_BUILD}"] + (["-B"] if force else []),
                   check=True)


def _opt(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Scalar-order CPU statement of prep / sweep / level search.

    ``omp=True`` loads the OpenMP build (same arithmetic, energies spread over
    host threads); used for the timed CPU baseline.
    """

    def __init__(self, omp: bool = False, threads: int | None = None, form: int = 0):
        """form 0: the 4-operation X form (default); form 1: the accurate 5-operation D form.  The
        table returned by prep() and taken by the sweeps is the one of that form."""
        build_oracle()
        self.form = int(form)
        self.lib = C.CDLL(str(_BUILD / ("liboracle_omp.so" if omp else "liboracle.so")))
        L = self.lib
        if omp and threads:
            L.orc_set_threads(C.c_int(int(threads)))
        L.orc_num_threads.restype = C.c_int
        L.orc_scale.restype = C.c_double
        L.orc_scale.argtypes = [C.c_double] * 3
        L.orc_morse_tabulate.argtypes = [C.c_double] * 5 + [C.c_uint32, _f64p]
        L.orc_lj_tabulate.argtypes = [C.c_double] * 4 + [C.c_uint32, _f64p]
        L.orc_prep_form.restype = C.c_int
        L.orc_prep_form.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_int, _f64p, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.orc_sweep_form.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_int, _f64p, C.c_uint64,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sweep_uniform_form.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_int, C.c_double, C.c_double,
                                             C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_solve_levels.restype = C.c_int
        L.orc_solve_levels.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                       C.c_uint32, _f64p, _f64p, C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint64)]
        L.orc_solve_levels_grid_form.restype = C.c_int
        L.orc_solve_levels_grid_form.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_int, C.c_double, C.c_double, C.c_uint64,
                                            C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                            C.c_uint32, _f64p, _f64p, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.orc_cooley_level.restype = C.c_int
        L.orc_cooley_level.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                       C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_spline_resample.restype = C.c_int
        L.orc_spline_resample.argtypes = [_f64p, _f64p, C.c_uint32, C.c_double, C.c_double, C.c_uint32, _f64p]
        L.orc_centrifugal.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_uint32, _f64p]
        L.orc_level_correction.restype = C.c_double
        L.orc_level_correction.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_double, _f64p, C.c_int64]
        L.orc_wavefunction.restype = C.c_int64
        L.orc_wavefunction.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_double, C.c_double, _f64p]

    @property
    def threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def scale(self, m0: float, m1: float, h: float) -> float:
        return float(self.lib.orc_scale(m0, m1, h))

    def morse(self, De, re, a, rmin, rmax, N) -> np.ndarray:
        V = np.empty(N, dtype=np.float64)
        self.lib.orc_morse_tabulate(De, re, a, rmin, rmax, N, V)
        return V

    def lj(self, De, re, rmin, rmax, N) -> np.ndarray:
        V = np.empty(N, dtype=np.float64)
        self.lib.orc_lj_tabulate(De, re, rmin, rmax, N, V)
        return V

    def centrifugal(self, V, s, rmin, h, J) -> np.ndarray:
        """V_J = V + J(J+1) (h^2 / 12 s) / r^2 on r_i = rmin + i h (J = 0: copy)."""
        V = np.ascontiguousarray(V, dtype=np.float64)
        out = np.empty_like(V)
        self.lib.orc_centrifugal(V, V.size, float(s), float(rmin), float(h), int(J), out)
        return out

    def prep(self, V: np.ndarray, s: float):
        """-> (F[n_steps], i0, n_steps, vmin)"""
        V = np.ascontiguousarray(V, dtype=np.float64)
        AB = np.empty(V.size, dtype=np.float64)
        i0, n, vmin = C.c_uint32(), C.c_uint32(), C.c_double()
        rc = self.lib.orc_prep_form(V, V.size, s, self.form, AB, C.byref(i0), C.byref(n), C.byref(vmin))
        if rc != 0:
            raise ValueError("orc_prep: unusable potential table")
        return AB[: n.value].copy(), i0.value, n.value, vmin.value

    def sweep(self, AB: np.ndarray, s: float, E: np.ndarray, tails: bool = True):
        E = np.ascontiguousarray(E, dtype=np.float64)
        n_steps = AB.size
        nodes = np.empty(E.size, dtype=np.uint32)
        mant = np.empty(E.size, dtype=np.float64) if tails else None
        expo = np.empty(E.size, dtype=np.int32) if tails else None
        self.lib.orc_sweep_form(AB, n_steps, s, self.form, E, E.size, _opt(nodes, np.uint32),
                           _opt(mant, np.float64), _opt(expo, np.int32))
        return nodes, mant, expo

    def sweep_uniform(self, AB, s, E0, dE, j0, nE, tails: bool = True):
        n_steps = AB.size
        nodes = np.empty(nE, dtype=np.uint32)
        mant = np.empty(nE, dtype=np.float64) if tails else None
        expo = np.empty(nE, dtype=np.int32) if tails else None
        self.lib.orc_sweep_uniform_form(AB, n_steps, s, self.form, E0, dE, j0, nE, _opt(nodes, np.uint32),
                                   _opt(mant, np.float64), _opt(expo, np.int32))
        return nodes, mant, expo

    def solve_levels(self, AB, s, E_lo, E_hi, n_coarse, vmin, vmax, M, rel_tol=1e-12,
                     max_rounds=8):
        """-> (levels[nlev], widths[nlev], n_below_hi, rounds, steps)"""
        n_steps = AB.size
        nlev = vmax - vmin + 1
        levels = np.empty(nlev, dtype=np.float64)
        widths = np.empty(nlev, dtype=np.float64)
        nb, st = C.c_uint32(), C.c_uint64()
        dE = (E_hi - E_lo) / float(n_coarse - 1)  # orc_solve_levels: uniform grid, j0 = 0
        rounds = self.lib.orc_solve_levels_grid_form(AB, n_steps, s, self.form, E_lo, dE, 0, n_coarse, vmin, vmax, M,
                                                     rel_tol, max_rounds, levels, widths, C.byref(nb), None,
                                                     C.byref(st))
        return levels, widths, nb.value, rounds, st.value

    def solve_levels_grid(self, AB, s, E0, dE, j0, n_coarse, vmin, vmax, M, rel_tol=1e-12, max_rounds=8):
        """Coarse grid E_j = E0 + (j0 + j) dE -> (levels, widths, n_last, n_first, rounds, steps)"""
        nlev = vmax - vmin + 1
        levels = np.empty(nlev, dtype=np.float64)
        widths = np.empty(nlev, dtype=np.float64)
        nb, nf, st = C.c_uint32(), C.c_uint32(), C.c_uint64()
        rounds = self.lib.orc_solve_levels_grid_form(AB, AB.size, s, self.form, E0, dE, j0, n_coarse, vmin, vmax, M, rel_tol,
                                                max_rounds, levels, widths, C.byref(nb), C.byref(nf), C.byref(st))
        return levels, widths, nb.value, nf.value, rounds, st.value

    def cooley_level(self, A, s, v, lo, hi, rel_tol=1e-12, max_iter=30, open_tail=False, seg_len=0):
        """Outward/inward matching search of level v inside the bracket [lo, hi] on the D-form table A
        -> (E, |last correction|, iterations)."""
        assert self.form == 1, "the Cooley search runs on the D-form table (Oracle(form=1))"
        E, wd = C.c_double(), C.c_double()
        it = self.lib.orc_cooley_level(np.ascontiguousarray(A), A.size, float(s), int(v), float(lo), float(hi),
                                       float(rel_tol), int(max_iter), int(bool(open_tail)), int(seg_len), C.byref(E), C.byref(wd))
        return E.value, wd.value, int(it)

    def solve_levels_cooley(self, A, s, E_lo, E_hi, n_coarse, vmin, vmax, rel_tol=1e-12, max_iter=30, open_tail=False,
                            seg_len=0):
        """Coarse sweep + bracketing as solve_levels, then the Cooley search per level
        -> (levels[nlev], widths[nlev], n_below_hi, iterations[nlev])."""
        dE = (E_hi - E_lo) / float(n_coarse - 1)
        nodes, _, _ = self.sweep_uniform(A, s, E_lo, dE, 0, n_coarse, tails=False)
        nlev = vmax - vmin + 1
        lev, wid, its = np.full(nlev, np.nan), np.full(nlev, np.nan), np.zeros(nlev, dtype=np.uint32)
        for l in range(nlev):
            v = vmin + l
            if nodes[-1] <= v or nodes[0] > v:
                continue
            j = int(np.argmax(nodes > v))
            lo, hi = E_lo + float(j - 1) * dE, E_lo + float(j) * dE
            lev[l], wid[l], its[l] = self.cooley_level(A, s, v, lo, hi, rel_tol, max_iter, open_tail, seg_len)
        return lev, wid, int(nodes[-1]), its

    def wavefunction(self, AB, s, E, h):
        """-> (psi[n_steps] on the integration window, match index m or -1)"""
        psi = np.empty(AB.size, dtype=np.float64)
        m = self.lib.orc_wavefunction(AB, AB.size, s, float(E), float(h), psi)
        return psi, int(m)

    def level_correction(self, AB, s, E, h):
        """First-order energy correction dE of a trial level E (Rayleigh quotient of the Numerov pencil)."""
        psi, m = self.wavefunction(AB, s, E, h)
        return float(self.lib.orc_level_correction(AB, AB.size, float(s), float(E), psi, m))

    def spline_resample(self, r, V, rmin, rmax, N):
        """Natural cubic spline through (r, V) resampled on N uniform points of [rmin, rmax]."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        out = np.empty(N, dtype=np.float64)
        if self.lib.orc_spline_resample(r, V, r.size, float(rmin), float(rmax), N, out) != 0:
            raise ValueError("orc_spline_resample: bad knots")
        return out


class QuadReference:
    """Independent binary128 solution of the discrete Numerov eigenproblem (oracle/numerov_quad.c):
    textbook recurrence with a division per step, bisection on the node count to 2^-100."""

    def __init__(self):
        build_oracle()
        self.lib = C.CDLL(str(_BUILD / "libquad.so"))
        L = self.lib
        L.quad_node_count.restype = C.c_uint32
        L.quad_node_count.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_double]
        L.quad_levels.restype = None
        L.quad_levels.argtypes = [_f64p, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, _f64p, _f64p, _f64p]

    def set_table_kind(self, kind: int) -> None:
        """0: table F_k = (1 - q_k)/12 (X form, default); 1: table A_k = 12 q_k (D form)."""
        self.lib.quad_set_table_kind(C.c_int(kind))

    def node_count(self, F, s, E) -> int:
        return int(self.lib.quad_node_count(np.ascontiguousarray(F), F.size, float(s), float(E)))

    def levels(self, F, s, v_min, v_max, brackets):
        """brackets[nlev, 2] = (E_lo, E_hi) per level -> (E as double, remainder): E = hi + lo to ~1e-30."""
        F = np.ascontiguousarray(F, dtype=np.float64)
        br = np.ascontiguousarray(brackets, dtype=np.float64).reshape(-1)
        n = v_max - v_min + 1
        hi, lo = np.empty(n), np.empty(n)
        self.lib.quad_levels(F, F.size, float(s), v_min, v_max, br, hi, lo)
        return hi, lo
