"""ctypes binding of the C ABI declared in ``include/epseon_cuda.h``.

This is the call path tests and ``bench.py`` use to reach the CUDA hot path
"through the C ABI".  It loads ``lib/libepseon_cuda.so`` and nothing else: no
oracle, no CPU fallback -- a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libepseon_cuda.so"

EPS_OK = 0
ERR_NAMES = {1: "EPS_ERR_INVALID", 2: "EPS_ERR_CUDA", 3: "EPS_ERR_RANGE", 4: "EPS_ERR_STATE",
             5: "EPS_ERR_NOMEM", 6: "EPS_ERR_CANCELLED"}
EPS_ERR_CANCELLED = 6

# every symbol include/epseon_cuda.h declares
SYMBOLS = [
    "eps_abi_version", "eps_device_count", "eps_device_get_props", "eps_ctx_create",
    "eps_ctx_destroy", "eps_last_error", "eps_sync", "eps_set_potentials", "eps_set_potentials_rot", "eps_get_curve_info",
    "eps_sweep", "eps_sweep_uniform", "eps_sweep_grid", "eps_solve_levels", "eps_solve_levels_grid", "eps_wavefunctions", "eps_level_corrections", "eps_spline_coefficients", "eps_spline_resample", "eps_set_option", "eps_get_counter", "eps_timer_start", "eps_timer_stop",
    "eps_stats_get", "eps_stats_reset", "eps_l2_flush", "eps_fp64_probe", "eps_host_alloc", "eps_host_free",
    "eps_request_stop", "eps_reset_stop", "eps_ctx_device_bytes", "eps_ctx_trim",
    "eps_group_create", "eps_group_destroy", "eps_group_size", "eps_group_ctx", "eps_group_last_error", "eps_group_last_ms",
    "eps_group_set_option", "eps_group_set_potentials", "eps_group_sweep_uniform", "eps_group_solve_levels",
    "eps_mailbox_create", "eps_mailbox_open", "eps_mailbox_destroy", "eps_mailbox_post_levels", "eps_mailbox_post",
    "eps_mailbox_collect", "eps_mailbox_slot_bytes", "eps_cooley_segment_length",
]


class EpsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class DeviceProps(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("ordinal", C.c_int32), ("cc_major", C.c_int32),
                ("cc_minor", C.c_int32), ("sm_count", C.c_int32), ("clock_khz", C.c_int32),
                ("driver_version", C.c_int32), ("runtime_version", C.c_int32),
                ("pci_domain", C.c_int32), ("pci_bus", C.c_int32), ("pci_device", C.c_int32),
                ("integrated", C.c_int32), ("max_threads_per_block", C.c_int32),
                ("max_grid", C.c_int32 * 3), ("max_block", C.c_int32 * 3), ("l2_bytes", C.c_int32),
                ("total_global_mem", C.c_uint64), ("shared_mem_per_block_optin", C.c_uint64),
                ("shared_mem_per_sm", C.c_uint64), ("uuid", C.c_uint8 * 16)]


class CurveInfo(C.Structure):
    _fields_ = [("i0", C.c_uint32), ("n_steps", C.c_uint32), ("scale", C.c_double),
                ("v_min", C.c_double), ("v_last", C.c_double)]


class SolveParams(C.Structure):
    _fields_ = [("v_min", C.c_uint32), ("v_max", C.c_uint32), ("n_coarse", C.c_uint32),
                ("refine_points", C.c_uint32), ("max_rounds", C.c_uint32), ("flags", C.c_uint32),
                ("rel_tol", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("sweep_launches", C.c_uint64), ("other_launches", C.c_uint64),
                ("grid_steps", C.c_uint64), ("sweep_ms", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint64)]


_lib = None


def load() -> C.CDLL:
    """Load libepseon_cuda.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() -- there is no "
                               "CPU fallback for the hot path")
        lib = C.CDLL(str(LIB_PATH))
        lib.eps_last_error.restype = C.c_char_p
        lib.eps_last_error.argtypes = [C.c_void_p]
        lib.eps_group_last_error.restype = C.c_char_p
        lib.eps_group_last_error.argtypes = [C.c_void_p]
        lib.eps_group_ctx.restype = C.c_void_p
        lib.eps_group_ctx.argtypes = [C.c_void_p, C.c_uint32]
        lib.eps_group_size.restype = C.c_uint32
        lib.eps_group_size.argtypes = [C.c_void_p]
        lib.eps_mailbox_slot_bytes.restype = C.c_size_t
        lib.eps_mailbox_slot_bytes.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().eps_device_count(C.byref(n))
    if rc != EPS_OK:
        raise EpsError(rc, (load().eps_last_error(None) or b"").decode())
    return n.value


def device_props(dev: int) -> DeviceProps:
    p = DeviceProps()
    rc = load().eps_device_get_props(dev, C.byref(p))
    if rc != EPS_OK:
        raise EpsError(rc, (load().eps_last_error(None) or b"").decode())
    return p


def cooley_segment_length(n_steps: int, n_items: int) -> int:
    lib = load()
    lib.eps_cooley_segment_length.restype = C.c_uint32
    return int(lib.eps_cooley_segment_length(C.c_uint32(n_steps), C.c_uint64(n_items)))


def _vec(x, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))


class Context:
    """One ``eps_ctx`` (one CUDA device, one stream)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.eps_ctx_create(device, C.byref(self.h))
        if rc != EPS_OK:
            raise EpsError(rc, (self.lib.eps_last_error(None) or b"").decode())
        self.n_curves = 0
        self.n_points = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.eps_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != EPS_OK:
            raise EpsError(rc, (self.lib.eps_last_error(self.h) or b"").decode())

    def sync(self):
        self._ck(self.lib.eps_sync(self.h))

    def request_stop(self):
        """Callable from any thread while another thread runs a compute call on this context."""
        self._ck(self.lib.eps_request_stop(self.h))

    def reset_stop(self):
        self._ck(self.lib.eps_reset_stop(self.h))

    def device_bytes(self) -> int:
        n = C.c_uint64()
        self._ck(self.lib.eps_ctx_device_bytes(self.h, C.byref(n)))
        return n.value

    def trim(self, drop_potentials: bool = False):
        self._ck(self.lib.eps_ctx_trim(self.h, C.c_int(1 if drop_potentials else 0)))
        if drop_potentials:
            self.n_curves = 0

    def set_potentials(self, V: np.ndarray, scale) -> None:
        V = np.ascontiguousarray(np.atleast_2d(V), dtype=np.float64)
        scale = _vec(scale, V.shape[0])
        self._ck(self.lib.eps_set_potentials(self.h, _ptr(V, np.float64), C.c_uint32(V.shape[0]),
                                             C.c_uint32(V.shape[1]), _ptr(scale, np.float64)))
        self.n_curves = V.shape[0]
        self.n_points = V.shape[1]

    def set_potentials_rot(self, V: np.ndarray, scale, r_min, grid_step, J) -> None:
        """Resident curves c*len(J) + j: V_J = V + J(J+1) (h^2 / 12 s) / r^2, expanded on the device."""
        V = np.ascontiguousarray(np.atleast_2d(V), dtype=np.float64)
        nC = V.shape[0]
        scale, r_min, grid_step = _vec(scale, nC), _vec(r_min, nC), _vec(grid_step, nC)
        J = np.ascontiguousarray(np.atleast_1d(J), dtype=np.uint32)
        self._ck(self.lib.eps_set_potentials_rot(self.h, _ptr(V, np.float64), C.c_uint32(nC), C.c_uint32(V.shape[1]),
                                                 _ptr(scale, np.float64), _ptr(r_min, np.float64),
                                                 _ptr(grid_step, np.float64), _ptr(J, np.uint32), C.c_uint32(J.size)))
        self.n_curves = nC * J.size
        self.n_points = V.shape[1]

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """Uninitialised numpy array in page-locked host memory (eps_host_alloc); freed with the array."""
        import weakref

        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._ck(self.lib.eps_host_alloc(self.h, C.c_size_t(max(n, 1)), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        lib, addr = self.lib, p.value
        weakref.finalize(buf, lambda: lib.eps_host_free(None, C.c_void_p(addr)))
        return arr

    def curve_info(self, curve: int = 0) -> CurveInfo:
        ci = CurveInfo()
        self._ck(self.lib.eps_get_curve_info(self.h, C.c_uint32(curve), C.byref(ci)))
        return ci

    def _outs(self, nE, nodes, tails):
        shape = (self.n_curves, nE)
        n = np.empty(shape, dtype=np.uint32) if nodes else None
        m = np.empty(shape, dtype=np.float64) if tails else None
        x = np.empty(shape, dtype=np.int32) if tails else None
        return n, m, x

    def sweep(self, E: np.ndarray, nodes: bool = True, tails: bool = True):
        E = np.ascontiguousarray(np.atleast_2d(E), dtype=np.float64)
        assert E.shape[0] == self.n_curves
        n, m, x = self._outs(E.shape[1], nodes, tails)
        self._ck(self.lib.eps_sweep(self.h, _ptr(E, np.float64), C.c_uint64(E.shape[1]),
                                    _ptr(n, np.uint32), _ptr(m, np.float64), _ptr(x, np.int32)))
        return n, m, x

    def sweep_uniform(self, E_lo, E_hi, nE: int, nodes: bool = True, tails: bool = True):
        lo, hi = _vec(E_lo, self.n_curves), _vec(E_hi, self.n_curves)
        n, m, x = self._outs(nE, nodes, tails)
        self._ck(self.lib.eps_sweep_uniform(self.h, _ptr(lo, np.float64), _ptr(hi, np.float64),
                                            C.c_uint64(nE), _ptr(n, np.uint32), _ptr(m, np.float64),
                                            _ptr(x, np.int32)))
        return n, m, x

    def sweep_grid(self, E0, dE, j0: int, nE: int, nodes: bool = True, tails: bool = True, out_nodes: np.ndarray | None = None):
        """Affine grid E_j = E0 + (j0 + j) dE (a slice of a global uniform grid).  out_nodes: the caller's
        [n_curves, nE] uint32 buffer for the node counts (page-locked, ``pinned_empty``, for a full-rate D2H)."""
        a, b = _vec(E0, self.n_curves), _vec(dE, self.n_curves)
        n, m, x = self._outs(nE, nodes and out_nodes is None, tails)
        if nodes and out_nodes is not None:
            assert out_nodes.dtype == np.uint32 and out_nodes.shape == (self.n_curves, nE) and out_nodes.flags.c_contiguous
            n = out_nodes
        self._ck(self.lib.eps_sweep_grid(self.h, _ptr(a, np.float64), _ptr(b, np.float64), C.c_uint32(j0),
                                         C.c_uint64(nE), _ptr(n, np.uint32), _ptr(m, np.float64),
                                         _ptr(x, np.int32)))
        return n, m, x

    SOLVE_COOLEY, SOLVE_OPEN_TAIL = 1, 2

    def solve_levels_grid(self, E0, dE, j0: int, n_coarse: int, v_min: int, v_max: int, refine_points: int,
                          rel_tol: float = 1e-12, max_rounds: int = 8, flags: int = 0):
        """-> (levels[nC, nlev], widths[nC, nlev], n_last[nC], n_first[nC])"""
        a, b = _vec(E0, self.n_curves), _vec(dE, self.n_curves)
        nlev = v_max - v_min + 1
        p = SolveParams(v_min, v_max, n_coarse, refine_points, max_rounds, flags, rel_tol)
        levels = np.empty((self.n_curves, nlev), dtype=np.float64)
        widths = np.empty((self.n_curves, nlev), dtype=np.float64)
        nl = np.empty(self.n_curves, dtype=np.uint32)
        nf = np.empty(self.n_curves, dtype=np.uint32)
        self._ck(self.lib.eps_solve_levels_grid(self.h, C.byref(p), _ptr(a, np.float64), _ptr(b, np.float64),
                                                C.c_uint32(j0), _ptr(levels, np.float64),
                                                _ptr(widths, np.float64), _ptr(nl, np.uint32),
                                                _ptr(nf, np.uint32)))
        return levels, widths, nl, nf

    def solve_levels(self, E_lo, E_hi, n_coarse: int, v_min: int, v_max: int, refine_points: int,
                     rel_tol: float = 1e-12, max_rounds: int = 8, flags: int = 0):
        """-> (levels[nC, nlev], widths[nC, nlev], n_below[nC]); flags=SOLVE_COOLEY: matching iteration"""
        lo, hi = _vec(E_lo, self.n_curves), _vec(E_hi, self.n_curves)
        nlev = v_max - v_min + 1
        p = SolveParams(v_min, v_max, n_coarse, refine_points, max_rounds, flags, rel_tol)
        levels = np.empty((self.n_curves, nlev), dtype=np.float64)
        widths = np.empty((self.n_curves, nlev), dtype=np.float64)
        nb = np.empty(self.n_curves, dtype=np.uint32)
        self._ck(self.lib.eps_solve_levels(self.h, C.byref(p), _ptr(lo, np.float64),
                                           _ptr(hi, np.float64), _ptr(levels, np.float64),
                                           _ptr(widths, np.float64), _ptr(nb, np.uint32)))
        return levels, widths, nb

    def wavefunctions(self, E, grid_step):
        """E[nC, nlev] (NaN rows skipped) -> (psi[nC, nlev, N], match_index[nC, nlev])"""
        E = np.ascontiguousarray(np.atleast_2d(E), dtype=np.float64)
        assert E.shape[0] == self.n_curves
        h = _vec(grid_step, self.n_curves)
        N = self.n_points
        psi = np.empty((self.n_curves, E.shape[1], N), dtype=np.float64)
        mi = np.empty(E.shape, dtype=np.uint32)
        self._ck(self.lib.eps_wavefunctions(self.h, _ptr(E, np.float64), C.c_uint32(E.shape[1]),
                                            _ptr(h, np.float64), _ptr(psi, np.float64), _ptr(mi, np.uint32)))
        return psi, mi

    def level_corrections(self, E, grid_step) -> np.ndarray:
        """E[nC, nlev] -> dE[nC, nlev]: first-order (Cooley / Rayleigh-quotient) energy corrections."""
        E = np.ascontiguousarray(np.atleast_2d(E), dtype=np.float64)
        assert E.shape[0] == self.n_curves
        h = _vec(grid_step, self.n_curves)
        dE = np.empty(E.shape, dtype=np.float64)
        self._ck(self.lib.eps_level_corrections(self.h, _ptr(E, np.float64), C.c_uint32(E.shape[1]),
                                                _ptr(h, np.float64), _ptr(dE, np.float64)))
        return dE

    def spline_resample(self, r, V, r_min: float, r_max: float, n_points: int) -> np.ndarray:
        """Natural cubic spline through (r, V), evaluated on the device on n_points uniform points."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        assert r.shape == V.shape and r.ndim == 1
        out = np.empty(n_points, dtype=np.float64)
        self._ck(self.lib.eps_spline_resample(self.h, _ptr(r, np.float64), _ptr(V, np.float64), C.c_uint32(r.size),
                                              C.c_double(r_min), C.c_double(r_max), C.c_uint32(n_points),
                                              _ptr(out, np.float64)))
        return out

    OPT_SCAN_SEGMENTS, OPT_SCAN_EXACT, OPT_CBANK, OPT_CBANK_SHAPE, OPT_CBANK_PDL, OPT_PREP_PARTS, OPT_FORM, OPT_PACK128 = 1, 2, 3, 4, 5, 6, 7, 8
    OPT_CBANK_GROUP, OPT_SCAN_COMBINE = 9, 10
    CNT_SCAN_LAUNCHES, CNT_SCAN_FLAGGED, CNT_CBANK_LAUNCHES = 1, 2, 3

    def set_option(self, option: int, value: int) -> None:
        self._ck(self.lib.eps_set_option(self.h, C.c_int(option), C.c_int64(value)))

    def counter(self, which: int) -> int:
        v = C.c_uint64()
        self._ck(self.lib.eps_get_counter(self.h, C.c_int(which), C.byref(v)))
        return v.value

    def timer_start(self):
        self._ck(self.lib.eps_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.eps_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def stats(self) -> Stats:
        s = Stats()
        self._ck(self.lib.eps_stats_get(self.h, C.byref(s)))
        return s

    def stats_reset(self):
        self._ck(self.lib.eps_stats_reset(self.h))

    def l2_flush(self):
        self._ck(self.lib.eps_l2_flush(self.h))

    def fp64_probe(self):
        t, ms = C.c_double(), C.c_float()
        self._ck(self.lib.eps_fp64_probe(self.h, C.byref(t), C.byref(ms)))
        return t.value, ms.value


SHARD_CURVES, SHARD_ENERGY = 0, 1


class Group:
    """``eps_group``: several CUDA devices in one process -- one context and one host thread per device,
    levels gathered on the first device with peer copies (include/epseon_cuda.h)."""

    def __init__(self, devices):
        self.lib = load()
        devs = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p()
        rc = self.lib.eps_group_create(devs, C.c_uint32(len(devices)), C.byref(self.h))
        if rc != EPS_OK:
            raise EpsError(rc, (self.lib.eps_last_error(None) or b"").decode())
        self.devices, self.n_curves, self.n_points = list(devices), 0, 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.eps_group_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != EPS_OK:
            raise EpsError(rc, (self.lib.eps_group_last_error(self.h) or b"").decode())

    @property
    def size(self) -> int:
        return int(self.lib.eps_group_size(self.h))

    def last_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.eps_group_last_ms(self.h, C.byref(ms)))
        return ms.value

    def set_option(self, option: int, value: int) -> None:
        self._ck(self.lib.eps_group_set_option(self.h, C.c_int(option), C.c_int64(value)))

    def counter(self, rank: int, which: int) -> int:
        v = C.c_uint64()
        rc = self.lib.eps_get_counter(C.c_void_p(self.lib.eps_group_ctx(self.h, rank)), C.c_int(which), C.byref(v))
        self._ck(rc)
        return v.value

    def set_potentials(self, V: np.ndarray, scale, shard: int) -> None:
        V = np.ascontiguousarray(np.atleast_2d(V), dtype=np.float64)
        scale = _vec(scale, V.shape[0])
        self._ck(self.lib.eps_group_set_potentials(self.h, _ptr(V, np.float64), C.c_uint32(V.shape[0]),
                                                   C.c_uint32(V.shape[1]), _ptr(scale, np.float64), C.c_int(shard)))
        self.n_curves, self.n_points = V.shape

    def sweep_uniform(self, E_lo, E_hi, nE: int, nodes: bool = True):
        lo, hi = _vec(E_lo, self.n_curves), _vec(E_hi, self.n_curves)
        n = np.empty((self.n_curves, nE), dtype=np.uint32) if nodes else None
        self._ck(self.lib.eps_group_sweep_uniform(self.h, _ptr(lo, np.float64), _ptr(hi, np.float64), C.c_uint64(nE),
                                                  _ptr(n, np.uint32)))
        return n

    def solve_levels(self, E_lo, E_hi, n_coarse: int, v_min: int, v_max: int, refine_points: int,
                     rel_tol: float = 1e-12, max_rounds: int = 8, flags: int = 0):
        """-> (levels[nC, nlev], widths[nC, nlev], n_below[nC]) of the WHOLE job"""
        lo, hi = _vec(E_lo, self.n_curves), _vec(E_hi, self.n_curves)
        nlev = v_max - v_min + 1
        p = SolveParams(v_min, v_max, n_coarse, refine_points, max_rounds, flags, rel_tol)
        levels = np.empty((self.n_curves, nlev), dtype=np.float64)
        widths = np.empty((self.n_curves, nlev), dtype=np.float64)
        nb = np.empty(self.n_curves, dtype=np.uint32)
        self._ck(self.lib.eps_group_solve_levels(self.h, C.byref(p), _ptr(lo, np.float64), _ptr(hi, np.float64),
                                                 _ptr(levels, np.float64), _ptr(widths, np.float64), _ptr(nb, np.uint32)))
        return levels, widths, nb


class Mailbox:
    """``eps_mailbox``: gather of small per-rank results into rank 0's device memory across PROCESSES
    (CUDA IPC + peer writes).  Rank 0: ``Mailbox.create(ctx, world, nbytes)`` -> ``.handle`` (64 bytes
    to hand to the other ranks by any means); rank r > 0: ``Mailbox.open(ctx, handle, world, r, nbytes)``."""

    def __init__(self, ctx: "Context", h, world: int, rank: int, handle: bytes):
        self.ctx, self.h, self.world, self.rank, self.handle = ctx, h, world, rank, handle
        self.slot = int(ctx.lib.eps_mailbox_slot_bytes(h))

    @classmethod
    def create(cls, ctx: "Context", world: int, bytes_per_rank: int) -> "Mailbox":
        h, buf = C.c_void_p(), (C.c_ubyte * 64)()
        ctx._ck(ctx.lib.eps_mailbox_create(ctx.h, C.c_uint32(world), C.c_size_t(bytes_per_rank), C.byref(h), buf))
        return cls(ctx, h, world, 0, bytes(buf))

    @classmethod
    def open(cls, ctx: "Context", handle: bytes, world: int, rank: int, bytes_per_rank: int) -> "Mailbox":
        h, buf = C.c_void_p(), (C.c_ubyte * 64).from_buffer_copy(handle)
        ctx._ck(ctx.lib.eps_mailbox_open(ctx.h, buf, C.c_uint32(world), C.c_uint32(rank), C.c_size_t(bytes_per_rank), C.byref(h)))
        return cls(ctx, h, world, rank, handle)

    def post_levels(self, seq: int) -> None:
        self.ctx._ck(self.ctx.lib.eps_mailbox_post_levels(self.h, C.c_uint32(seq)))

    def post(self, payload: np.ndarray, seq: int) -> None:
        a = np.ascontiguousarray(payload)
        self.ctx._ck(self.ctx.lib.eps_mailbox_post(self.h, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes), C.c_uint32(seq)))

    def collect(self, seq: int, nbytes: int, timeout_s: float = 60.0) -> np.ndarray:
        """Rank 0: -> uint8 array [world, nbytes] (the head of every slot) once every rank has posted `seq`."""
        out = np.empty((self.world, nbytes), dtype=np.uint8)
        self.ctx._ck(self.ctx.lib.eps_mailbox_collect(self.h, C.c_uint32(seq), out.ctypes.data_as(C.c_void_p),
                                                      C.c_size_t(nbytes), C.c_double(timeout_s)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.eps_mailbox_destroy(self.h)
            self.h = None
