"""Cooperative cancellation (eps_request_stop, bound by TaskHandle::cancel -- reference
task_handle.hpp:136-144) and release of a context's grow-only device memory (eps_ctx_trim)."""
import threading
import time

import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    c = cabi.Context(0)
    yield c
    c.close()


def test_request_stop_interrupts_a_long_sweep(ctx):
    """A 2^23-energy sweep on the 200k grid (~0.4 s, 51 queued chunk launches) is stopped from another
    thread: the call returns EPS_ERR_CANCELLED well before it would have finished, the flag stays up
    until eps_reset_stop, and the context then returns the same bits as before."""
    from epseon_backend_b200 import cabi

    nE = 1 << 23
    w = W.c5(nE=nE)
    ctx.set_potentials(w["V"], w["s"])
    ref, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], 4096, tails=False)
    t0 = time.perf_counter()
    ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, tails=False)
    full = time.perf_counter() - t0

    threading.Timer(0.05, ctx.request_stop).start()
    t0 = time.perf_counter()
    with pytest.raises(cabi.EpsError) as err:
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], nE, tails=False)
    stopped = time.perf_counter() - t0
    assert err.value.code == cabi.EPS_ERR_CANCELLED
    assert stopped < 0.6 * full, (stopped, full)
    with pytest.raises(cabi.EpsError) as err:  # still up
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], 4096, tails=False)
    assert err.value.code == cabi.EPS_ERR_CANCELLED
    ctx.reset_stop()
    again, _, _ = ctx.sweep_uniform(w["E_lo"], w["E_hi"], 4096, tails=False)
    assert np.array_equal(again, ref)


def test_request_stop_between_refinement_rounds(ctx, oracle):
    from epseon_backend_b200 import cabi

    w = W.c2()
    ctx.set_potentials(w["V"], w["s"])
    ctx.request_stop()
    with pytest.raises(cabi.EpsError) as err:
        ctx.solve_levels(w["E_lo"], w["E_hi"], 65536, 0, 16, 4457, 1e-10, 8)
    assert err.value.code == cabi.EPS_ERR_CANCELLED
    ctx.reset_stop()
    lev, _, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-10, 8)
    F, *_ = oracle.prep(w["V"], w["s"])
    lev_o, *_ = oracle.solve_levels(F, w["s"], w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-10, 8)
    assert np.array_equal(lev[0].view(np.uint64), lev_o.view(np.uint64)) and nb[0] == 17


def test_trim_releases_scratch_and_keeps_the_curves(ctx):
    w = W.c1()
    ctx.set_potentials(w["V"], w["s"])
    n0, m0, x0 = ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1 << 16)
    lev = ctx.solve_levels(w["E_lo"], w["E_hi"], 1024, 0, 16, 256, 1e-12, 10)[0]
    h = (10.0 - 0.2) / (w["N"] - 1)
    ctx.wavefunctions(lev, h)
    big = ctx.device_bytes()
    ctx.trim()
    small = ctx.device_bytes()
    assert small < big / 4 and small < (4 << 20), (big, small)
    n1, m1, x1 = ctx.sweep_uniform(w["E_lo"], w["E_hi"], 1 << 16)  # scratch comes back on demand
    assert np.array_equal(n0, n1) and np.array_equal(m0.view(np.uint64), m1.view(np.uint64)) and np.array_equal(x0, x1)
    ctx.trim(drop_potentials=True)
    assert ctx.device_bytes() == 0
    from epseon_backend_b200 import cabi

    with pytest.raises(cabi.EpsError) as err:
        ctx.n_curves = 1
        ctx.sweep_uniform(w["E_lo"], w["E_hi"], 16)
    assert err.value.code == 4  # EPS_ERR_STATE: a new eps_set_potentials is needed


def test_task_cancel_through_python_api():
    """cancel() returns a bool, the task ends with status 'cancelled' (not failed), a later task on the
    same device (same pooled context) is unaffected."""
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200.device.gpu import _libepseon_gpu as g

    dev = g.EpseonComputeContext.create().get_device_interface(0)

    def submit(n_curves, N):
        cfgs = [g.MorsePotentialConfig(5500.0 + c, 2.2, 1.6, 0.4, 10.0, N) for c in range(n_curves)]
        cfg = (dev.get_task_configurator("float64").set_hardware_config(N, 1024, 1 << 24)
               .set_morse_potential(cfgs).set_vibwa_algorithm(20.0, 20.0, 0.1, 1.0, 0, 7))
        return dev.submit_task(cfg)

    h = submit(2048, 16500)  # tens of milliseconds of tabulation + solve
    time.sleep(0.002)
    asked = h.cancel()
    h.wait()
    assert isinstance(asked, bool) and h.is_done() and not h.has_failed()
    if asked:
        assert h.was_cancelled() and h.get_status_message() == "cancelled"
    assert h.cancel() is False  # nothing running any more
    h2 = submit(2, 10000)
    h2.wait()
    assert not h2.has_failed() and not h2.was_cancelled() and h2.get_status_message() == "done"
    assert np.all(np.isfinite(np.array(h2.get_levels())))
    g.release_device_memory()
