"""Outward/inward matching (Cooley) level search, EPS_SOLVE_COOLEY.

CPU: the oracle's statement of the iteration against the k-section search and the analytic Morse
spectrum; the open tail on the reference's docs-example curve (docs/examples/example.py:23-30), whose
top levels feel the wall at max_r.  GPU: the CUDA kernel against the oracle (same operations, same
summation order: bit-identical) and against the k-section levels."""
import numpy as np
import pytest

from tests import workloads as W


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def _docs_example():
    De, re, a, rmin, rmax, N, m = 500.0, 2.6, 1.3, 0.0, 10.0, 16500, 87.62
    V = W.morse(De, re, a, rmin, rmax, N)
    s = W.scale(m, m, W.grid_h(rmin, rmax, N))
    return V, s, De, W.morse_levels(De, a, m, m)


def test_oracle_cooley_equals_ksection(oracle_d):
    for w in (W.c1(), W.c2(N=40_000)):
        A, *_ = oracle_d.prep(w["V"], w["s"])
        ref, *_ = oracle_d.solve_levels(A, w["s"], w["E_lo"], w["E_hi"], 2048, 0, 16, 256, 1e-13, 12)
        lev, wid, nb, its = oracle_d.solve_levels_cooley(A, w["s"], w["E_lo"], w["E_hi"], 1024, 0, 16, 1e-12, 30)
        assert nb == 17 and np.all(its <= 6) and np.all(its >= 2)
        assert np.max(np.abs(lev - ref) / ref) < 2e-13
        assert np.max(wid / lev) < 1e-12  # the last correction is below the tolerance


def test_oracle_cooley_keeps_the_bracketed_level(oracle_d):
    """A bracket that holds exactly level v must return level v even from a poor start: brackets as
    wide as the coarse grid allows (64 points over the whole well)."""
    w = W.c1()
    A, *_ = oracle_d.prep(w["V"], w["s"])
    ref, *_ = oracle_d.solve_levels(A, w["s"], w["E_lo"], w["E_hi"], 2048, 0, 16, 256, 1e-13, 12)
    lev, wid, nb, its = oracle_d.solve_levels_cooley(A, w["s"], w["E_lo"], w["E_hi"], 64, 0, 16, 1e-12, 40)
    found = np.isfinite(lev)
    assert found.sum() >= 12  # (levels sharing a coarse interval are reported once, as by the k-section search)
    assert np.max(np.abs(lev[found] - ref[found]) / ref[found]) < 2e-13


def test_open_tail_removes_the_wall_on_the_docs_example(oracle_d):
    """docs/examples/example.py: De = 500, a = 1.3, re = 2.6 on r <= 10.  v = 26 lies 0.98 below De and
    its tail reaches the wall: the box level is 2.6e-6 too high; the open tail brings it to 1.4e-8.
    v = 27 (0.03 below De, turning point at 10.5 > max_r) is not representable on this grid."""
    V, s, De, exact = _docs_example()
    A, *_ = oracle_d.prep(V, s)
    box, _, nb, _ = oracle_d.solve_levels_cooley(A, s, 0.0, De - 1e-6, 8192, 0, 26, 1e-12, 40)
    opn, _, _, its = oracle_d.solve_levels_cooley(A, s, 0.0, De - 1e-6, 8192, 0, 26, 1e-12, 40, open_tail=True)
    assert nb == 27 and len(exact) == 28
    rel_box, rel_opn = (box - exact[:27]) / exact[:27], (opn - exact[:27]) / exact[:27]
    assert 1e-6 < rel_box[26] < 1e-5 and abs(rel_opn[26]) < 5e-8
    assert 1e-10 < rel_box[25] < 1e-9 and abs(rel_opn[25]) < 1e-10
    assert np.max(np.abs(rel_opn[:25] - rel_box[:25])) < 1e-12  # bound states do not feel the difference
    assert np.all(its[:27] <= 8)


@pytest.fixture()
def ctx():
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200 import cabi

    c = cabi.Context(0)
    c.set_option(c.OPT_FORM, 1)
    yield c
    c.close()


@pytest.mark.gpu
def test_cuda_cooley_bit_identical_to_oracle(oracle_d, ctx):
    for w, n_coarse in ((W.c1(), 1024), (W.c2(), 4096), (W.c2(N=30_000), 65)):
        A, *_ = oracle_d.prep(w["V"], w["s"])
        ctx.set_potentials(w["V"], w["s"])
        lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], n_coarse, 0, 16, 1, 1e-12, 40, flags=ctx.SOLVE_COOLEY)
        lev_o, wid_o, nb_o, its = oracle_d.solve_levels_cooley(A, w["s"], w["E_lo"], w["E_hi"], n_coarse, 0, 16, 1e-12, 40)
        assert nb[0] == nb_o
        assert _same_bits(lev[0], lev_o), np.abs(lev[0] - lev_o) / lev_o
        assert _same_bits(wid[0], wid_o)
        ref, _, _ = ctx.solve_levels(w["E_lo"], w["E_hi"], 4096, 0, 16, 256, 1e-13, 12)
        ok = np.isfinite(lev[0])
        assert np.max(np.abs(lev[0][ok] - ref[0][ok]) / ref[0][ok]) < 2e-13


@pytest.mark.gpu
def test_cuda_cooley_batch_and_open_tail(oracle_d, ctx):
    w = W.c4(nC=40, N=6000, nE=512)
    ctx.set_potentials(w["V"], w["s"])
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 512, 0, 7, 1, 1e-12, 40, flags=ctx.SOLVE_COOLEY)
    ref, _, nb_ref = ctx.solve_levels(w["E_lo"], w["E_hi"], 512, 0, 7, 64, 1e-13, 14)
    assert np.array_equal(nb, nb_ref) and np.max(np.abs(lev - ref) / ref) < 2e-13
    from epseon_backend_b200 import cabi

    w = W.c4(nC=160, N=6000, nE=512)  # 1280 (curve, level) items: the batch rule for the segment length
    ctx.set_potentials(w["V"], w["s"])
    lev, wid, nb = ctx.solve_levels(w["E_lo"], w["E_hi"], 512, 0, 7, 1, 1e-12, 40, flags=ctx.SOLVE_COOLEY)
    for c in (0, 17, 159):
        A, *_ = oracle_d.prep(w["V"][c], w["s"])
        L = cabi.cooley_segment_length(A.size, 160 * 8)
        assert L > cabi.cooley_segment_length(A.size, 8)
        lev_o, *_ = oracle_d.solve_levels_cooley(A, w["s"], w["E_lo"][c], w["E_hi"][c], 512, 0, 7, 1e-12, 40, seg_len=L)
        assert _same_bits(lev[c], lev_o)
    V, s, De, exact = _docs_example()
    A, *_ = oracle_d.prep(V, s)
    ctx.set_potentials(V, s)
    opn, _, nb = ctx.solve_levels(0.0, De - 1e-6, 8192, 0, 26, 1, 1e-12, 40, flags=ctx.SOLVE_COOLEY | ctx.SOLVE_OPEN_TAIL)
    opn_o, *_ = oracle_d.solve_levels_cooley(A, s, 0.0, De - 1e-6, 8192, 0, 26, 1e-12, 40, open_tail=True)
    assert _same_bits(opn[0], opn_o) and nb[0] == 27
    assert abs(opn[0][26] - exact[26]) / exact[26] < 5e-8 and abs(opn[0][25] - exact[25]) / exact[25] < 1e-10


@pytest.mark.gpu
def test_cooley_needs_accurate_tables_and_falls_back_on_short_windows(ctx):
    from epseon_backend_b200 import cabi

    w = W.c1()
    ctx.set_option(ctx.OPT_FORM, 0)
    ctx.set_potentials(w["V"], w["s"])
    with pytest.raises(cabi.EpsError) as e:
        ctx.solve_levels(w["E_lo"], w["E_hi"], 512, 0, 3, 1, 1e-12, 40, flags=ctx.SOLVE_COOLEY)
    assert e.value.code == 4
    ctx.set_option(ctx.OPT_FORM, 1)
    V = W.morse(5500.0, 2.2, 1.6, 1.0, 8.0, 200) * 1e-3  # 200-point window: k-section is used instead
    s = W.scale(20.0, 20.0, W.grid_h(1.0, 8.0, 200))
    ctx.set_potentials(V, s)
    a = ctx.solve_levels(float(V.min()), float(V[-1]) * 0.9, 128, 0, 2, 32, 1e-12, 12, flags=ctx.SOLVE_COOLEY)
    b = ctx.solve_levels(float(V.min()), float(V[-1]) * 0.9, 128, 0, 2, 32, 1e-12, 12)
    assert _same_bits(a[0], b[0])


@pytest.mark.gpu
def test_level_search_through_python_api(oracle_d):
    import __graft_entry__ as ge

    ge.build()
    from epseon_backend_b200.device.gpu import _libepseon_gpu as g

    dev = g.EpseonComputeContext.create().get_device_interface(0)
    V, s, De, exact = _docs_example()
    out = {}
    for mode in ("ksection", "cooley", "cooley_open"):
        cfg = (dev.get_task_configurator("float64").set_hardware_config(16500, 8192, 1 << 24)
               .set_morse_potential([g.MorsePotentialConfig(500.0, 2.6, 1.3, 0.0, 10.0, 16500)])
               .set_vibwa_algorithm(87.62, 87.62, 0.1, 1e-6, 0, 26).set_level_search(mode))
        h = dev.submit_task(cfg)
        h.wait()
        assert not h.has_failed(), h.get_status_message()
        out[mode] = np.array(h.get_levels())[0]
    assert np.max(np.abs(out["cooley"] - out["ksection"]) / out["ksection"]) < 2e-12
    assert abs(out["cooley_open"][26] - exact[26]) / exact[26] < 5e-8 < abs(out["ksection"][26] - exact[26]) / exact[26]
    with pytest.raises(RuntimeError):
        dev.get_task_configurator("float64").set_level_search("newton")
