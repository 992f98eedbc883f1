"""GPU edge cases of the C-ABI (libupcgpu.so) against the CPU oracle: degenerate grids (one mass row, one rapidity
column, a single cell), empty inputs on every vector entry point, calls made out of order, parameter blocks the
library must refuse.  The reference has no such tests of its own (no unit tests at all); the cases
are the ones its grid driver and event loop admit (src/UpcCrossSection.cpp:536-550 loops over whatever BINS_M /
BINS_Y say, src/UpcGenerator.cpp:867-889 runs NEVENTS = 0 as an empty loop).
"""
import dataclasses

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_POINT = 1e-9
RTOL_FF = 1e-7


@pytest.fixture(scope="module")
def capi():
    from upcgen_b200 import capi as m
    m.lib()
    return m


def _gpu(capi, name, extra):
    from upcgen_b200.config import named_config
    P = named_config(name, extra)
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    return P, g


@pytest.mark.parametrize("base,extra,tol", [
    ("cfg1", "BINS_M 1\nBINS_Y 1\n", RTOL_POINT),                        # a single cell (m = MMIN, y = YMIN)
    ("cfg1", "BINS_M 1\nBINS_Y 2\n", RTOL_POINT),                        # one mass row; y = -6 and the self-mirrored y = 0
    ("cfg1", "BINS_M 3\nBINS_Y 1\n", RTOL_POINT),                        # one rapidity column
    ("cfg2", "BINS_M 1\nBINS_Y 1\n", RTOL_FF),                           # the same with QAGS rows and the breakup table
    ("cfg2", "BINS_M 2\nBINS_Y 3\nYMIN -1\nYMAX 2\n", RTOL_FF),          # asymmetric, fewer rows than a warp has lanes
    ("cfg2", "BINS_M 33\nBINS_Y 2\n", RTOL_FF),                          # one full lane group of mass rows plus one row
    ("cfg3", "BINS_Y 1\n", RTOL_POINT),                                  # polarised pair of tables, one column (grid forced to 1000 m rows)
])
def test_degenerate_grids_every_cell_vs_oracle(capi, get_oracle, base, extra, tol):
    P, g = _gpu(capi, base, extra)
    _, o = get_oracle(base, extra)
    try:
        table = g.fill_lumi()
        st = g.fill_stats()
        assert st["qags_errors"] == 0
        if base == "cfg3":
            ref = o.fill_lumi(im_step=37)
            sel = np.isfinite(ref[0])
            assert sel.sum() >= 27
            e = max(np.max(np.abs(table[0][sel] / ref[0][sel] - 1)), np.max(np.abs(table[1][sel] / ref[1][sel] - 1)))
        else:
            ref = o.fill_lumi()
            assert table.shape == (P.nm, P.ny) and np.all(np.isfinite(table)) and np.all(table > 0)
            e = np.max(np.abs(table / ref - 1))
        print(base, "grid", P.nm, "x", P.ny, "max rel", e)
        assert e < tol
    finally:
        g.close()


def test_single_cell_table_through_fold_sampler_and_events(capi, get_oracle):
    """The whole chain on a 1 x 1 grid with one z bin: the fold is one multiply, the 2-D CDF is (0, 1), every candidate
    falls into bin (0, 0), and the events equal the oracle's on the same Philox slots."""
    extra = "BINS_M 1\nBINS_Y 1\nBINS_Z 1\n"
    P, g = _gpu(capi, "cfg1", extra)
    _, o = get_oracle("cfg1", extra)
    try:
        lumi = g.fill_lumi()
        m = P.mmin + P.dm * np.arange(P.nm)
        cs, _, tot = g.fold_sigma(sig_m=o.sigma_m(m))
        ocs, _, otot = o.fold(lumi)
        assert cs.shape == (1, 1) and np.array_equal(cs, ocs) and tot == pytest.approx(otot, rel=1e-12)
        cszm = o.cs_zm(0)
        assert cszm.shape == (1, 1)
        g.sampler_build(cszm=cszm)
        s2, sz, _ = g.sampler_cdf()
        assert np.array_equal(s2, [0., 1.]) and np.array_equal(sz.ravel(), [0., 1.])
        k, yb, mb, y, mm = g.sample_ym(np.array([[0., 0.], [0.5, 0.25], [1 - 2.0 ** -32, 1 - 2.0 ** -32]]))
        assert np.all(k == 0) and np.all(yb == 0) and np.all(mb == 0)
        assert np.array_equal(y, P.ymin + np.array([0., 0.5, 1 - 2.0 ** -32]) * (P.ymax - P.ymin))
        ev = g.generate(5, 0, 64)
        for i in range(64):
            acc, pdg, stt, mo, p4, aux = o.generate_event(5, i, s2, sz, None, None)
            assert ev["npart"][i] == len(pdg) and np.array_equal(ev["pdg"][i, :len(pdg)], pdg)
            assert ev["aux"][i, 0] == aux[0] and ev["aux"][i, 1] == aux[1] and ev["aux"][i, 2] == aux[2]
            if len(pdg):
                assert np.max(np.abs(ev["p4"][i, :len(pdg)] - p4)) <= 1e-9 * np.abs(p4).max()
    finally:
        g.close()


def test_empty_inputs_are_no_ops(capi):
    """n = 0 on every vector entry point: UPCGPU_OK, nothing written, nothing launched with an empty grid (a CUDA
    'invalid configuration' would surface as an error on the next call, which the last lines check)."""
    P, g = _gpu(capi, "cfg2", "BINS_M 8\nBINS_Y 5\n")
    try:
        z = np.zeros(0)
        assert g.eval_table(0, z).size == 0
        assert g.breakup_raw(z, 2).size == 0
        assert g.flux_point(z, z).size == 0
        out, ne = g.flux_form(z, z, with_neval=True)
        assert out.size == 0 and ne.size == 0
        assert g.lumi_cells(z, z).size == 0
        fp, fn = g.photon_flux(z, z)
        assert fp.size == 0 and fn.size == 0
        lumi = g.fill_lumi()
        m = P.mmin + P.dm * np.arange(P.nm)
        g.fold_sigma(sig_m=np.ones_like(m))
        g.sampler_build(cszm=np.ones((P.nm, P.nz)))
        k, yb, mb, y, mm = g.sample_ym(np.zeros((0, 2)))
        assert k.size == 0 and y.size == 0
        assert g.sample_z(np.zeros(0, np.int32), z).size == 0
        ev = g.generate(1, 0, 0)
        assert ev["n_accepted"] == 0 and ev["npart"].size == 0
        ev = g.generate_packed(1, 0, 0)
        assert ev["n_accepted"] == 0
        assert g.generate_device(1, 0, 0) == 0
        # the context is still healthy: a real call after the empty ones
        again = g.fill_lumi()
        assert np.array_equal(again, lumi)
        assert g.generate(1, 0, 16)["npart"].size == 16
    finally:
        g.close()


def test_calls_out_of_order_fail_loudly(capi):
    """Events before a sampler, a sampler before a fold: an error code and a message, never a crash or stale data."""
    P, g = _gpu(capi, "cfg1", "BINS_M 4\nBINS_Y 3\n")
    try:
        with pytest.raises(capi.UpcGpuError):
            g.generate(1, 0, 8)
        with pytest.raises(capi.UpcGpuError):
            g.sample_ym(np.zeros((4, 2)))
        with pytest.raises(capi.UpcGpuError):
            g.sampler_build(cszm=np.ones((P.nm, P.nz)))      # no folded table on the device yet
        g.fill_lumi()
        g.fold_sigma(sig_m=np.ones(P.nm))
        g.sampler_build(cszm=np.ones((P.nm, P.nz)))
        assert g.generate(1, 0, 8)["npart"].size == 8
    finally:
        g.close()


@pytest.mark.parametrize("field,value", [
    ("nm", 0), ("ny", 0), ("nm", -3), ("mmax", 1.0), ("ymax", -7.0), ("mmin", 0.0), ("breakup_mode", 0),
    ("breakup_mode", 5), ("Z", 0), ("A", 0), ("sqrts", -1.0),
])
def test_create_refuses_bad_parameter_blocks(capi, field, value):
    """upcgpu_create validates what UpcGenerator::init would have run into later (a zero-sized grid, an empty mass or
    rapidity range, an unknown breakup mode): UPCGPU_EINVAL and a message, no context."""
    from upcgen_b200.config import named_config
    P = named_config("cfg1", "BINS_M 4\nBINS_Y 3\n")
    cp = capi.to_cparams(P)
    names = [f for f, _ in cp._fields_]
    assert field in names, names
    setattr(cp, field, value)
    import ctypes as C
    L = capi.lib()
    h = C.c_void_p()
    rc = L.upcgpu_create(C.byref(cp), 0, C.byref(h))
    assert rc != 0 and not h.value
    msg = L.upcgpu_last_error(None)
    assert msg and b"upcgpu_create" in msg
