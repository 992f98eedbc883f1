"""The oracle against the REFERENCE's own translation units (src/UpcCrossSection.cpp,
src/UpcTwoPhotonDilep.cpp, src/UpcTwoPhotonALP.cpp) compiled unmodified against the GSL/ROOT shim
(oracle/refshim -> oracle/_ref/libupcref.so).  This pins the oracle's restatement of the
reference's own code (loops, grids, clamps, quirks); the third-party numerics behind the shim are
pinned separately (tests/test_oracle_pins.py).  Each case runs in a subprocess because the
reference keeps its tables in process-global state."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference)")

_CASE = r"""
import json, sys, tempfile
sys.path.insert(0, {root!r})
import numpy as np
from oracle import pyoracle, pyref
from upcgen_b200.config import named_config
cfg, extra, with_bk = {cfg!r}, {extra!r}, {with_bk}
P = named_config(cfg, extra)
ref = pyref.Reference(P, with_breakup_table=with_bk)
o = pyoracle.Oracle(P, nbc=0)
L = ref.L
out = {{}}
out["rho0"] = [L.upcref_rho0(), o.rho0()]
out["gtot"] = [L.upcref_gtot(), P.gtot]
gy, gc = ref.gaa(); ob, og, oc, ota = o.gaa()
out["gaa_maxabs"] = float(np.max(np.abs(gy - og))); out["gaa_c_maxabs"] = float(np.max(np.abs(gc - oc)))
out["ff_knots"] = float(max(abs(L.upcref_formfac_knot(i) - o.formfac_table(i, 1)[0][0]) for i in (0, 1, 777, 500000, 999999)))
b = np.exp(np.linspace(np.log(0.4), np.log(300.), 12)); k = np.exp(np.linspace(np.log(5e-3), np.log(3e3), 9))
fp = max(abs(L.upcref_flux_point(float(x), float(y)) / max(o.flux_point(float(x), float(y)), 1e-300) - 1) for x in b for y in k if o.flux_point(float(x), float(y)) > 1e-280)
out["flux_point_rel"] = float(fp)
if not P.is_point:
    bb = np.exp(np.linspace(np.log(0.05 * P.R), np.log(2 * P.R), 7)); kk = np.exp(np.linspace(np.log(1e-2), np.log(1e3), 6))
    out["flux_form_equal"] = bool(all(L.upcref_flux_form(float(x), float(y)) == o.flux_form(float(x), float(y)) for x in bb for y in kk))
if P.breakup_mode > 1:
    bs = [1e-6, 0.334, 6.68, 13.36, 15.0, 19.5, 20.0, 25.0]
    out["breakup_raw_equal"] = bool(all(L.upcref_breakup_raw(x, P.breakup_mode) == o.breakup_raw([x], P.breakup_mode)[0] for x in bs))
    if with_bk:
        xs = list(np.random.default_rng(0).uniform(1e-6, 20, 200)) + [20.0]
        out["breakup_spline_maxabs"] = float(max(abs(L.upcref_breakup_spline(float(x)) - o.breakup_spline([x])[0]) for x in xs))
cells = [(P.mmin, P.ymin), (P.mmin + P.dm * (P.nm - 1), P.ymin + P.dy * (P.ny - 1)), (P.mmin + P.dm * (P.nm // 2), P.ymin + P.dy * (P.ny // 2)), (P.mmin + 3 * P.dm, 0.37)]
if with_bk or P.breakup_mode == 1:
    if P.use_pol:
        out["lumi_equal"] = bool(all(ref.lumi_pol(m, y) == o.lumi_pol(m, y) for m, y in cells))
    else:
        vals = [(L.upcref_lumi(m, y), o.lumi(m, y)) for m, y in cells]
        out["lumi_equal"] = bool(all(a == b for a, b in vals)); out["lumi_vals"] = vals
    if P.proc_id in (11, 13, 15, 51) and P.nm * P.ny <= 400:
        with tempfile.TemporaryDirectory() as d:
            cs, ratio, tot = ref.grid_and_fold(d, nthreads=3)
        if P.use_pol:
            ls, lp = o.fill_lumi(); ocs, oratio, otot = o.fold(None, ls, lp)
            out["ratio_equal"] = bool(np.array_equal(ratio, oratio))
        else:
            ocs, _, otot = o.fold(o.fill_lumi())
        out["fold_equal"] = bool(np.array_equal(cs, ocs)); out["tot_rel"] = abs(tot / otot - 1)
ms = [P.mmin + P.dm * i for i in (0, 1, P.nm // 2, P.nm - 1)]
if P.proc_id in (11, 13, 15, 51):
    out["sigma_equal"] = bool(all(L.upcref_sigma_m(m) == o.sigma_m([m])[0] for m in ms))
    if P.proc_id != 51:
        out["sigma_pol_equal"] = bool(all(L.upcref_sigma_m_pol(m, ps) == o.sigma_m_pol([m], ps)[0] for m in ms for ps in (0, 1)))
        out["sigma_zm_equal"] = bool(all(L.upcref_sigma_zm(z, m) == o.L.upco_sigma_zm(o.h, z, m) and L.upcref_sigma_zm_pol(z, m, 1) == o.L.upco_sigma_zm_pol(o.h, z, m, 1) for m in ms for z in (-0.99, -0.3, 0.0, 0.5)))
print("RESULT " + json.dumps(out))
"""


def run_case(tmp_path, cfg, extra="", with_bk=False, timeout=900):
    script = tmp_path / "case.py"
    script.write_text(_CASE.format(root=ROOT, cfg=cfg, extra=extra, with_bk=with_bk))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=timeout, cwd=tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _common(d):
    assert d["rho0"][0] == d["rho0"][1]
    assert d["gtot"][0] == pytest.approx(d["gtot"][1], rel=1e-15)
    assert d["gaa_maxabs"] == 0.0 and d["gaa_c_maxabs"] == 0.0
    assert d["ff_knots"] == 0.0
    assert d["flux_point_rel"] == 0.0
    assert d.get("sigma_equal", True) and d.get("sigma_pol_equal", True) and d.get("sigma_zm_equal", True)


def test_reference_point_flux_no_breakup(tmp_path):
    """cfg1 physics on a 16x9 grid: tables, fluxes, cells, the OpenMP grid driver and the fold."""
    d = run_case(tmp_path, "cfg1", "BINS_M 16\nBINS_Y 9\n")
    _common(d)
    assert d["lumi_equal"], d.get("lumi_vals")
    assert d["fold_equal"] and d["tot_rel"] < 1e-13


def test_reference_form_factor_flux(tmp_path):
    """cfg2/cfg4 physics without the breakup table: fluxForm through the reference's own integrand
    and gsl_integration_qags call; cells bit-equal."""
    d = run_case(tmp_path, "cfg2", "BREAKUP_MODE 1\nBINS_M 8\nBINS_Y 6\n")
    _common(d)
    assert d["flux_form_equal"] and d["lumi_equal"], d.get("lumi_vals")
    assert d["fold_equal"]


def test_reference_polarised_and_alp(tmp_path):
    d = run_case(tmp_path, "cfg1", "USE_POLARIZED_CS 1\nLEP_A 0.0011\nBINS_M 10\nBINS_Y 6\n")
    _common(d)
    assert d["lumi_equal"] and d["fold_equal"] and d["ratio_equal"]
    d = run_case(tmp_path, "cfg5", "BREAKUP_MODE 1\nBINS_M 6\nBINS_Y 5\n")
    _common(d)
    assert d["lumi_equal"] and d["fold_equal"]


def test_reference_breakup_raw_all_modes(tmp_path):
    """calcBreakupProb itself (STARlight-derived, TMath::BesselK1) for XNXN / 0N0N / 0NXN."""
    for cfg in ("cfg2", "cfg5", "cfg3"):
        d = run_case(tmp_path, cfg, "BINS_M 6\nBINS_Y 5\n")
        assert d["breakup_raw_equal"], cfg


@pytest.mark.slow
def test_reference_full_breakup_table_and_cells(tmp_path):
    """The reference's full 1e6-knot breakup spline (prepareBreakupProb, ~1 min) against the oracle's
    truncated 21 001-knot one: spline values on [0,20] and whole cells with XNXN breakup."""
    d = run_case(tmp_path, "cfg2", "BINS_M 8\nBINS_Y 6\n", with_bk=True, timeout=1800)
    _common(d)
    assert d["breakup_spline_maxabs"] < 1e-15
    assert d["flux_form_equal"] and d["lumi_equal"], d.get("lumi_vals")
    assert d["fold_equal"]
