"""The C++ host facade (upcgen_b200/host: UpcCrossSection / UpcSampler / UpcGenerator with the
reference's names) and the upcgen command line, on the GPU, against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "upcgen_b200", "host")


def _build():
    subprocess.check_call(["make", "-s", "-C", HOST])
    exe = os.path.join(ROOT, "tests", "cpp", "facade_check")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", HOST, "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "facade_check.cpp"), "-L", HOST, "-lupcgen_host",
                           "-L", os.path.join(ROOT, "upcgen_b200"), "-lupcgpu", f"-Wl,-rpath,{HOST}",
                           f"-Wl,-rpath,{os.path.join(ROOT, 'upcgen_b200')}"])
    return exe


def test_facade_classes_match_oracle(tmp_path, oracle_mod):
    from upcgen_b200.config import named_config
    exe = _build()
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    P = named_config("cfg2", "BINS_M 24\nBINS_Y 10\n")
    o = oracle_mod.Oracle(P)
    assert d["rho0"] == pytest.approx(o.rho0(), rel=1e-13)
    assert d["fluxPoint"] == pytest.approx(o.flux_point(10.0, 1.0), rel=1e-12)
    assert d["fluxForm"] == pytest.approx(o.flux_form(5.0, 1.0), rel=1e-9)
    assert d["breakup"] == pytest.approx(o.breakup_raw([15.0], 2)[0], abs=1e-13)
    assert d["lumi"] == pytest.approx(o.lumi(10.0, 0.5), rel=1e-9)
    lumi = o.fill_lumi()
    cs, _, tot = o.fold(lumi)
    assert d["totCS"] == pytest.approx(tot, rel=1e-9)
    assert d["cs00"] == pytest.approx(cs[0, 0], rel=1e-9) and d["cs_last"] == pytest.approx(cs[-1, -1], rel=1e-9)
    assert abs(d["sum_last"] - 1) < 1e-12
    ye = P.ymin + P.dy * np.arange(P.ny + 1)
    me = P.mmin + P.dm * np.arange(P.nm + 1)
    for y, m, yb, mb in d["samples"]:
        assert ye[0] <= y < ye[-1] and me[0] <= m < me[-1]
        assert yb == oracle_mod.get_bin(P.ny, y, ye[0], ye[-1]) and mb == oracle_mod.get_bin(P.nm, m, me[0], me[-1])
    # 1-D sampler: the CDF is GSL's sequential one; draws avoid the empty bin [2,3)
    assert d["s1sum"] == list(oracle_mod.pdf_init(np.array([1., 2., 0., 4., 3.])))
    assert all(0 <= v < 5 and not (2 <= v < 3) for v in d["s1"])
    # the cache file written by prepareTwoPhotonLumi is picked up by a second run
    assert os.path.exists(tmp_path / "twoPhotonLumi.root")
    out2 = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert "Found pre-calculated" in out2.stderr
    assert json.loads(out2.stdout.strip().splitlines()[-1])["totCS"] == d["totCS"]


def test_upcgen_cli_writes_hepmc(tmp_path):
    subprocess.check_call(["make", "-s", "-C", HOST])
    par = """NUCLEUS_Z 82
NUCLEUS_A 208
SQRTS 5020
PROC_ID 13
NEVENTS 3000
DO_PT_CUT 1
PT_MIN 0.5
MMIN 4
MMAX 30
BINS_M 30
BINS_Y 14
BINS_Z 50
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
SEED 4242
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "my.in").write_text(par)
    r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "my.in", "-nthreads", "4"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Number of rejected events" in r.stderr
    lines = (tmp_path / "events.hepmc").read_text().splitlines()
    assert lines[0] == "HepMC::Version 3.02.04" and lines[1] == "HepMC::Asciiv3-START_EVENT_LISTING"
    assert lines[-1] == "HepMC::Asciiv3-END_EVENT_LISTING"
    ev = [l for l in lines if l.startswith("E ")]
    assert len(ev) == 3000 and ev[0] == "E 0 0 2" and ev[-1].startswith("E 2999 ")
    parts = [l.split() for l in lines if l.startswith("P ")]
    assert len(parts) == 6000
    p = np.array([[float(x) for x in q[4:9]] for q in parts])
    pdg = np.array([int(q[3]) for q in parts])
    assert set(np.abs(pdg)) == {13} and pdg[0::2].tolist() == (-pdg[1::2]).tolist()
    assert np.all(np.hypot(p[:, 0], p[:, 1]) >= 0.5)                       # PT_MIN cut
    assert np.allclose(p[:, 4], 0.1056583745, atol=2e-6)                   # printed mass
    pair = p[0::2, :4] + p[1::2, :4]
    minv = np.sqrt(pair[:, 3] ** 2 - pair[:, 0] ** 2 - pair[:, 1] ** 2 - pair[:, 2] ** 2)
    assert minv.min() >= 4 - 1e-6 and minv.max() <= 30 + 1e-6


def test_upcgen_cli_reads_a_reference_style_lumi_cache_and_writes_events_root(tmp_path, oracle_mod):
    """(f2) twoPhotonLumi.root -- here written by the ROOT-less writer from the ORACLE's table, i.e. the file a run of
    the reference would have left behind -- is picked up instead of recomputing (src/UpcCrossSection.cpp:481-491), and
    the cross section is the oracle's.  (f3) USE_ROOT_OUTPUT 1 writes events.root with the tree "particles" and its nine
    branches (src/UpcGenerator.cpp:842-857), consistent with events.hepmc written by the same run."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_root_file import read_tree
    from upcgen_b200 import capi
    from upcgen_b200.config import UpcParams
    subprocess.check_call(["make", "-s", "-C", HOST])
    par = """NUCLEUS_Z 82
NUCLEUS_A 208
SQRTS 5020
PROC_ID 13
NEVENTS 2000
MMIN 4
MMAX 30
BINS_M 20
BINS_Y 10
BINS_Z 50
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 0
SEED 7
USE_ROOT_OUTPUT 1
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "my.in").write_text(par)
    P = UpcParams.from_text(par).init()
    o = oracle_mod.Oracle(P)
    lumi = o.fill_lumi()
    _, _, tot = o.fold(lumi)
    capi.root_write_th2d(str(tmp_path / "twoPhotonLumi.root"), {"hD2LDMDY": lumi}, P.nm, P.mmin, P.mmax, P.ny, P.ymin, P.ymax)
    r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "my.in", "-debug", "1"], cwd=tmp_path, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Found pre-calculated unpolarized 2D luminosity" in r.stderr
    line = [l for l in r.stdout.splitlines() if "total cross section" in l][0]
    assert float(line.split()[4]) == pytest.approx(tot, rel=1e-5)     # printed with 6 digits
    t = read_tree((tmp_path / "events.root").read_bytes(), "particles")
    assert t["entries"] == 4000 and list(t["cols"]) == ["eventNumber", "pdgCode", "particleID", "statusID", "motherID",
                                                        "px", "py", "pz", "e"]
    c = {k: v[1] for k, v in t["cols"].items()}
    assert c["eventNumber"].tolist() == np.repeat(np.arange(2000), 2).tolist()
    assert c["particleID"].tolist() == [1, 2] * 2000 and set(c["statusID"]) == {23} and set(c["motherID"]) == {0}
    # -debug 1: the cross section table and its projections next to the tree (src/UpcGenerator.cpp:900-917)
    from test_root_file import read_keys, key_data, parse_hist
    fb = (tmp_path / "events.root").read_bytes()
    _, keys = read_keys(fb)
    names = [(k["cls"], k["name"]) for k in keys if k["cls"] in ("TH1D", "TH2D", "TTree")]
    assert names == [("TH1D", "hNucCSYM_py"), ("TH1D", "hNucCSYM_px"), ("TH2D", "hNucCSYM"), ("TTree", "particles")], names
    h2 = parse_hist(key_data(fb, [k for k in keys if k["name"] == "hNucCSYM"][0]), "TH2D")
    ocs, _, _ = o.fold(lumi)
    assert np.array_equal(h2["cells"].reshape(P.nm + 2, P.ny + 2)[1:-1, 1:-1], ocs.T)
    assert np.allclose(h2["axes"][0]["edges"], P.ymin + P.dy * np.arange(P.ny + 1), rtol=0, atol=1e-12)
    assert np.allclose(h2["axes"][1]["edges"], P.mmin + P.dm * np.arange(P.nm + 1), rtol=0, atol=1e-12)
    hm = parse_hist(key_data(fb, [k for k in keys if k["name"] == "hNucCSYM_py"][0]), "TH1D")
    assert np.allclose(hm["cells"][1:-1], ocs.sum(axis=0), rtol=1e-13) and hm["cells"][0] == 0 and hm["cells"][-1] == 0
    hep = [l.split() for l in (tmp_path / "events.hepmc").read_text().splitlines() if l.startswith("P ")]
    assert len(hep) == 4000
    assert [int(q[3]) for q in hep] == c["pdgCode"].tolist()
    for j, name in enumerate(("px", "py", "pz", "e")):
        assert np.allclose([float(q[4 + j]) for q in hep], c[name], rtol=2e-8, atol=1e-12)   # HepMC prints 9 digits


def test_upcgen_cli_pi0_pairs_from_tabulated_cross_sections(tmp_path):
    """PROC_ID 111 end to end through the C++ drop-in: UpcTwoPhotonDipion reads sigma(m) and dsigma/dz from ROOT files
    (cross_sections/pi0pi0, src/UpcTwoPhotonDipion.cpp) -- here files of the same form written by this repository's
    ROOT writer, because the reference's do not travel to the GPU box --, the generator forces its grid (:69-103),
    and every event is a pi0 pair with both pi0 decayed uniformly into photons (:799-803): six particles, two decay
    vertices.  The total cross section equals the fold of the same table through the C-ABI."""
    from upcgen_b200 import capi
    from upcgen_b200.config import UpcParams
    subprocess.check_call(["make", "-s", "-C", HOST])
    xs = tmp_path / "xs" / "pi0pi0"
    xs.mkdir(parents=True)
    mc = 0.025 + 0.05 * np.arange(100)                  # bin centres of the reference's sigma(m) histogram: 100 bins on [0, 5]
    sig_tab = np.where(mc > 0.27, 30.0 / (0.2 + mc) ** 2, 0.0)
    capi.root_write_th1d(str(xs / "cross_section_m.root"), "hCrossSectionM", sig_tab, 0.0, 5.0)
    zc = -0.99 + 0.02 * np.arange(100)
    zm_tab = np.outer(1.0 + 0.5 * zc ** 2, sig_tab)    # [z bin][m bin]: x = z, y = m (UpcTwoPhotonTabulated::calcCrossSectionZM)
    capi.root_write_th2d(str(xs / "cross_section_zm.root"), {"hCrossSectionZM": zm_tab}, 100, -1.0, 1.0, 100, 0.0, 5.0)
    n = 4000
    par = f"""NUCLEUS_Z 82
NUCLEUS_A 208
WS_R 6.68
WS_A 0.447
SQRTS 5020
PROC_ID 111
NEVENTS {n}
YMIN -4
YMAX 4
BINS_Y 32
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
SEED 5
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "pi0.in").write_text(par)
    r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "pi0.in"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=600, env={**os.environ, "UPCGEN_CROSS_SEC_DIR": str(tmp_path / "xs")})
    assert r.returncode == 0, r.stderr[-3000:]
    tot_cli = float([l for l in r.stdout.splitlines() if "total cross section" in l][0].split()[4])
    # the same fold through the C-ABI: sigma(m) looked up as the plug-in does (TAxis::FindBin on the lower bin edges)
    P = UpcParams.from_text(par).init()
    assert (P.nm, P.mmin, P.mmax, P.nz) == (91, 0.275, 5.0, 100)
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    g.fill_lumi()
    m = P.mmin + P.dm * np.arange(P.nm)
    sig = sig_tab[np.minimum((100 * (m - 0.0) / (5.0 - 0.0)).astype(int), 99)]   # TAxis::FindBin: 1 + int(nbins (x - xmin) / (xmax - xmin))
    _, _, tot = g.fold_sigma(sig_m=sig)
    g.close()
    assert tot_cli == pytest.approx(tot, rel=1e-5)        # six printed digits
    lines = (tmp_path / "events.hepmc").read_text().splitlines()
    ev = [l for l in lines if l.startswith("E ")]
    assert len(ev) == n and all(l.split()[2:] == ["2", "6"] for l in ev)
    parts = [l.split() for l in lines if l.startswith("P ")]
    pdg = np.array([int(q[3]) for q in parts]).reshape(n, 6)
    mo = np.array([int(q[2]) for q in parts]).reshape(n, 6)
    st = np.array([int(q[9]) for q in parts]).reshape(n, 6)
    p = np.array([[float(x) for x in q[4:8]] for q in parts]).reshape(n, 6, 4)
    assert np.all(pdg == [111, 111, 22, 22, 22, 22]) and np.all(mo == [0, 0, 1, 1, 2, 2]) and np.all(st == [23, 23, 33, 33, 33, 33])
    assert np.allclose(p[:, 2] + p[:, 3], p[:, 0], rtol=1e-6, atol=1e-7) and np.allclose(p[:, 4] + p[:, 5], p[:, 1], rtol=1e-6, atol=1e-7)
    mass2 = lambda q: q[..., 3] ** 2 - (q[..., :3] ** 2).sum(-1)
    assert np.allclose(np.sqrt(np.maximum(mass2(p[:, :2]), 0)), 0.1349770, atol=2e-4)    # nine printed digits of a boosted pi0


def test_upcgen_cli_honours_the_lumi_lock_file(tmp_path):
    """The lock protocol of prepareTwoPhotonLumi (src/UpcCrossSection.cpp:465-478, :587-591): while another generator's
    .lumiIsCalculated exists in the luminosity directory the run waits (in one-second steps); once it is gone the run
    takes the lock, fills and caches the table, and removes its own lock file."""
    import time
    subprocess.check_call(["make", "-s", "-C", HOST])
    par = """NUCLEUS_Z 82
NUCLEUS_A 208
SQRTS 5020
PROC_ID 13
NEVENTS 10
MMIN 4
MMAX 30
BINS_M 6
BINS_Y 4
BINS_Z 10
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 0
SEED 3
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "my.in").write_text(par)
    lock = tmp_path / ".lumiIsCalculated"
    lock.write_text("")
    err = open(tmp_path / "stderr.txt", "w")
    p = subprocess.Popen([os.path.join(HOST, "upcgen"), "-parfile", "my.in"], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=err, text=True)
    try:
        deadline = time.time() + 120
        while "waiting" not in (tmp_path / "stderr.txt").read_text():      # the tables are made first
            assert p.poll() is None, (tmp_path / "stderr.txt").read_text()
            assert time.time() < deadline
            time.sleep(0.2)
        time.sleep(1.5)
        assert p.poll() is None and not (tmp_path / "twoPhotonLumi.root").exists() and lock.exists()
        lock.unlink()
        out, _ = p.communicate(timeout=120)
    finally:
        if p.poll() is None:
            p.kill()
        err.close()
    assert p.returncode == 0, (tmp_path / "stderr.txt").read_text()
    assert "total cross section" in out
    assert (tmp_path / "twoPhotonLumi.root").exists() and not lock.exists()
    assert len([l for l in (tmp_path / "events.hepmc").read_text().splitlines() if l.startswith("E ")]) == 10
