"""Statistical compatibility of the GPU event stage (Philox-driven, csrc/upc_events.cu) with the reference
(BASELINE.json north_star: "sampled pair kinematics (m, y, pT, eta) must be statistically compatible with the
reference under chi2 / KS tests").

Two kinds of comparison, all through the C-ABI (upcgpu_generate with host buffers):
  * against the distribution the reference DEFINES: chi2 of the sampled (yBin, mBin) counts against cs / sum(cs)
    (src/UpcGenerator.cpp:737, include/UpcSampler.h:118-121); KS of cos(theta) inside single mass bins against the
    cumulative dsigma/dz table (:753); KS of the photon pT against the 5000-bin cumulative pdf (getPhotonPt,
    src/UpcCrossSection.cpp:1021-1051);
  * against a sample DRAWN BY THE REFERENCE'S OWN CODE: src/UpcGenerator.cpp's generateEvent (compiled unmodified
    into oracle/_ref, MT19937 / mt19937_64 streams, reference slot order) fed with the SAME luminosity table the GPU
    computed; two-sample KS on pair mass, rapidity, pair pT and on the pT and eta of the final-state particles.
All samples are seeded, so the p-values are fixed numbers; the acceptance level is 1e-3 per test.
"""
import os
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
P_MIN = 1e-3


@pytest.fixture(scope="module")
def capi():
    from upcgen_b200 import capi as c
    c.lib()
    return c


def _setup(capi, name, extra=""):
    """tables -> lumi fill -> fold -> samplers on the GPU; returns (P, gpu, lumi, cs)."""
    from upcgen_b200.config import named_config
    P = named_config(name, extra)
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    lumi = g.fill_lumi()
    cs, _, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
    g.sampler_build(cszm=None if P.ignore_csz else capi.elem_cs_zm(P))
    return P, g, lumi, cs


def _generate(g, n, seed=20261018, chunk=1 << 20):
    out = {"npart": [], "p4": [], "aux": [], "pdg": []}
    for first in range(0, n, chunk):
        ev = g.generate(seed, first, min(chunk, n - first))
        for k in out:
            out[k].append(ev[k])
    return {k: np.concatenate(v) for k, v in out.items()}


def _pt(p4):
    return np.hypot(p4[..., 0], p4[..., 1])


def _eta(p4):
    p = np.sqrt(p4[..., 0] ** 2 + p4[..., 1] ** 2 + p4[..., 2] ** 2)
    c = p4[..., 2] / p
    return -0.5 * np.log((1 - c) / (1 + c))


def test_chi2_of_sampled_bins_against_sigma_table(capi):
    """4e6 (y, m) draws: counts per (yBin, mBin) against N cs / sum(cs); bins with small expectation merged (in order
    of expectation) until every group expects >= 20 events."""
    from scipy import stats
    P, g, lumi, cs = _setup(capi, "cfg1", "PROC_ID 13\n")
    n = 4_000_000
    ev = _generate(g, n)
    y, m = ev["aux"][:, 0], ev["aux"][:, 1]
    iy = np.clip(np.floor((y - P.ymin) / P.dy).astype(np.int64), 0, P.ny - 1)
    im = np.clip(np.floor((m - P.mmin) / P.dm).astype(np.int64), 0, P.nm - 1)
    counts = np.bincount(iy * P.nm + im, minlength=P.ny * P.nm).astype(float)
    expect = (cs / cs.sum()).ravel() * n
    assert counts[expect == 0].sum() == 0          # nothing is drawn where the cross section vanishes
    order = np.argsort(expect, kind="stable")
    e_sorted, c_sorted = expect[order], counts[order]
    # merge ascending until each group holds >= 20 expected events
    csum = np.cumsum(e_sorted)
    gid = np.floor(csum / 20.0).astype(np.int64)
    gid = np.minimum(gid, gid[-1] - 1) if gid[-1] > 0 else gid   # the last partial group joins its neighbour
    ge = np.bincount(gid, weights=e_sorted)
    gc = np.bincount(gid, weights=c_sorted)
    keep = ge > 0
    chi2 = float(np.sum((gc[keep] - ge[keep]) ** 2 / ge[keep]))
    dof = int(keep.sum()) - 1
    p = float(stats.chi2.sf(chi2, dof))
    print(f"chi2 = {chi2:.1f} for {dof} dof (p = {p:.3g}); {keep.sum()} groups from {np.count_nonzero(expect)} bins")
    assert p > P_MIN
    # marginals as well (far more events per bin: tighter in absolute terms)
    for ax, nbin, idx in ((0, P.ny, iy), (1, P.nm, im)):
        e1 = cs.sum(axis=1 - ax) / cs.sum() * n
        c1 = np.bincount(idx, minlength=nbin).astype(float)
        sel = e1 >= 20
        chi = float(np.sum((c1[sel] - e1[sel]) ** 2 / e1[sel]))
        pp = float(stats.chi2.sf(chi, int(sel.sum()) - 1))
        print(f"marginal axis {ax}: chi2 = {chi:.1f} / {int(sel.sum()) - 1} (p = {pp:.3g})")
        assert pp > P_MIN
    # inside a bin the draw is uniform (delta in y, r2 in m): KS of the in-bin fractions
    fy = (y - (P.ymin + P.dy * iy)) / P.dy
    fm = (m - (P.mmin + P.dm * im)) / P.dm
    assert stats.kstest(fy[:200000], "uniform").pvalue > P_MIN
    assert stats.kstest(fm[:200000], "uniform").pvalue > P_MIN
    g.close()


def test_ks_of_cos_theta_inside_mass_bins(capi, oracle_mod):
    """cos(theta) of the events whose mass falls into one bin follows that bin's cumulative dsigma/dz table
    (samplersCsZ[mBin], src/UpcGenerator.cpp:753): KS against the piecewise-linear cdf built by the ORACLE from its own
    fillCrossSectionZM -- for bins whose getBinY index equals the true bin, and for the bins where the reference's
    integer arithmetic (include/UpcSampler.h:130-133) picks the lower neighbour."""
    from scipy import stats
    from oracle import pyoracle
    P, g, lumi, cs = _setup(capi, "cfg1", "PROC_ID 13\n")
    o = pyoracle.Oracle(P)
    ev = _generate(g, 2_000_000)
    m, z = ev["aux"][:, 1], ev["aux"][:, 2]
    ze = P.zmin + P.dz * np.arange(P.nz + 1)
    mb = np.array([pyoracle.get_bin(P.nm, float(x), P.mmin, P.mmax) for x in m[:400000]])  # the reference's index
    zz = z[:400000]
    ozm = o.cs_zm(0)
    checked = 0
    for b in (0, 1, 3, 10, 40):
        sel = zz[mb == b]
        if sel.size < 500:
            continue
        cdf = pyoracle.pdf_init(ozm[b])
        p = stats.kstest(sel, lambda x: np.interp(x, ze, cdf)).pvalue
        print(f"m bin {b}: {sel.size} events, KS p = {p:.3g}")
        assert p > P_MIN
        checked += 1
    assert checked >= 3
    g.close()


def test_ks_of_photon_pt_against_the_cumulative_pdf(capi):
    """Photon pT (aux[3], aux[4]) of the photons whose energy falls on one integer-MeV key against that key's
    5000-bin cumulative pdf (upcgpu_photon_pt_cdf at the key's centre), and -- two-sample -- against draws of the
    reference's own getPhotonPt at that energy."""
    from scipy import stats
    P, g, lumi, cs = _setup(capi, "cfg1", "PROC_ID 13\n")
    ev = _generate(g, 2_000_000)
    y, m = ev["aux"][:, 0], ev["aux"][:, 1]
    k1 = (m / 2 * np.exp(y) * 1e3).astype(np.int64)
    k2 = (m / 2 * np.exp(-y) * 1e3).astype(np.int64)
    keys = np.concatenate([k1, k2]); pts = np.concatenate([ev["aux"][:, 3], ev["aux"][:, 4]])
    uniq, cnt = np.unique(keys, return_counts=True)
    best = uniq[np.argsort(cnt)[-3:]]
    edges = 6 * 0.1973269718 / P.R / 5000 * np.arange(5001)
    for key in best:
        sel = pts[keys == key]
        cdf = g.photon_pt_cdf((key + 0.5) * 1e-3)
        p = stats.kstest(sel, lambda x: np.interp(x, edges, cdf)).pvalue
        print(f"key {key} MeV: {sel.size} photons, KS p = {p:.3g}")
        assert sel.size > 100 and p > P_MIN
    g.close()


def _reference_sample(P, cfg, extra, lumi, n):
    from oracle import pyref
    from upcgen_b200.config import config_text
    d = tempfile.mkdtemp()
    # BREAKUP_MODE only enters the luminosity table, which is injected: 1 skips the reference's minute-long table
    ref = pyref.RefGenerator(config_text(cfg, extra + "BREAKUP_MODE 1\n"), d, lumi=lumi,
                             grid=(P.nm, P.ny, P.mmin, P.mmax, P.ymin, P.ymax))
    # The reference seeds the 2-D sampler and all nm z samplers alike (src/UpcGenerator.cpp:688-700, Q6): the k-th draw
    # of every mass bin uses the same uniform, so its pooled cos(theta) sample is not i.i.d. (the first draw of each of
    # the 1001 bins lands on one quantile -- a KS test of the lepton pT sees that at p ~ 1e-18).  The streams are
    # re-seeded one by one; the sampling CODE stays the reference's.
    ref.reseed_z(777)
    return ref, ref.generate(n)


def _two_sample(name, a, b):
    from scipy import stats
    p = stats.ks_2samp(a, b).pvalue
    print(f"  {name}: KS two-sample p = {p:.3g}  ({a.size} GPU vs {b.size} reference)")
    assert p > P_MIN, name
    return p


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libupcref.so")),
                    reason="oracle/_ref not built")
def test_two_sample_vs_reference_generate_event_ditau(capi):
    """cfg1 (the repo's parameters.in: ditau, point flux, photon pT on) with kinematic cuts: GPU events against events
    drawn by the reference's generateEvent from the same luminosity table."""
    extra = "DO_PT_CUT 1\nPT_MIN 0.3\nDO_ETA_CUT 1\nETA_MIN -4\nETA_MAX 4\n"
    P, g, lumi, cs = _setup(capi, "cfg1", extra)
    n_gpu, n_ref = 1_000_000, 60_000
    ev = _generate(g, n_gpu)
    ref, rv = _reference_sample(P, "cfg1", extra, lumi, n_ref)
    # acceptance of the cuts: binomial compatibility
    a_gpu, a_ref = np.mean(ev["npart"] > 0), np.mean(rv["npart"] > 0)
    sig = np.sqrt(a_ref * (1 - a_ref) * (1 / n_gpu + 1 / n_ref))
    print(f"acceptance: GPU {a_gpu:.5f}, reference {a_ref:.5f}  ({abs(a_gpu - a_ref) / sig:.2f} sigma)")
    assert abs(a_gpu - a_ref) < 4 * sig
    ga, ra = ev["p4"][ev["npart"] > 0], rv["p4"][rv["npart"] > 0]
    pair_g, pair_r = ga[:, 0] + ga[:, 1], ra[:, 0] + ra[:, 1]
    mass = lambda q: np.sqrt(np.maximum(q[:, 3] ** 2 - q[:, 0] ** 2 - q[:, 1] ** 2 - q[:, 2] ** 2, 0))
    rap = lambda q: 0.5 * np.log((q[:, 3] + q[:, 2]) / (q[:, 3] - q[:, 2]))
    _two_sample("pair mass", mass(pair_g), mass(pair_r))
    _two_sample("pair rapidity", rap(pair_g), rap(pair_r))
    _two_sample("pair pT", _pt(pair_g), _pt(pair_r))
    _two_sample("lepton pT", _pt(ga[:, 0]), _pt(ra[:, 0]))
    _two_sample("lepton eta", _eta(ga[:, 0]), _eta(ra[:, 0]))
    _two_sample("second lepton eta", _eta(ga[:, 1]), _eta(ra[:, 1]))
    # charge assignment is a fair coin in both
    qg, qr = np.mean(ev["pdg"][ev["npart"] > 0][:, 0] > 0), np.mean(rv["pdg"][rv["npart"] > 0][:, 0] > 0)
    assert abs(qg - 0.5) < 4 * 0.5 / np.sqrt(ga.shape[0]) and abs(qr - 0.5) < 4 * 0.5 / np.sqrt(ra.shape[0])
    # photon pT at one energy: GPU table draws vs the reference's getPhotonPt (TH1::GetRandom on its own histogram)
    e0 = 2.0005
    cdf = g.photon_pt_cdf(e0)
    edges = 6 * 0.1973269718 / P.R / 5000 * np.arange(5001)
    u = np.random.default_rng(5).random(100000)
    mine = np.interp(u, cdf, edges)
    _two_sample("photon pT at 2 GeV", mine, ref.photon_pt(e0, 100000))
    g.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libupcref.so")),
                    reason="oracle/_ref not built")
def test_two_sample_vs_reference_generate_event_alp(capi):
    """cfg5 (Xe-Xe ALP, photon pT on, uniform two-photon decay): the ALP's pT and rapidity and the decay photons'
    pT / eta against the reference's generateEvent."""
    P, g, lumi, cs = _setup(capi, "cfg5")
    n_gpu, n_ref = 1_000_000, 60_000
    ev = _generate(g, n_gpu)
    ref, rv = _reference_sample(P, "cfg5", "", lumi, n_ref)
    assert np.all(ev["npart"] == 3) and np.all(rv["npart"] == 3)
    ga, ra = ev["p4"], rv["p4"]
    rap = lambda q: 0.5 * np.log((q[:, 3] + q[:, 2]) / (q[:, 3] - q[:, 2]))
    _two_sample("ALP pT", _pt(ga[:, 0]), _pt(ra[:, 0]))
    _two_sample("ALP rapidity", rap(ga[:, 0]), rap(ra[:, 0]))
    _two_sample("decay photon pT", _pt(ga[:, 1]), _pt(ra[:, 1]))
    _two_sample("decay photon eta", _eta(ga[:, 2]), _eta(ra[:, 2]))
    # momentum conservation of the decay in both samples
    assert np.max(np.abs(ga[:, 1] + ga[:, 2] - ga[:, 0])) < 1e-9 * np.max(np.abs(ga[:, 0]))
    g.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libupcref.so")),
                    reason="oracle/_ref not built")
def test_two_sample_vs_reference_generate_event_pi0_pairs(capi):
    """PROC_ID 111 (pi0 pairs, each decayed uniformly into two photons): GPU events against the reference's own
    UpcGenerator + src/UpcTwoPhotonDipion.cpp, both fed the SAME sigma(m) and dsigma/dz histograms -- placed into the
    shim's in-memory ROOT files for the reference, looked up with TAxis::FindBin arithmetic for the GPU side (the
    reference's own pi0 pi0 files do not travel to the GPU box; the plug-in's look-ups on them are pinned bit for bit by
    tests/test_reference_tabulated.py)."""
    from oracle import pyref
    from upcgen_b200.config import named_config, config_text
    extra = "PROC_ID 111\nBINS_Y 40\nDO_PT_CUT 1\nPT_MIN 0.1\n"
    P = named_config("cfg1", extra)
    assert (P.nm, P.mmin, P.mmax, P.nz) == (91, 0.275, 5.0, 100)
    # the two histograms of cross_sections/pi0pi0: sigma(m) on 100 bins of [0, 5], dsigma/dz on 100 x 100 bins of [-1, 1] x [0, 5]
    mc = 0.025 + 0.05 * np.arange(100)
    zc = -0.99 + 0.02 * np.arange(100)
    h1 = np.concatenate([[0.0], np.where(mc > 0.27, 25.0 / (0.3 + mc) ** 3, 0.0), [0.0]])
    h2 = np.zeros((102, 102))                                  # [m bin][z bin], z (x axis) fastest
    h2[1:-1, 1:-1] = np.outer(h1[1:-1], 1.0 + 0.8 * zc ** 2)
    d = pyref.cross_sec_dir()
    pyref.put_hist(f"{d}/pi0pi0/cross_section_m.root", "hCrossSectionM", h1, (100, 0.0, 5.0))
    pyref.put_hist(f"{d}/pi0pi0/cross_section_zm.root", "hCrossSectionZM", h2, (100, -1.0, 1.0), (100, 0.0, 5.0))
    findbin = lambda x, n, lo, hi: 0 if x < lo else (n + 1 if not x < hi else 1 + int(n * (x - lo) / (hi - lo)))
    m = P.mmin + P.dm * np.arange(P.nm)
    z = P.zmin + P.dz * np.arange(P.nz)
    sig = np.array([h1[findbin(x, 100, 0.0, 5.0)] for x in m])
    hc = 0.1973269718
    cszm = np.array([[h2[findbin(a, 100, 0.0, 5.0), findbin(b, 100, -1.0, 1.0)] * (hc * hc * 1e7) / P.dm for b in z] for a in m])
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    lumi = g.fill_lumi()
    cs, _, tot = g.fold_sigma(sig_m=sig)
    g.sampler_build(cszm=cszm)
    n_gpu, n_ref = 600_000, 50_000
    ev = _generate(g, n_gpu)
    ref = pyref.RefGenerator(config_text("cfg1", extra + "BREAKUP_MODE 1\n"), tempfile.mkdtemp(), lumi=lumi,
                             grid=(P.nm, P.ny, P.mmin, P.mmax, P.ymin, P.ymax))
    assert ref.totcs() == pytest.approx(tot, rel=1e-12)       # the reference folded the same table with the same sigma(m)
    ref.reseed_z(4242)
    rv = ref.generate(n_ref)
    a_gpu, a_ref = np.mean(ev["npart"] > 0), np.mean(rv["npart"] > 0)
    sg = np.sqrt(a_ref * (1 - a_ref) * (1 / n_gpu + 1 / n_ref))
    print(f"acceptance: GPU {a_gpu:.5f}, reference {a_ref:.5f}  ({abs(a_gpu - a_ref) / sg:.2f} sigma)")
    assert abs(a_gpu - a_ref) < 4 * sg
    ga, ra = ev["p4"][ev["npart"] > 0], rv["p4"][rv["npart"] > 0]
    assert np.all(ev["npart"][ev["npart"] > 0] == 6) and np.all(rv["npart"][rv["npart"] > 0] == 6)
    assert np.all(rv["pdg"][rv["npart"] > 0] == [111, 111, 22, 22, 22, 22]) and np.all(rv["mother"][rv["npart"] > 0] == [0, 0, 1, 1, 2, 2])
    mass = lambda q: np.sqrt(np.maximum(q[:, 3] ** 2 - q[:, 0] ** 2 - q[:, 1] ** 2 - q[:, 2] ** 2, 0))
    rap = lambda q: 0.5 * np.log((q[:, 3] + q[:, 2]) / (q[:, 3] - q[:, 2]))
    _two_sample("pair mass", mass(ga[:, 0] + ga[:, 1]), mass(ra[:, 0] + ra[:, 1]))
    _two_sample("pair rapidity", rap(ga[:, 0] + ga[:, 1]), rap(ra[:, 0] + ra[:, 1]))
    _two_sample("pi0 pT", _pt(ga[:, 0]), _pt(ra[:, 0]))
    _two_sample("pi0 eta", _eta(ga[:, 1]), _eta(ra[:, 1]))
    _two_sample("photon pT (first decay)", _pt(ga[:, 2]), _pt(ra[:, 2]))
    _two_sample("photon eta (second decay)", _eta(ga[:, 5]), _eta(ra[:, 5]))
    _two_sample("photon energy (second decay)", ga[:, 4, 3], ra[:, 4, 3])
    g.close()
