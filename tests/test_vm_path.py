"""The vector-meson 1-D path: UpcCrossSection::calcPhotonFlux / calcNucCrossSectionY
(src/UpcCrossSection.cpp:700-748), getMomentumVM (:1076-1104), the UpcPhotoNuclearVM plug-in
(src/UpcPhotoNuclearVM.cpp) and the vector-meson branch of the event loop (src/UpcGenerator.cpp:425-472, :701-712).

  * the oracle's restatement of calcPhotonFlux == the REFERENCE's own method (oracle/_ref) bit for bit (CPU);
  * upcgpu_photon_flux == the oracle within the north-star tolerances (GPU);
  * the plug-in's pieces against independent evaluations with scipy (CPU): the power-law dsigma/dt, the integrated
    squared form factor (scipy.integrate.quad of the oracle's form factor), the LTA shadowing spline
    (scipy CubicSpline, not-a-knot = ROOT's TSpline3 default) and sigma(y) assembled from them;
  * `upcgen` with PROC_ID 443: cross section == flux x sigma summed over y, events on the J/psi mass shell (GPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOST = os.path.join(ROOT, "upcgen_b200", "host")
REF = "/root/reference/cross_sections"
from oracle import pyref  # noqa: E402

HC, M_PROT, M_NEUT = 0.1973269718, 0.9382720813, 0.939565346

_FLUX_CASE = r"""
import json, sys
sys.path.insert(0, {root!r})
import numpy as np
from oracle import pyoracle, pyref
from upcgen_b200.config import named_config
P = named_config("cfg1", {extra!r})
ref = pyref.Reference(P)
o = pyoracle.Oracle(P)
ok = True; worst = 0.0
for M in (3.0969, 9.3987):
    for Y in (-5.5, -3.0, -0.7, 0.0, 1.9, 4.2, 5.9):
        a, b = ref.L.upcref_photon_flux(M, Y), o.photon_flux(M, Y)
        ok &= (a == b)
        worst = max(worst, abs(a - b) / max(abs(a), 1e-300))
print("RESULT " + json.dumps(dict(equal=bool(ok), worst=worst)))
"""


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("extra", ["FLUX_POINT 1\n", "FLUX_POINT 0\n"])
def test_oracle_photon_flux_equals_reference(extra):
    r = subprocess.run([sys.executable, "-c", _FLUX_CASE.format(root=ROOT, extra=extra)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert out["equal"], out


def _vm_check(pdg, shad, dght, P, rho0, env=None):
    subprocess.check_call(["make", "-s", "-C", HOST])
    exe = os.path.join(ROOT, "tests", "cpp", "vm_check")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I", HOST, "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "vm_check.cpp"), "-L", HOST, "-lupcgen_host",
                           "-L", os.path.join(ROOT, "upcgen_b200"), "-lupcgpu", f"-Wl,-rpath,{HOST}",
                           f"-Wl,-rpath,{os.path.join(ROOT, 'upcgen_b200')}"])
    r = subprocess.run([exe, str(pdg), str(shad), str(dght), str(P.Z), str(P.A), repr(P.R), repr(P.a), repr(P.sqrts),
                        repr(rho0)], capture_output=True, text=True, timeout=120, env={**os.environ, **(env or {})})
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def _py_sigma(P, o, pdg, y, rg=None):
    """sigma(y) of src/UpcPhotoNuclearVM.cpp:340-381 from independent pieces."""
    from scipy import integrate
    mPart, c0, pw = {443: (3.0969, 342., 0.4), 100443: (3.6861, 56.8, 0.4), 553: (9.3987, 0.902, 0.447)}[pdg]
    mNucl = (P.Z * M_PROT + (P.A - P.Z) * M_NEUT) / P.A
    w = mPart / 2 * np.exp(y)
    Wgp2 = 4 * w * 0.5 * P.sqrts
    mmin = M_PROT + mPart
    ds = c0 * (1 - mmin ** 2 / Wgp2) ** 1.5 * (Wgp2 * 1e-4) ** pw if np.sqrt(Wgp2) > mmin else 0.0
    x = mPart ** 2 / Wgp2
    tmin = x * x * mNucl ** 2
    phi, _ = integrate.quad(lambda t: o.formfac([t])[0] ** 2, tmin, tmin + 1, epsabs=0, epsrel=1e-11, limit=400,
                            points=None)
    c2, r = (1.0, 1.0) if rg is None else (0.81, rg(x))
    return c2 * ds * r * r * phi * 1e-6


def test_vm_plugin_impulse_approximation(get_oracle):
    P, o = get_oracle("cfg1")
    d = _vm_check(443, 0, 13, P, o.rho0())
    assert d["mPart"] == 3.0969 and d["mDght"] == 0.1056583745
    ys = -6 + 0.5 * np.arange(25)
    mine = np.array(d["sigma"])
    ref = np.array([_py_sigma(P, o, 443, y) for y in ys])
    sel = ref > 0
    assert np.array_equal(mine == 0, ref == 0)
    assert np.max(np.abs(mine[sel] / ref[sel] - 1)) < 1e-8
    # the power law alone
    for W, v in zip((3.0, 10., 100.), d["dsdt"]):
        mmin = M_PROT + 3.0969
        exp = 342. * (1 - mmin ** 2 / W ** 2) ** 1.5 * (W * W * 1e-4) ** 0.4 if W > mmin else 0.
        assert v == pytest.approx(exp, rel=1e-14)
    d2 = _vm_check(553, 0, 11, P, o.rho0())
    assert d2["mPart"] == 9.3987 and np.array(d2["sigma"])[12] == pytest.approx(_py_sigma(P, o, 553, 0.0), rel=1e-8)
    assert "error" in _vm_check(443, 1, 13, P, o.rho0())      # EPS09 grids are not shipped by the reference either


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
def test_vm_plugin_lta_shadowing(get_oracle):
    """SHADOWING 4: the Guzey-Zhalov table through a not-a-knot cubic spline inside (1e-5, 1e-1), linear outside."""
    from scipy.interpolate import CubicSpline
    P, o = get_oracle("cfg1")
    d = _vm_check(443, 4, 13, P, o.rho0(), env={"UPCGEN_CROSS_SEC_DIR": REF})
    tab = np.loadtxt(os.path.join(REF, "vm", "lta", "LT2013_pb208_cteq6l1_m12_Q2_3.dat"))[:37]
    xs_t, rg_t = tab[:, 0], tab[:, 2]
    cs = CubicSpline(xs_t, rg_t, bc_type="not-a-knot")

    def rg(x):
        if 1e-5 < x < 1e-1:
            return float(cs(x))
        if x <= xs_t[0]:
            lo, up = 0, 1
        elif x >= xs_t[-1]:
            lo, up = len(xs_t) - 2, len(xs_t) - 1
        else:
            lo = int(np.searchsorted(xs_t, x, side="right") - 1); up = lo + 1
        return rg_t[up] + (x - xs_t[up]) * (rg_t[lo] - rg_t[up]) / (xs_t[lo] - xs_t[up])

    for x, v in zip((1e-6, 5e-6, 2e-5, 1e-4, 1.3e-3, 2e-2, 0.09, 0.5), d["rg"]):
        assert v == pytest.approx(rg(x), rel=1e-10), x
    ys = -6 + 0.5 * np.arange(25)
    ref = np.array([_py_sigma(P, o, 443, y, rg) for y in ys])
    mine = np.array(d["sigma"])
    sel = ref > 0
    assert np.max(np.abs(mine[sel] / ref[sel] - 1)) < 1e-8


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
@pytest.mark.parametrize("shad,model", [(2, 2), (3, 1)])
def test_vm_plugin_fgs10_nodes(shad, model, get_oracle):
    """SHADOWING 2 / 3 for psi(2S) (mu^2 = 4 = the grid's first Q^2 row): a bicubic patch returns the table's glue
    column at the table's own x nodes, clamps x to [9.99999975e-6, 0.95] (src/UpcPhotoNuclearVM.cpp:291-294) and is
    smooth between nodes; J/psi (mu^2 = 3) keeps the impulse approximation (:366-374)."""
    P, o = get_oracle("cfg1")
    d = _vm_check(100443, shad, 13, P, o.rho0(), env={"UPCGEN_CROSS_SEC_DIR": REF})
    assert "error" not in d, d
    tok = open(os.path.join(REF, "vm", "lta", f"QCDEvolution_pb208proton_2009_model{model}.dat")).read().split()
    assert float(tok[0]) == 4.0
    rows = np.array(tok[1:1 + 90 * 9], dtype=float).reshape(90, 9)
    xs_t, glue = rows[:, 0], rows[:, 7]
    rg = np.array(d["rg"])
    # vm_check's probes: below the grid (clamped to the first node), the first node, the node 3e-5 (9 printed digits:
    # 2.99999992E-05, so the probe sits 8e-13 off it), between nodes, above 0.95 (clamped)
    assert rg[0] == rg[1] == pytest.approx(glue[0], rel=1e-14)
    assert rg[2] == pytest.approx(glue[4], rel=1e-6)
    for x, v in zip((1.2e-4, 1.3e-3, 2e-2, 0.4), rg[3:7]):
        k = np.searchsorted(xs_t, x)
        lo, hi = sorted((glue[k - 1], glue[k]))
        pad = 0.05 * (hi - lo) + 1e-3
        assert lo - pad <= v <= hi + pad, (x, v, lo, hi)
    k95 = np.searchsorted(xs_t, 0.95)
    assert min(glue[k95 - 1:k95 + 1]) - 0.05 <= rg[7] <= max(glue[k95 - 1:k95 + 1]) + 0.05
    # J/psi: the option leaves the impulse approximation in place
    a = _vm_check(443, shad, 13, P, o.rho0(), env={"UPCGEN_CROSS_SEC_DIR": REF})
    b = _vm_check(443, 0, 13, P, o.rho0())
    assert a["sigma"] == b["sigma"]


_VM_REF_CASE = r"""
import json, sys
sys.path.insert(0, {root!r})
import numpy as np
from oracle import pyref
from upcgen_b200.config import named_config
P = named_config("cfg1")
ref = pyref.Reference(P)
ys = -6 + 0.5 * np.arange(25)
out = {{}}
for pdg, shad, dght in {cases!r}:
    sig, mpart = ref.vm_sigma_y(pdg, shad, dght, ys)
    out[f"{{pdg}}_{{shad}}_{{dght}}"] = dict(sigma=sig.tolist(), mPart=mpart)
print("RESULT " + json.dumps(out))
"""


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("cases", [[(443, 4, 13), (553, 0, 11)], [(100443, 4, 13), (443, 0, 11)],
                                   [(100443, 2, 13), (443, 2, 13)], [(553, 2, 11)], [(553, 3, 11)], [(100443, 3, 13)]])
def test_vm_plugin_equals_the_references_own(cases, get_oracle):
    """The host plug-in against the REFERENCE's src/UpcPhotoNuclearVM.cpp, compiled unmodified into oracle/_ref (TF1 /
    TGraph / TSpline3 behind it are shim restatements: QAGS at 1e-12, not-a-knot spline, linear TGraph::Eval): sigma(y)
    on 25 rapidities for the impulse approximation, the LTA shadowing (SHADOWING 4) and the FGS10 grids (SHADOWING 2 / 3:
    gsl_spline2d's bicubic behind the reference is the shim's power-basis restatement of GSL's bicubic.c, the plug-in
    evaluates the same patch in Hermite form; one FGS10 instance per case list -- the reference keeps "initialised" in a
    function-level static but the spline in the instance, so a second instance of a process dereferences a null
    spline), J/psi, psi(2S), Upsilon.  The
    two sides integrate the squared form factor with different adaptive rules: agreement is 3e-15, the bar 1e-12."""
    r = subprocess.run([sys.executable, "-c", _VM_REF_CASE.format(root=ROOT, cases=cases)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    P, o = get_oracle("cfg1")
    for pdg, shad, dght in cases:
        d = _vm_check(pdg, shad, dght, P, o.rho0(), env={"UPCGEN_CROSS_SEC_DIR": REF})
        want = ref[f"{pdg}_{shad}_{dght}"]
        assert d["mPart"] == want["mPart"]
        mine, theirs = np.array(d["sigma"]), np.array(want["sigma"])
        assert np.array_equal(mine == 0, theirs == 0), (pdg, shad)
        sel = theirs > 0
        assert sel.sum() > 10
        e = np.max(np.abs(mine[sel] / theirs[sel] - 1))
        print(pdg, shad, "max rel", e)
        assert e < 1e-12, (pdg, shad, e)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("extra,tol", [("FLUX_POINT 1\nBREAKUP_MODE 1\n", 1e-9), ("FLUX_POINT 0\nBREAKUP_MODE 1\n", 1e-7),
                                       ("FLUX_POINT 0\nBREAKUP_MODE 2\n", 1e-7), ("FLUX_POINT 1\nBREAKUP_MODE 4\n", 1e-9)])
def test_gpu_photon_flux_vs_oracle(extra, tol, get_oracle):
    from upcgen_b200 import capi
    P, o = get_oracle("cfg1", extra)
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    Y = np.array([-5.5, -3.0, -0.7, 0.0, 1.9, 4.2, 5.9])
    for M in (3.0969, 9.3987):
        fp, fn = g.photon_flux(np.full(Y.size, M), Y)
        ref_p = np.array([o.photon_flux(M, y) for y in Y])
        ref_n = np.array([o.photon_flux(M, -y) for y in Y])
        assert np.max(np.abs(fp / ref_p - 1)) < tol and np.max(np.abs(fn / ref_n - 1)) < tol
    g.close()


@pytest.mark.gpu
def test_upcgen_cli_jpsi(tmp_path, get_oracle):
    """PROC_ID 443 end to end through the C++ drop-in: the cross section is sum_y (flux(+y) sigma(+y) + flux(-y)
    sigma(-y)) dy with the oracle's flux and the independent sigma(y); events are a J/psi on its mass shell (status 23)
    and two muons (status 33, mother 1) whose invariant mass is the J/psi's."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    par = """NUCLEUS_Z 82
NUCLEUS_A 208
WS_R 6.68
WS_A 0.447
SQRTS 5020
PROC_ID 443
SHADOWING 0
DECAY_PDG 13
NEVENTS 3000
YMIN -4
YMAX 4
BINS_Y 16
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
SEED 11
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "vm.in").write_text(par)
    r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "vm.in"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    from upcgen_b200.config import UpcParams
    P = UpcParams.from_text(par)
    from oracle import pyoracle
    o = pyoracle.Oracle(P)
    tot = 0.0
    for iy in range(P.ny):
        y = P.ymin + P.dy * iy
        tot += o.photon_flux(3.0969, y) * _py_sigma(P, o, 443, y) + o.photon_flux(3.0969, -y) * _py_sigma(P, o, 443, -y)
    tot *= P.dy
    line = [l for l in r.stdout.splitlines() if "total cross section" in l][0]
    assert float(line.split()[4]) == pytest.approx(tot, rel=1e-5)
    lines = (tmp_path / "events.hepmc").read_text().splitlines()
    ev = [l for l in lines if l.startswith("E ")]
    assert len(ev) == 3000 and ev[0] == "E 0 1 3"        # three particles, one decay vertex
    parts = [l.split() for l in lines if l.startswith("P ")]
    p = np.array([[float(x) for x in q[4:9]] for q in parts]).reshape(3000, 3, 5)
    pdg = np.array([int(q[3]) for q in parts]).reshape(3000, 3)
    st = np.array([int(q[9]) for q in parts]).reshape(3000, 3)
    mo = np.array([int(q[2]) for q in parts]).reshape(3000, 3)
    assert np.all(pdg[:, 0] == 443) and np.all(np.abs(pdg[:, 1:]) == 13) and np.all(pdg[:, 1] == -pdg[:, 2])
    assert np.all(st == [23, 33, 33]) and np.all(mo == [0, 1, 1])
    m = lambda q: np.sqrt(np.maximum(q[:, 3] ** 2 - q[:, 0] ** 2 - q[:, 1] ** 2 - q[:, 2] ** 2, 0))
    assert np.allclose(m(p[:, 0]), 3.0969, atol=2e-5)                      # mass bin is 2e-6 wide
    assert np.allclose(m(p[:, 1] + p[:, 2]), m(p[:, 0]), rtol=3e-5)         # the muons carry the J/psi (9 printed digits)
    assert np.allclose((p[:, 1] + p[:, 2])[:, :4], p[:, 0, :4], rtol=1e-7, atol=1e-7)
    assert np.allclose(p[:, 1, 4], 0.1056583745, atol=2e-6)
    rap = 0.5 * np.log((p[:, 0, 3] + p[:, 0, 2]) / (p[:, 0, 3] - p[:, 0, 2]))
    assert rap.min() >= -4 - 1e-9 and rap.max() <= 4 + 1e-9 and abs(rap.mean()) < 0.3
    assert np.all(np.hypot(p[:, 0, 0], p[:, 0, 1]) < 1.5)                   # photon + pomeron pT: a few hundred MeV at most


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flux_point", [1, 0])
def test_upcgen_cli_jpsi_vs_the_references_generator(tmp_path, flux_point):
    """PROC_ID 443 through the C++ drop-in (photon flux on the GPU) against the REFERENCE's own UpcGenerator run on the
    CPU in the same configuration (src/UpcGenerator.cpp: init, computeNuclXsection's VM branch, generateEvent with
    getMomentumVM and twoPartDecayVM; src/UpcPhotoNuclearVM.cpp): the total cross section agrees to the printed digits
    and 8000 events of each are compatible in the J/psi's rapidity and pT and the muons' pT and eta (two-sample KS)."""
    from scipy import stats
    subprocess.check_call(["make", "-s", "-C", HOST])
    n = 8000
    par = f"""NUCLEUS_Z 82
NUCLEUS_A 208
WS_R 6.68
WS_A 0.447
SQRTS 5020
PROC_ID 443
SHADOWING 0
DECAY_PDG 13
NEVENTS {n}
YMIN -4
YMAX 4
BINS_Y 40
FLUX_POINT {flux_point}
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
SEED 31
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""
    (tmp_path / "vm.in").write_text(par)
    r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "vm.in"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    mine_tot = float([l for l in r.stdout.splitlines() if "total cross section" in l][0].split()[4])
    parts = [l.split() for l in (tmp_path / "events.hepmc").read_text().splitlines() if l.startswith("P ")]
    mine = np.array([[float(x) for x in q[4:8]] for q in parts]).reshape(n, 3, 4)

    mp = 3.0969
    ref = pyref.RefGenerator(par, str(tmp_path / "ref"), lumi=np.zeros((1, 40)), grid=(1, 40, mp - 1e-6, mp + 1e-6, -4., 4.))
    assert mine_tot == pytest.approx(ref.totcs(), rel=1e-5)      # six printed digits
    ev = ref.generate(n)
    assert ev["n_accepted"] == n
    theirs = ev["p4"][:, :3, :]

    def obs(p):
        jp, mu = p[:, 0], p[:, 1:3].reshape(-1, 4)
        rap = 0.5 * np.log((jp[:, 3] + jp[:, 2]) / (jp[:, 3] - jp[:, 2]))
        pt_mu = np.hypot(mu[:, 0], mu[:, 1])
        eta = np.arcsinh(mu[:, 2] / pt_mu)
        return {"y": rap, "pt": np.hypot(jp[:, 0], jp[:, 1]), "pt_mu": pt_mu, "eta_mu": eta}
    a, b = obs(mine), obs(theirs)
    for k in a:
        pv = stats.ks_2samp(a[k], b[k]).pvalue
        print("FLUX_POINT", flux_point, k, "KS p =", pv)
        assert pv > 1e-3, (k, pv)
