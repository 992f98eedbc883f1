"""The light-by-light (PROC_ID 22) and pi0 pi0 (111) plug-ins against the REFERENCE's own src/UpcTwoPhotonLbyL.cpp and
src/UpcTwoPhotonDipion.cpp, compiled unmodified into oracle/_ref.  Those read TH1D / TH2D objects from ROOT files; the
shim's TFile finds them in memory, where this test places what an independent Python parser (tests/test_root_hist.py)
read from the reference's own cross_sections/*.root.  The host plug-ins (upcgen_b200/host/UpcTwoPhotonTabulated.cpp over
the ROOT-less reader UpcRootHist.cpp) must give the same sigma(m) and dsigma/dz, bit for bit, on bin centres, on the
lower edges the generator's grid uses, and beyond both ends.  CPU only; runs where /root/reference is mounted."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/cross_sections"
from oracle import pyref  # noqa: E402

_CASE = r"""
import json, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import numpy as np
from oracle import pyref
from upcgen_b200.config import named_config
from test_root_hist import py_read_hist
proc, sub = {proc}, {sub!r}
d = pyref.cross_sec_dir()
c1 = py_read_hist(f"{ref}/{{sub}}/cross_section_m.root", "hCrossSectionM")
c2 = py_read_hist(f"{ref}/{{sub}}/cross_section_zm.root", "hCrossSectionZM")
pyref.put_hist(f"{{d}}/{{sub}}/cross_section_m.root", "hCrossSectionM", c1[3], c1[1])
pyref.put_hist(f"{{d}}/{{sub}}/cross_section_zm.root", "hCrossSectionZM", c2[3], c2[1], c2[2])
P = named_config("cfg1", "PROC_ID %d\nUSE_POLARIZED_CS 0\nBINS_Y 8\n" % proc)
ref = pyref.Reference(P)                       # UpcCrossSection::setElemProcess(proc) -> the reference's plug-in
m = np.array({m!r}); z = np.array({z!r})
sm = [ref.L.upcref_sigma_m(float(x)) for x in m]
szm = [[ref.L.upcref_sigma_zm(float(a), float(b)) for a in z] for b in m]
print("RESULT " + json.dumps(dict(sm=sm, szm=szm)))
"""


@pytest.mark.skipif(not (pyref.available() and os.path.isdir(REF)), reason="needs oracle/_ref and the reference's cross_sections")
@pytest.mark.parametrize("proc,sub,mlo,mhi,nm,zlo,zhi,nz", [(22, "lbyl", 0.05, 50.0, 1000, -0.99, 0.99, 198),
                                                            (111, "pi0pi0", 0.275, 5.0, 91, -1.0, 1.0, 100)])
def test_host_plugins_equal_the_references_own(proc, sub, mlo, mhi, nm, zlo, zhi, nz):
    import ctypes as C
    from upcgen_b200 import capi
    dm, dz = (mhi - mlo) / nm, (zhi - zlo) / nz
    rng = np.random.default_rng(proc)
    m = np.concatenate([mlo + dm * np.arange(0, nm, max(1, nm // 60)), rng.uniform(mlo, mhi, 40), [mlo - 0.01, mhi, mhi + 3.0, 0.0]])
    z = np.concatenate([zlo + dz * np.arange(0, nz, max(1, nz // 25)), rng.uniform(zlo, zhi, 15), [zlo - 0.5, zhi, 1.5]])
    r = subprocess.run([sys.executable, "-c", _CASE.format(root=ROOT, ref=REF, proc=proc, sub=sub, m=m.tolist(), z=z.tolist())],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    os.environ["UPCGEN_CROSS_SEC_DIR"] = REF
    L = capi.lib()
    out = np.zeros(m.size)
    assert L.upcgpu_elem_sigma_m(proc, 0., 0., 0., 0, m.ctypes.data, m.size, out.ctypes.data) == 0
    assert np.array_equal(out, np.array(ref["sm"]))                       # sigma(m): the same bin, the same content
    assert np.count_nonzero(out) > 40
    # dsigma/dz through the C-ABI's table filler on a grid whose lower edges are the probe points: one (m, z) per call
    L.upcgpu_elem_fill_cs_zm.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int,
                                         C.c_double, C.c_double, C.c_int, C.c_void_p]
    hc = 0.1973269718
    worst = 0
    for im in range(0, m.size, 7):
        for iz in range(0, z.size, 5):
            one = np.zeros(1)
            # fillCrossSectionZM evaluates at the lower edges: a 1 x 1 grid starting at (m, z); width 1 in both
            assert L.upcgpu_elem_fill_cs_zm(proc, 0., 0., 0., 0, float(z[iz]), float(z[iz]) + 1.0, 1, float(m[im]), float(m[im]) + 1.0,
                                            1, one.ctypes.data) == 0
            dm1 = ((float(m[im]) + 1.0) - float(m[im])) / 1                 # the grid step the call derives (not exactly 1)
            want = ref["szm"][im][iz] * (hc * hc * 1e7) / dm1             # src/UpcCrossSection.cpp:337-362
            assert one[0] == want, (m[im], z[iz], one[0], want)
            worst += one[0] != 0
    assert worst > 10
