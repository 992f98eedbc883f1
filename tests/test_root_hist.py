"""The ROOT-less TH1D/TH2D reader (upcgen_b200/host/UpcRootHist.cpp) and the file-based elementary processes
(UpcTwoPhotonLbyL / UpcTwoPhotonDipion, reference src/UpcTwoPhotonLbyL.cpp:32-80), through the C-ABI entry points
upcgpu_elem_sigma_m / upcgpu_elem_fill_cs_zm.

The data files are the reference's own (cross_sections/lbyl, cross_sections/pi0pi0); they are not copied into this
repository, so these tests run where /root/reference is mounted and skip elsewhere.  ROOT is not available to
cross-check against: what pins the reader is an independent parse of the same files in Python (below: TFile header,
TKey chain, zlib blocks, byte-counted TH1/TH2 streaming -- written separately from the C++ one), the exact
consumption of every byte count, cells == (nx + 2)(ny + 2), and the axes being the grids the reference hard-codes for
these processes (src/UpcGenerator.cpp:69-103).
"""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

REF = "/root/reference/cross_sections"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lookup_semantics_cpp(tmp_path):
    """tests/cpp/roothist_check.cpp: FindBin on uniform and variable-width axes, edges, under-/overflow, clamping."""
    import subprocess
    exe = tmp_path / "roothist_check"
    host = os.path.join(ROOT_DIR, "upcgen_b200", "host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", host, "-o", str(exe),
                           os.path.join(ROOT_DIR, "tests", "cpp", "roothist_check.cpp"),
                           os.path.join(host, "UpcRootHist.cpp"), os.path.join(host, "UpcLz4.cpp"), "-lz"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ROOTHIST_OK" in r.stdout, r.stdout + r.stderr


def py_read_hist(path, name):
    f = open(path, "rb").read()
    assert f[:4] == b"root"
    begin, end = struct.unpack(">ii", f[8:16])
    pos, best = begin, None
    while pos < end:
        nbytes, = struct.unpack(">i", f[pos:pos + 4])
        if nbytes < 0:
            pos -= nbytes
            continue
        if nbytes == 0:
            break
        version, objlen, _, keylen, cycle = struct.unpack(">hiIhh", f[pos + 4:pos + 18])
        p = pos + 18 + (16 if version > 1000 else 8)
        strs = []
        for _ in range(2):
            n = f[p]; p += 1
            strs.append(f[p:p + n].decode("latin1")); p += n
        if strs[1] == name and strs[0] in ("TH1D", "TH2D") and (best is None or cycle > best[0]):
            best = (cycle, strs[0], pos, nbytes, objlen, keylen)
        pos += nbytes
    _, cl, pos, nbytes, objlen, keylen = best
    raw = f[pos + keylen:pos + nbytes]
    buf, q = b"", 0
    while len(buf) < objlen:
        assert raw[q:q + 2] == b"ZL"
        csz = raw[q + 3] | raw[q + 4] << 8 | raw[q + 5] << 16
        buf += zlib.decompress(raw[q + 9:q + 9 + csz])
        q += 9 + csz
    P = [0]

    def rd(fmt):
        v, = struct.unpack(">" + fmt, buf[P[0]:P[0] + struct.calcsize(fmt)])
        P[0] += struct.calcsize(fmt)
        return v

    def obj():
        v = rd("I")
        assert v & 0x40000000
        e = P[0] + (v & 0x3FFFFFFF)
        rd("H")
        return e

    def skip():
        P[0] = obj()

    def axis():
        e = obj(); skip(); skip()
        nb, lo, hi, n = rd("i"), rd("d"), rd("d"), rd("i")
        assert n == 0
        P[0] = e
        return nb, lo, hi

    def th1():
        e = obj(); skip(); skip(); skip(); skip()
        rd("i")
        ax, ay, _ = axis(), axis(), axis()
        P[0] = e
        return ax, ay

    e0 = obj()
    if cl == "TH2D":
        e2 = obj(); ax, ay = th1(); P[0] = e2
    else:
        ax, ay = th1()
    n = rd("i")
    cells = np.frombuffer(buf, dtype=">f8", count=n, offset=P[0]).astype(np.float64)
    assert P[0] + 8 * n == e0 == len(buf)
    return cl, ax, ay, cells


@pytest.fixture(scope="module")
def lib():
    from upcgen_b200 import capi
    os.environ["UPCGEN_CROSS_SEC_DIR"] = REF
    L = C.CDLL(capi.SO_PATH)
    L.upcgpu_elem_sigma_m.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    L.upcgpu_elem_fill_cs_zm.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]
    return L


def sigma_m(L, proc, m, which=0):
    m = np.ascontiguousarray(m, dtype=np.float64)
    out = np.zeros_like(m)
    rc = L.upcgpu_elem_sigma_m(proc, 0., 0., 0., which, m.ctypes.data, m.size, out.ctypes.data)
    assert rc == 0
    return out


@needs_ref
@pytest.mark.parametrize("proc,sub,nz,zlo,zhi,nm,mlo,mhi", [
    (22, "lbyl", 198, -0.99, 0.99, 1000, 0.05, 50.0),      # the grid src/UpcGenerator.cpp:74-79 forces
    (111, "pi0pi0", 100, -1.0, 1.0, 100, 0.0, 5.0),
])
def test_histograms_and_lookups(lib, proc, sub, nz, zlo, zhi, nm, mlo, mhi):
    cl1, ax1, _, c1 = py_read_hist(f"{REF}/{sub}/cross_section_m.root", "hCrossSectionM")
    cl2, ax2, ay2, c2 = py_read_hist(f"{REF}/{sub}/cross_section_zm.root", "hCrossSectionZM")
    assert (cl1, cl2) == ("TH1D", "TH2D")
    assert ax1 == (nm, mlo, mhi) and ax2 == (nz, zlo, zhi) and ay2 == (nm, mlo, mhi)
    assert c1.size == nm + 2 and c2.size == (nz + 2) * (nm + 2)
    # sigma(m): bin centres, lower edges (a lower edge belongs to its bin), under- and overflow
    dm = (mhi - mlo) / nm
    centres = mlo + dm * (np.arange(nm) + 0.5)
    assert np.array_equal(sigma_m(lib, proc, centres), c1[1:nm + 1])
    edges = mlo + dm * np.arange(nm)
    got = sigma_m(lib, proc, edges)
    want = c1[[0 if x < mlo else (nm + 1 if not x < mhi else 1 + int(nm * (x - mlo) / (mhi - mlo))) for x in edges]]
    assert np.array_equal(got, want)                       # TAxis::FindBin arithmetic, incl. edges that round down a bin
    assert np.array_equal(sigma_m(lib, proc, [mlo - 1e-3, mhi, mhi + 7.]), c1[[0, nm + 1, nm + 1]])
    # polarised parts do not exist
    assert not sigma_m(lib, proc, centres[:5], 1).any() and not sigma_m(lib, proc, centres[:5], 2).any()
    # dsigma/dz table on the process's own grid: fillCrossSectionZM = content * (hc)^2 * 1e7 / dm at the LOWER edges
    out = np.zeros((nm, nz))
    gm_lo, gm_hi = (0.05, 50.0) if proc == 22 else (0.275, 5.0)
    g_nm = 1000 if proc == 22 else 91
    out = np.zeros((g_nm, nz))
    rc = lib.upcgpu_elem_fill_cs_zm(proc, 0., 0., 0., 0, zlo, zhi, nz, gm_lo, gm_hi, g_nm, out.ctypes.data)
    assert rc == 0
    h2 = c2.reshape(nm + 2, nz + 2)                        # [m bin][z bin], x (= z) fastest
    gdm, dz = (gm_hi - gm_lo) / g_nm, (zhi - zlo) / nz
    hc = 0.1973269718
    for im in (0, 1, g_nm // 2, g_nm - 1):
        for iz in (0, 1, nz // 2, nz - 1):
            m, z = gm_lo + gdm * im, zlo + dz * iz
            bz = 0 if z < zlo else (nz + 1 if not z < zhi else 1 + int(nz * (z - zlo) / (zhi - zlo)))
            bm = 0 if m < mlo else (nm + 1 if not m < mhi else 1 + int(nm * (m - mlo) / (mhi - mlo)))
            assert out[im, iz] == h2[bm, bz] * (hc * hc * 1e7) / gdm
    # consistency of the two files (a cell-order or axis mix-up in the reader would break it): the z integral of the
    # (z, m) histogram is proportional to sigma(m), with one constant for all mass bins
    integ = h2[1:nm + 1, 1:nz + 1].sum(axis=1) * dz
    sel = c1[1:nm + 1] > 0
    assert sel.sum() > nm // 2
    r = integ[sel] / c1[1:nm + 1][sel]
    print(proc, "z integral / sigma(m): mean", r.mean(), "relative spread", r.std() / r.mean())
    assert r.std() / r.mean() < 1e-3


@needs_ref
def test_missing_directory_is_an_error(lib):
    os.environ["UPCGEN_CROSS_SEC_DIR"] = "/nonexistent"
    try:
        m = np.array([1.0]); out = np.zeros(1)
        assert lib.upcgpu_elem_sigma_m(22, 0., 0., 0., 0, m.ctypes.data, 1, out.ctypes.data) != 0
    finally:
        os.environ["UPCGEN_CROSS_SEC_DIR"] = REF


@needs_ref
def test_c_abi_reader_equals_python_parse():
    """upcgpu_root_hist_read (the entry point a luminosity-cache comparison uses, tools/compare_lumi_root.py) against the
    independent Python parse: axes and every cell, under- and overflow included."""
    from upcgen_b200 import capi
    for sub, nm_ in (("lbyl", "hCrossSectionZM"), ("pi0pi0", "hCrossSectionZM")):
        path = f"{REF}/{sub}/cross_section_zm.root"
        cl, ax, ay, cells = py_read_hist(path, nm_)
        h = capi.root_hist_read(path, nm_)
        assert h["dim"] == 2 and (h["nx"], h["xlo"], h["xhi"]) == ax and (h["ny"], h["ylo"], h["yhi"]) == ay
        assert np.array_equal(h["cells"].ravel(), cells)
    h1 = capi.root_hist_read(f"{REF}/lbyl/cross_section_m.root", "hCrossSectionM")
    assert h1["dim"] == 1 and h1["cells"].shape == (1002,)
    with pytest.raises(capi.UpcGpuError):
        capi.root_hist_read(f"{REF}/lbyl/cross_section_m.root", "noSuchObject")


def test_reader_survives_corrupted_files(tmp_path):
    """These files are external input (a luminosity cache left by another job, cross sections from
    UPCGEN_CROSS_SEC_DIR): 600 truncated, byte-flipped and word-overwritten copies of an uncompressed, a zlib and an LZ4
    file go through the reader built with AddressSanitizer and UBSan -- every one must end in a value or an error
    message, never in a memory fault."""
    import subprocess
    from upcgen_b200 import capi
    host = os.path.join(ROOT_DIR, "upcgen_b200", "host")
    exe = str(tmp_path / "roothist_fuzz")
    r = subprocess.run(["/usr/bin/g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                        "-I", host, "-o", exe, os.path.join(ROOT_DIR, "tests", "cpp", "roothist_fuzz.cpp"),
                        os.path.join(host, "UpcRootHist.cpp"), os.path.join(host, "UpcLz4.cpp"), "-lz"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain: " + r.stderr[-200:])
    rng = np.random.default_rng(99)
    table = np.outer(np.linspace(1, 2, 60), np.ones(20))
    files = []
    try:
        for comp in (0, 101, 409):
            capi.root_set_compression(comp)
            p = str(tmp_path / f"h{comp}.root")
            capi.root_write_th2d(p, {"hD2LDMDY": table}, 60, 3.56, 50.0, 20, -6.0, 6.0)
            files.append(open(p, "rb").read())
    finally:
        capi.root_set_compression(0)
    names = [str(tmp_path / f"h{c}.root") for c in (0, 101, 409)]
    for it in range(600):
        f = bytearray(files[it % 3])
        mode = rng.integers(0, 3)
        if mode == 0:
            for _ in range(rng.integers(1, 6)):
                f[rng.integers(0, len(f))] = rng.integers(0, 256)
        elif mode == 1:
            f = f[:rng.integers(0, len(f))]
        else:
            i = rng.integers(0, len(f) - 4)
            f[i:i + 4] = bytes(rng.integers(0, 256, 4).astype(np.uint8))
        names.append(str(tmp_path / f"x{it}.root"))
        open(names[-1], "wb").write(bytes(f))
    r = subprocess.run([exe] + names, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    ok, bad = [int(x) for x in r.stdout.split("ok")[1].replace("bad", "").split()]
    assert ok >= 3 and bad > 100 and ok + bad == 603
