"""The ROOT-less WRITER (upcgen_b200/host/UpcRootFile.cpp, through upcgpu_root_write_th2d / upcgpu_root_write_tree):
the luminosity cache twoPhotonLumi[Pol].root and events.root as the reference writes them
(src/UpcCrossSection.cpp:493-507, :578-585; src/UpcGenerator.cpp:842-857).  No GPU involved.

ROOT is not available here to read the files back.  What pins the writer instead:
  * the TH2D object it streams is BYTE-IDENTICAL to the one ROOT 6.22/09 wrote into the reference's own
    cross_sections/lbyl/cross_section_zm.root when given the same name, axes and cells (runs where /root/reference is
    mounted): class versions, member order, default attributes, byte counts -- everything ROOT's own streamer put there;
  * the file-level records (header, directory, key list, streamer-info list, free segments) are walked by an
    independent parser written here from the format description, every byte accounted for;
  * the TTree is read back by an independent Python parser (below) that follows ROOT's streaming rules -- byte counts,
    class tags and back references, the branches' basket tables, the basket keys -- and returns the columns.
"""
import os
import struct
import sys
import zlib

import numpy as np
import pytest

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT_DIR)
REF = "/root/reference/cross_sections"


class Buf:
    def __init__(self, b, base=0):
        self.b, self.p, self.base = b, 0, base   # base: offset of b[0] inside its key buffer (fKeylen)

    def rd(self, fmt):
        n = struct.calcsize(">" + fmt)
        v = struct.unpack(">" + fmt, self.b[self.p:self.p + n])
        self.p += n
        return v[0] if len(v) == 1 else v

    def tstring(self):
        n = self.rd("B")
        if n == 255:
            n = self.rd("I")
        s = self.b[self.p:self.p + n].decode("latin1")
        self.p += n
        return s

    def cstring(self):
        e = self.b.index(b"\0", self.p)
        s = self.b[self.p:e].decode("latin1")
        self.p = e + 1
        return s

    def obj(self):
        v = self.rd("I")
        assert v & 0x40000000, hex(v)
        end = self.p + (v & 0x3FFFFFFF)
        return end, self.rd("H")


def read_keys(f):
    """TFile header + the chain of keys; returns (header dict, [key dict])."""
    assert f[:4] == b"root"
    ver, begin, end, seekfree, nbfree, nfree, nbname, units, compress, seekinfo, nbinfo = struct.unpack(">iiiiiiibiii", f[4:45])
    hdr = dict(ver=ver, begin=begin, end=end, seekfree=seekfree, nbfree=nbfree, nfree=nfree, nbname=nbname, units=units,
               compress=compress, seekinfo=seekinfo, nbinfo=nbinfo)
    keys, pos = [], begin
    while pos < end:
        nbytes, version, objlen, dt, keylen, cycle, seekkey, seekpdir = struct.unpack(">ihiIhhii", f[pos:pos + 26])
        assert nbytes > 0 and version == 4
        b = Buf(f[pos + 26:pos + keylen])
        cls, name, title = b.tstring(), b.tstring(), b.tstring()
        keys.append(dict(pos=pos, nbytes=nbytes, objlen=objlen, keylen=keylen, cycle=cycle, seekkey=seekkey,
                         seekpdir=seekpdir, cls=cls, name=name, title=title, hdrlen=26 + b.p))
        pos += nbytes
    assert pos == end
    return hdr, keys


def key_data(f, k):
    raw = f[k["pos"] + k["keylen"]:k["pos"] + k["nbytes"]]
    if len(raw) == k["objlen"]:
        return raw
    if raw[:2] == b"L4":
        return l4_unframe(raw, k["objlen"])
    out, q = b"", 0
    while len(out) < k["objlen"]:
        assert raw[q:q + 2] == b"ZL"
        csz = raw[q + 3] | raw[q + 4] << 8 | raw[q + 5] << 16
        out += zlib.decompress(raw[q + 9:q + 9 + csz])
        q += 9 + csz
    return out


def check_file_records(f):
    """header / directory / key list / streamer info / free segments are consistent with each other."""
    hdr, keys = read_keys(f)
    assert hdr["end"] == len(f) and hdr["begin"] == 100 and hdr["units"] == 4
    d = keys[0]
    assert d["cls"] == "TFile" and d["pos"] == 100 and d["seekkey"] == 100 and d["seekpdir"] == 0
    b = Buf(key_data(f, d))
    name, title = b.tstring(), b.tstring()
    assert name == d["name"] and title == ""
    assert hdr["nbname"] == d["keylen"] + b.p
    ver, dc, dm, nbkeys, nbname, seekdir, seekparent, seekkeys = b.rd("hIIiiiii")
    assert ver == 5 and seekdir == 100 and seekparent == 0 and nbname == hdr["nbname"]
    uuid = b.b[b.p:b.p + 18]
    assert uuid == f[45:63] and len(b.b) - b.p - 18 == 12
    # key list
    kl = [k for k in keys if k["pos"] == seekkeys]
    assert len(kl) == 1 and kl[0]["nbytes"] == nbkeys and kl[0]["cls"] == "TFile"
    kb = Buf(key_data(f, kl[0]))
    listed = []
    for _ in range(kb.rd("i")):
        nbytes, version, objlen, dt, keylen, cycle, seekkey, seekpdir = kb.rd("ihiIhhii")
        cls, nm, ti = kb.tstring(), kb.tstring(), kb.tstring()
        listed.append((seekkey, cls, nm))
        real = [k for k in keys if k["pos"] == seekkey][0]
        assert (real["nbytes"], real["objlen"], real["keylen"], real["cls"], real["name"]) == (nbytes, objlen, keylen, cls, nm)
        assert seekpdir == 100
    assert kb.p == len(kb.b)
    # streamer info: a TList
    si = [k for k in keys if k["pos"] == hdr["seekinfo"]][0]
    assert si["cls"] == "TList" and si["name"] == "StreamerInfo" and si["nbytes"] == hdr["nbinfo"]
    # free segments: one, from the end of the file
    fr = [k for k in keys if k["pos"] == hdr["seekfree"]][0]
    assert fr["nbytes"] == hdr["nbfree"] and hdr["nfree"] == 1
    v, first, last = struct.unpack(">hii", key_data(f, fr))
    assert v == 1 and first == hdr["end"] and last == 2000000000
    return hdr, keys, listed


def test_th2d_roundtrip_and_records(tmp_path):
    from upcgen_b200 import capi
    rng = np.random.default_rng(1)
    nm, ny = 37, 11
    t0, t1 = rng.random((nm, ny)) * 1e3, rng.random((nm, ny))
    path = str(tmp_path / "twoPhotonLumiPol.root")
    capi.root_write_th2d(path, {"hD2LDMDY_s": t0, "hD2LDMDY_p": t1}, nm, 3.56, 50.0, ny, -6.0, 6.0)
    f = open(path, "rb").read()
    hdr, keys, listed = check_file_records(f)
    assert [(c, n) for _, c, n in listed] == [("TH2D", "hD2LDMDY_s"), ("TH2D", "hD2LDMDY_p")]
    # read back with the product's reader (the one the light-by-light plug-in uses on the reference's files)
    for name, t in (("hD2LDMDY_s", t0), ("hD2LDMDY_p", t1)):
        h = capi.root_hist_read(path, name)
        assert (h["dim"], h["nx"], h["ny"]) == (2, nm, ny)
        assert (h["xlo"], h["xhi"], h["ylo"], h["yhi"]) == (3.56, 50.0, -6.0, 6.0)
        assert np.array_equal(h["cells"][1:-1, 1:-1].T, t)       # bin (im + 1, iy + 1) = table[im][iy]
        assert np.all(h["cells"][0] == 0) and np.all(h["cells"][:, 0] == 0)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
def test_th2d_bytes_equal_what_root_wrote(tmp_path):
    """Same name, axes and cells as the TH2D in the reference's cross_section_zm.root -> the same streamed object,
    byte for byte, as ROOT 6.22/09 produced."""
    from upcgen_b200 import capi
    src = os.path.join(REF, "lbyl", "cross_section_zm.root")
    f = open(src, "rb").read()
    _, keys = read_keys(f)
    k = [k for k in keys if k["cls"] == "TH2D"][0]
    real = key_data(f, k)
    h = capi.root_hist_read(src, k["name"])
    nx, ny = h["nx"], h["ny"]
    path = str(tmp_path / "copy.root")
    # all (nx + 2) x (ny + 2) cells, the overflow row included (this histogram has entries there)
    import ctypes as C
    L = capi.lib()
    cells = np.ascontiguousarray(h["cells"], dtype=np.float64)
    names = (C.c_char_p * 1)(k["name"].encode())
    ptrs = (C.c_void_p * 1)(cells.ctypes.data)
    L.upcgpu_root_write_th2d.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                         C.c_double, C.c_double, C.c_void_p]
    assert L.upcgpu_root_write_th2d(path.encode(), 1, names, nx, h["xlo"], h["xhi"], ny, h["ylo"], h["yhi"], ptrs) == 0
    g = open(path, "rb").read()
    _, keys2 = read_keys(g)
    mine = key_data(g, [q for q in keys2 if q["cls"] == "TH2D"][0])
    assert len(mine) == len(real)
    diff = [i for i in range(len(real)) if real[i] != mine[i]]
    assert not diff, (len(diff), diff[:20])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
def test_record_checker_accepts_files_root_wrote():
    """check_file_records -- the statement of how header, directory, key list, streamer info and free segments hang
    together that the writer's output is held to -- holds for the four files ROOT itself wrote for the reference."""
    for p in ("lbyl/cross_section_zm.root", "lbyl/cross_section_m.root", "pi0pi0/cross_section_m.root", "pi0pi0/cross_section_zm.root"):
        hdr, keys, listed = check_file_records(open(os.path.join(REF, p), "rb").read())
        assert len(listed) == 1 and listed[0][1] in ("TH1D", "TH2D")


# ---------------------------------------------------------------------------------------------------------------
# an independent reader of flat TTrees (TTree v20, TBranch v13, TLeaf{I,D}, TBasket) following ROOT's streaming rules
def read_tree(f, name):
    hdr, keys, listed = check_file_records(f)
    k = [k for k in keys if k["cls"] == "TTree" and k["name"] == name][0]
    b = Buf(key_data(f, k), base=k["keylen"])
    classes = {}   # tag position -> class name
    objects = {}   # tag position -> leaf dict

    def tnamed():
        end, v = b.obj()
        assert v == 1
        assert b.rd("H") == 1
        b.rd("I"); b.rd("I")
        n, t = b.tstring(), b.tstring()
        assert b.p == end
        return n, t

    def skip_obj(version):
        end, v = b.obj()
        assert v == version
        b.p = end

    def tobjarray():
        end, v = b.obj()
        assert v == 3
        assert b.rd("H") == 1
        b.rd("I"); b.rd("I")
        assert b.tstring() == ""
        n, low = b.rd("i"), b.rd("i")
        assert low == 0
        return end, n

    def object_any():
        """WriteObjectAny: byte count, class tag (new: 0xFFFFFFFF + name; known: position | 0x80000000), object."""
        start = b.p
        first = b.rd("I")
        if first == 0:
            return None, None, start
        if not (first & 0x40000000):          # reference to an object already streamed
            return "ref", first, start
        end = b.p + (first & 0x3FFFFFFF)
        tagpos = b.p
        tag = b.rd("I")
        if tag == 0xFFFFFFFF:
            cls = b.cstring()
            classes[b.base + tagpos + 2] = cls
        else:
            assert tag & 0x80000000
            cls = classes[tag & 0x7FFFFFFF]
        return cls, end, start

    def tio_features():
        end, v = b.obj()
        b.p = end

    end_tree, v = b.obj()
    assert v == 20
    tname, ttitle = tnamed()
    skip_obj(2); skip_obj(2); skip_obj(2)
    entries, totbytes, zipbytes, saved, flushed = b.rd("qqqqq")
    weight = b.rd("d")
    timer, scan, update, deflen, ncluster = b.rd("iiiii")
    maxentries, maxloop, maxvirt, autosave, autoflush, estimate = b.rd("qqqqqq")
    assert ncluster == 0 and b.rd("B") == 0 and b.rd("B") == 0
    tio_features()
    end_br, nbr = tobjarray()
    branches = []
    for _ in range(nbr):
        cls, end, start = object_any()
        assert cls == "TBranch"
        e2, v = b.obj()
        assert v == 13
        bname, btitle = tnamed()
        skip_obj(2)
        compress, basketsize, entryoffsetlen, writebasket = b.rd("iiii")
        entrynumber = b.rd("q")
        tio_features()
        offset, maxbaskets, splitlevel = b.rd("iii")
        bentries, firstentry, btot, bzip = b.rd("qqqq")
        e3, n3 = tobjarray(); assert n3 == 0 and b.p == e3
        e4, nleaves = tobjarray()
        assert nleaves == 1
        lcls, lend, lstart = object_any()
        assert lcls in ("TLeafI", "TLeafD")
        e5, v5 = b.obj(); assert v5 == 1
        e6, v6 = b.obj(); assert v6 == 2
        lname, ltitle = tnamed()
        flen, lentype, loffset, isrange, isunsigned, leafcount = b.rd("iiiBBI")
        assert b.p == e6 and flen == 1 and leafcount == 0
        b.rd("ii" if lcls == "TLeafI" else "dd")
        assert b.p == e5 == lend and b.p == e4
        objects[b.base + lstart + 2] = dict(name=lname, cls=lcls)
        e7, n7 = tobjarray(); assert n7 == 0 and b.p == e7
        assert b.rd("B") == 1
        bbytes = [b.rd("i") for _ in range(maxbaskets)]
        assert b.rd("B") == 1
        bentry = [b.rd("q") for _ in range(maxbaskets)]
        assert b.rd("B") == 1
        bseek = [b.rd("q") for _ in range(maxbaskets)]
        assert b.tstring() == ""
        assert b.p == e2 == end
        assert entrynumber == bentries == entries and compress == hdr["compress"]
        branches.append(dict(name=bname, title=btitle, leaf=lcls, lentype=lentype, nbaskets=writebasket, bytes=bbytes,
                             entry=bentry, seek=bseek, tot=btot, zip=bzip))
    assert b.p == end_br
    end_lv, nlv = tobjarray()
    assert nlv == nbr
    for i in range(nlv):
        kind, ref, _ = object_any()
        assert kind == "ref" and objects[ref]["name"] == branches[i]["name"]
    assert b.p == end_lv
    assert b.rd("I") == 0 and b.rd("i") == 0 and b.rd("i") == 0
    assert b.rd("IIII") == (0, 0, 0, 0)
    assert b.p == end_tree == len(b.b)
    # baskets
    cols = {}
    by_pos = {k["pos"]: k for k in keys}
    total = total_unzipped = 0
    for br in branches:
        vals = []
        br_tot = br_zip = 0
        for i in range(br["nbaskets"]):
            k = by_pos[br["seek"][i]]
            assert k["cls"] == "TBasket" and k["name"] == br["name"] and k["title"] == name and k["nbytes"] == br["bytes"][i]
            hb = Buf(f[k["pos"] + k["hdrlen"]:k["pos"] + k["keylen"]])
            ver, bufsize, nevsize, nev, last, flag = hb.rd("hiiiiB")
            assert hb.p == len(hb.b) and ver == 3 and flag == 0
            assert nevsize == br["lentype"] and last == k["keylen"] + k["objlen"] and k["objlen"] == nev * nevsize
            nxt = br["entry"][i + 1]
            assert br["entry"][i] + nev == nxt
            dt = ">i4" if br["leaf"] == "TLeafI" else ">f8"
            vals.append(np.frombuffer(key_data(f, k), dtype=dt, count=nev))     # inflated if the basket is an L4 record
            total += k["nbytes"]; br_zip += k["nbytes"]
            total_unzipped += k["keylen"] + k["objlen"]; br_tot += k["keylen"] + k["objlen"]
            assert k["nbytes"] <= k["keylen"] + k["objlen"]
        assert br["entry"][br["nbaskets"]] == entries and (br["tot"], br["zip"]) == (br_tot, br_zip)
        cols[br["name"]] = (br["title"], np.concatenate(vals) if vals else np.zeros(0))
    assert total == zipbytes and total_unzipped == totbytes
    if hdr["compress"] == 0:
        assert totbytes == zipbytes
    assert all(k[1] != "TBasket" for k in listed)       # baskets are not in the directory's key list
    return dict(name=tname, title=ttitle, entries=entries, autosave=autosave, cols=cols, totbytes=totbytes, zipbytes=zipbytes)


def test_tree_roundtrip(tmp_path):
    """events.root as src/UpcGenerator.cpp:842-857 declares it: tree "particles", nine branches."""
    from upcgen_b200 import capi
    rng = np.random.default_rng(2)
    n = 2_300_001    # more than one basket per branch (2^20 entries each)
    cols = {"eventNumber": ("I", np.repeat(np.arange(n // 3 + 1), 3)[:n]), "pdgCode": ("I", rng.choice([13, -13, 22], n)),
            "particleID": ("I", np.tile([1, 2, 3], n // 3 + 1)[:n]), "statusID": ("I", np.full(n, 23)),
            "motherID": ("I", rng.integers(0, 3, n)), "px": ("D", rng.normal(size=n)), "py": ("D", rng.normal(size=n)),
            "pz": ("D", rng.normal(size=n) * 50), "e": ("D", rng.random(n) * 100)}
    path = str(tmp_path / "events.root")
    capi.root_write_tree(path, "particles", "Generated particles", cols)
    t = read_tree(open(path, "rb").read(), "particles")
    assert t["name"] == "particles" and t["title"] == "Generated particles" and t["entries"] == n and t["autosave"] == 0
    assert list(t["cols"]) == list(cols)
    for name, (typ, vals) in cols.items():
        title, got = t["cols"][name]
        assert title == f"{name}/{typ}"
        assert np.array_equal(got, np.asarray(vals, dtype=got.dtype.newbyteorder("=")))


def test_tree_empty_and_small(tmp_path):
    from upcgen_b200 import capi
    for n in (0, 1, 5):
        path = str(tmp_path / f"e{n}.root")
        capi.root_write_tree(path, "particles", "Generated particles", {"pdgCode": ("I", np.arange(n)), "e": ("D", np.arange(n) * 0.5)})
        t = read_tree(open(path, "rb").read(), "particles")
        assert t["entries"] == n
        assert np.array_equal(t["cols"]["e"][1], np.arange(n) * 0.5)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
def test_th1d_bytes_equal_what_root_wrote(tmp_path):
    """The TH1D the writer streams (upcgpu_root_write_th1d) is, byte for byte, the hCrossSectionM that ROOT 6.22/09 wrote
    into the reference's cross_sections/lbyl/cross_section_m.root when given its name, axis and cells; and the reader gets
    the cells back."""
    from upcgen_b200 import capi
    src = os.path.join(REF, "lbyl", "cross_section_m.root")
    f = open(src, "rb").read()
    _, keys = read_keys(f)
    k = [k for k in keys if k["cls"] == "TH1D"][0]
    real = key_data(f, k)
    h = capi.root_hist_read(src, k["name"])
    path = str(tmp_path / "copy.root")
    capi.root_write_th1d(path, k["name"], None, h["xlo"], h["xhi"], cells=h["cells"].ravel())
    g = open(path, "rb").read()
    check_file_records(g)
    _, keys2 = read_keys(g)
    mine = key_data(g, [q for q in keys2 if q["cls"] == "TH1D"][0])
    assert mine == real
    back = capi.root_hist_read(path, k["name"])
    assert back["dim"] == 1 and np.array_equal(back["cells"], h["cells"])


def parse_hist(data, cls):
    """Independent parse of a streamed TH1D / TH2D (uniform or variable bins): name, the three axes with their fXbins,
    entries, the four (seven) statistics sums, the cells.  Member order: TH1 v8 / TH2 v5 / TAxis v10 (ROOT 6.2x)."""
    b = Buf(data)

    def skip():
        e, _ = b.obj()
        b.p = e

    def tnamed():
        e, v = b.obj()
        assert v == 1
        assert b.rd("H") == 1            # TObject version
        b.rd("I"); b.rd("I")             # fUniqueID, fBits
        n, t = b.tstring(), b.tstring()
        assert b.p == e
        return n, t

    def axis():
        e, v = b.obj()
        assert v == 10
        name, _ = tnamed()
        skip()                           # TAttAxis
        nb, lo, hi = b.rd("i"), b.rd("d"), b.rd("d")
        ne = b.rd("i")
        edges = np.array([b.rd("d") for _ in range(ne)])
        b.p = e
        return dict(name=name, n=nb, lo=lo, hi=hi, edges=edges)

    def th1():
        e, v = b.obj()
        assert v == 8
        name, title = tnamed()
        skip(); skip(); skip()           # TAttLine, TAttFill, TAttMarker
        ncells = b.rd("i")
        ax = [axis(), axis(), axis()]
        b.rd("h"); b.rd("h")             # fBarOffset, fBarWidth
        entries = b.rd("d")
        stats = [b.rd("d") for _ in range(4)]
        b.p = e
        return dict(name=name, title=title, ncells=ncells, axes=ax, entries=entries, stats=stats)

    e0, v0 = b.obj()
    if cls == "TH2D":
        assert v0 == 4
        e2, v2 = b.obj()
        assert v2 == 5
        h = th1()
        scale = b.rd("d")
        assert scale == 1.0
        h["stats"] += [b.rd("d") for _ in range(3)]
        assert b.p == e2
    else:
        assert v0 == 3
        h = th1()
    n = b.rd("i")
    assert n == h["ncells"]
    h["cells"] = np.frombuffer(data, dtype=">f8", count=n, offset=b.p).astype(np.float64)
    assert b.p + 8 * n == e0 == len(data)
    return h


def test_sigma_debug_histograms(tmp_path):
    """What the reference adds to events.root at debug level > 0 (src/UpcGenerator.cpp:900-917): TH2D hNucCSYM over
    the (y, m) bin edges (variable-bin axes), bin (iy + 1, im + 1) = cs[iy][im], entries ny * nm, no statistics; then
    ProjectionY() -> hNucCSYM_py over m and ProjectionX() -> hNucCSYM_px over y, written first.  The projections are
    checked against TH2::DoProjection's arithmetic restated here: cells = sequential sums over the integrated axis,
    statistics recomputed from the cells at the bin centres (TH1::ResetStats), entries floor(total + 0.5)."""
    from upcgen_b200 import capi
    rng = np.random.default_rng(7)
    ny, nm = 12, 29
    ye = -6.0 + 1.0 * np.arange(ny + 1)
    me = 3.56 + (50 - 3.56) / nm * np.arange(nm + 1)
    cs = rng.random((ny, nm)) * 1e4
    cs[3, :] = 0.0
    path = str(tmp_path / "events.root")
    capi.root_write_sigma_hists(path, ye, me, cs)
    f = open(path, "rb").read()
    hdr, keys, listed = check_file_records(f)
    assert [(c, n) for _, c, n in listed] == [("TH1D", "hNucCSYM_py"), ("TH1D", "hNucCSYM_px"), ("TH2D", "hNucCSYM")]
    by_name = {k["name"]: k for k in keys if k["cls"] in ("TH1D", "TH2D")}
    h2 = parse_hist(key_data(f, by_name["hNucCSYM"]), "TH2D")
    assert h2["title"] == "" and h2["entries"] == ny * nm and h2["stats"] == [0.0] * 7
    ax, ay, az = h2["axes"]
    assert (ax["name"], ax["n"], ax["lo"], ax["hi"]) == ("xaxis", ny, ye[0], ye[-1]) and np.array_equal(ax["edges"], ye)
    assert (ay["name"], ay["n"], ay["lo"], ay["hi"]) == ("yaxis", nm, me[0], me[-1]) and np.array_equal(ay["edges"], me)
    assert (az["n"], az["lo"], az["hi"], az["edges"].size) == (1, 0.0, 1.0, 0)
    cells = h2["cells"].reshape(nm + 2, ny + 2)          # x (= y_pair) fastest
    assert np.array_equal(cells[1:-1, 1:-1], cs.T)
    assert np.all(cells[0] == 0) and np.all(cells[-1] == 0) and np.all(cells[:, 0] == 0) and np.all(cells[:, -1] == 0)

    def projection(onx):
        edges = ye if onx else me
        nout = edges.size - 1
        c = np.zeros(nout + 2)
        tot = 0.0
        for ob in range(nout + 2):
            cont = 0.0
            line = cells[:, ob] if onx else cells[ob, :]
            for v in line:
                cont += float(v)
            c[ob] = cont
            tot += cont
        st = [0.0, 0.0, 0.0, 0.0]
        for bn in range(1, nout + 1):
            x = 0.5 * (edges[bn - 1] + edges[bn])
            w = c[bn]
            err = abs(np.sqrt(abs(w)))
            st[0] += w; st[1] += err * err; st[2] += w * x; st[3] += w * x * x
        return c, st, np.floor(tot + 0.5)

    for name, onx, edges in (("hNucCSYM_py", False, me), ("hNucCSYM_px", True, ye)):
        h = parse_hist(key_data(f, by_name[name]), "TH1D")
        c, st, ent = projection(onx)
        assert h["title"] == "" and np.array_equal(h["cells"], c) and h["stats"] == [float(s) for s in st] and h["entries"] == ent
        a = h["axes"][0]
        assert (a["n"], a["lo"], a["hi"]) == (edges.size - 1, edges[0], edges[-1]) and np.array_equal(a["edges"], edges)
        assert (h["axes"][1]["n"], h["axes"][1]["edges"].size) == (1, 0)
        # the product's reader (TAxis::FindBin on variable bins) finds the cells again
        back = capi.root_hist_read(path, name)
        assert back["dim"] == 1 and np.array_equal(back["cells"].ravel(), c)
    assert h2["cells"].sum() == pytest.approx(cs.sum(), rel=1e-13)
    back = capi.root_hist_read(path, "hNucCSYM")
    assert back["dim"] == 2 and (back["nx"], back["ny"]) == (ny, nm) and np.array_equal(back["cells"], cells)


def test_uniform_histograms_parse_with_the_same_parser(tmp_path):
    """parse_hist on the writer's uniform-axis objects (the ones pinned byte for byte against ROOT's own): no fXbins."""
    from upcgen_b200 import capi
    path = str(tmp_path / "h.root")
    t = np.arange(15.0).reshape(5, 3)
    capi.root_write_th2d(path, {"hD2LDMDY": t}, 5, 1.0, 2.0, 3, -1.0, 1.0)
    f = open(path, "rb").read()
    _, keys = read_keys(f)
    h = parse_hist(key_data(f, [k for k in keys if k["cls"] == "TH2D"][0]), "TH2D")
    assert h["name"] == "hD2LDMDY" and h["axes"][0]["edges"].size == 0 and h["axes"][0]["n"] == 5 and h["entries"] == 15
    assert np.array_equal(h["cells"].reshape(5, 7)[1:-1, 1:-1], t.T)


# ---- LZ4 ("L4" records) ---------------------------------------------------------------------------------------
def lz4_block_decode(src, out_n):
    """Independent decoder of one LZ4 block (lz4_Block_format.md): token, literal run, offset, match."""
    out = bytearray()
    ip, n = 0, len(src)
    while ip < n:
        tok = src[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[ip]; ip += 1
                lit += b
                if b != 255:
                    break
        out += src[ip:ip + lit]
        assert ip + lit <= n
        ip += lit
        if ip >= n:
            break
        off = src[ip] | src[ip + 1] << 8
        ip += 2
        assert 0 < off <= len(out)
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[ip]; ip += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        if off >= ml:
            out += out[len(out) - off:len(out) - off + ml]
        else:
            for _ in range(ml):
                out.append(out[-off])
    assert len(out) == out_n
    return bytes(out)


def l4_unframe(raw, objlen):
    """ROOT's L4 framing: 9-byte header, 8-byte big-endian XXH64 of the LZ4 bytes, the LZ4 block; chained."""
    import xxhash
    out, q = b"", 0
    while len(out) < objlen:
        assert raw[q:q + 2] == b"L4" and raw[q + 2] == 1
        csz = raw[q + 3] | raw[q + 4] << 8 | raw[q + 5] << 16
        usz = raw[q + 6] | raw[q + 7] << 8 | raw[q + 8] << 16
        payload = raw[q + 17:q + 9 + csz]
        assert int.from_bytes(raw[q + 9:q + 17], "big") == xxhash.xxh64(payload, seed=0).intdigest()
        out += lz4_block_decode(payload, usz)
        q += 9 + csz
    assert q == len(raw) and len(out) == objlen
    return out


@pytest.fixture(scope="module")
def lz4_check(tmp_path_factory):
    import subprocess
    d = tmp_path_factory.mktemp("lz4")
    exe = str(d / "lz4_check")
    host = os.path.join(ROOT_DIR, "upcgen_b200", "host")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", host, "-o", exe,
                           os.path.join(ROOT_DIR, "tests", "cpp", "lz4_check.cpp"), os.path.join(host, "UpcLz4.cpp")])

    def run(data: bytes):
        import json
        import subprocess as sp
        (d / "in.bin").write_bytes(data)
        r = sp.run([exe, str(d / "in.bin"), str(d / "out")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout), (d / "out.lz4").read_bytes(), (d / "out.l4").read_bytes()

    return run


def test_lz4_and_xxh64_against_independent_implementations(lz4_check):
    """UpcLz4.cpp: XXH64 equals the xxhash module's on every length class (0, < 4, < 8, < 32, 32 k + r); the compressor's
    blocks are decoded by the independent decoder above -- short inputs that must stay literal (the last 5 bytes, no
    match in the last 12), literal runs and matches beyond the 15 / 270 thresholds of the length encoding, overlapping
    matches (runs), incompressible bytes; the L4 framing carries the right sizes and checksum, chains blocks above
    16 MB, and refuses data that does not shrink (ROOT stores those as they are)."""
    import xxhash
    rng = np.random.default_rng(11)
    ints = np.repeat(np.arange(3000, dtype=">i4"), 3).tobytes()
    cases = [b"", b"a", b"abc", b"abcdefg", b"0123456789a", b"0123456789ab", b"0123456789abc", b"x" * 12, b"x" * 13,
             b"ab" * 40, b"x" * 5000, rng.bytes(3000), bytes(range(256)) * 3 + b"q" * 700 + bytes(range(256)) * 2,
             ints, rng.bytes(300) + b"\0" * 100000 + rng.bytes(20), b"abcdefghij" * 31 + rng.bytes(31) + b"abcdefghij" * 7]
    for data in cases:
        res, blk, framed = lz4_check(data)
        assert res["n"] == len(data)
        assert int(res["xxh64"], 16) == xxhash.xxh64(data, seed=0).intdigest(), len(data)
        assert res["block_roundtrip"] and res["frames_ok"]
        assert lz4_block_decode(blk, len(data)) == data
        if res["shrunk"]:
            assert len(framed) < len(data) and l4_unframe(framed, len(data)) == data
        else:
            assert framed == b""
    assert not lz4_check(rng.bytes(3000))[0]["shrunk"] and lz4_check(ints)[0]["shrunk"]
    # more than one 0xffffff-byte chunk
    big = np.repeat(np.arange(1_500_000, dtype=">i4"), 3).tobytes()        # 18 MB
    res, blk, framed = lz4_check(big)
    assert res["block_roundtrip"] and res["frames_ok"] and res["shrunk"]
    assert framed[6] | framed[7] << 8 | framed[8] << 16 == 0xffffff
    assert l4_unframe(framed, len(big)) == big


def test_lz4_compressed_files_roundtrip(tmp_path):
    """upcgpu_root_set_compression(409) -- the reference's setting for events.root (src/UpcGenerator.cpp:843): the tree's
    baskets and the objects above 256 bytes become L4 records (sizes, XXH64 checksum and LZ4 stream checked by the
    independent decoders above), fCompress is recorded in the file header and in every branch, fTotBytes / fZipBytes
    account for both sizes, the columns and histograms come back bit for bit -- through the Python parser and through
    the product's reader (UpcRootHist.cpp inflates L4 next to zlib).  Incompressible columns stay as they are."""
    from upcgen_b200 import capi
    rng = np.random.default_rng(5)
    n = 1_200_001
    cols = {"eventNumber": ("I", np.repeat(np.arange(n // 3 + 1), 3)[:n]), "pdgCode": ("I", rng.choice([13, -13, 22], n)),
            "particleID": ("I", np.tile([1, 2, 3], n // 3 + 1)[:n]), "statusID": ("I", np.full(n, 23)),
            "motherID": ("I", rng.integers(0, 3, n)), "px": ("D", rng.normal(size=n)), "e": ("D", np.round(rng.random(n) * 100, 1))}
    assert capi.root_set_compression(409) == 0
    try:
        with pytest.raises(capi.UpcGpuError):
            capi.root_set_compression(505)                # zstd: neither read nor written
        path = str(tmp_path / "events.root")
        capi.root_write_tree(path, "particles", "Generated particles", cols)
        hpath = str(tmp_path / "twoPhotonLumi.root")
        table = np.outer(np.linspace(1, 2, 300), np.ones(40))          # smooth: compresses well
        capi.root_write_th2d(hpath, {"hD2LDMDY": table}, 300, 3.56, 50.0, 40, -6.0, 6.0)
    finally:
        assert capi.root_set_compression(0) == 409
    f = open(path, "rb").read()
    hdr, keys = read_keys(f)
    assert hdr["compress"] == 409
    t = read_tree(f, "particles")
    assert t["entries"] == n and t["zipbytes"] < 0.62 * t["totbytes"]
    for name, (typ, vals) in cols.items():
        got = t["cols"][name][1]
        assert np.array_equal(got, np.asarray(vals, dtype=got.dtype.newbyteorder("=")))
    size = {name: sum(k["nbytes"] for k in keys if k["cls"] == "TBasket" and k["name"] == name) for name in cols}
    raw = {name: sum(k["keylen"] + k["objlen"] for k in keys if k["cls"] == "TBasket" and k["name"] == name) for name in cols}
    assert size["statusID"] < 0.01 * raw["statusID"] and size["eventNumber"] < 0.65 * raw["eventNumber"]
    assert size["px"] == raw["px"]                       # random mantissas: stored as they are
    tk = [k for k in keys if k["cls"] == "TTree"][0]
    assert tk["nbytes"] < tk["keylen"] + tk["objlen"]    # the tree record itself is above 256 bytes and shrinks
    g = open(hpath, "rb").read()
    hh, hkeys = read_keys(g)
    check_file_records(g)
    hk = [k for k in hkeys if k["cls"] == "TH2D"][0]
    assert hh["compress"] == 409 and hk["nbytes"] < 0.2 * (hk["keylen"] + hk["objlen"])
    h = parse_hist(key_data(g, hk), "TH2D")
    assert np.array_equal(h["cells"].reshape(42, 302)[1:-1, 1:-1], table.T)
    back = capi.root_hist_read(hpath, "hD2LDMDY")
    assert np.array_equal(back["cells"][1:-1, 1:-1].T, table)
    # and with the default setting nothing is compressed
    capi.root_write_th2d(hpath, {"hD2LDMDY": table}, 300, 3.56, 50.0, 40, -6.0, 6.0)
    g = open(hpath, "rb").read()
    hh, hkeys = read_keys(g)
    hk = [k for k in hkeys if k["cls"] == "TH2D"][0]
    assert hh["compress"] == 0 and hk["nbytes"] == hk["keylen"] + hk["objlen"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference's cross_sections directory is not mounted")
def test_zlib_records_have_roots_framing(tmp_path):
    """Setting 101 (ROOT's default, the one the reference's luminosity cache is written with): the TH2D record of the
    reference's own cross_sections/lbyl/cross_section_zm.root, written again with the same name, axes and cells, has
    the same key header fields and the same 'ZL' 8 framing as the record ROOT 6.22/09 wrote, and inflates to the same
    bytes; the compressed stream itself may differ with the zlib build (reported, not required)."""
    from upcgen_b200 import capi
    src = os.path.join(REF, "lbyl", "cross_section_zm.root")
    f = open(src, "rb").read()
    hdr, keys = read_keys(f)
    k = [q for q in keys if q["cls"] == "TH2D"][0]
    real_raw = f[k["pos"] + k["keylen"]:k["pos"] + k["nbytes"]]
    real = key_data(f, k)
    h = capi.root_hist_read(src, k["name"])
    assert capi.root_set_compression(hdr["compress"]) == 0 and hdr["compress"] == 101
    try:
        path = str(tmp_path / "copy.root")
        import ctypes as C
        L = capi.lib()
        cells = np.ascontiguousarray(h["cells"], dtype=np.float64)     # [ny + 2][nx + 2]: the overflow row has entries
        names = (C.c_char_p * 1)(k["name"].encode())
        ptrs = (C.c_void_p * 1)(cells.ctypes.data)
        L.upcgpu_root_write_th2d.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                             C.c_double, C.c_double, C.c_void_p]
        assert L.upcgpu_root_write_th2d(path.encode(), 1, names, h["nx"], h["xlo"], h["xhi"], h["ny"], h["ylo"], h["yhi"], ptrs) == 0
    finally:
        capi.root_set_compression(0)
    g = open(path, "rb").read()
    check_file_records(g)
    hdr2, keys2 = read_keys(g)
    k2 = [q for q in keys2 if q["cls"] == "TH2D"][0]
    mine_raw = g[k2["pos"] + k2["keylen"]:k2["pos"] + k2["nbytes"]]
    assert hdr2["compress"] == 101 and (k2["objlen"], k2["keylen"]) == (k["objlen"], k["keylen"])
    assert mine_raw[:3] == real_raw[:3] == b"ZL\x08" and mine_raw[6:9] == real_raw[6:9]     # method, uncompressed size
    assert mine_raw[9:11] == real_raw[9:11]                                                 # zlib header of a level-1 stream
    assert key_data(g, k2) == real
    print("compressed bytes: ROOT", len(real_raw), "here", len(mine_raw), "identical" if mine_raw == real_raw else "different stream")
    back = capi.root_hist_read(path, k["name"])
    assert np.array_equal(back["cells"], h["cells"])
