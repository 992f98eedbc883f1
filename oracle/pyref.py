"""ctypes loader for oracle/_ref/libupcref.so: the REFERENCE's own src/UpcCrossSection.cpp,
src/UpcTwoPhotonDilep.cpp, src/UpcTwoPhotonALP.cpp, src/UpcGenerator.cpp and include/UpcSampler.h compiled
unmodified against the GSL/ROOT shim (oracle/refshim).  TEST INFRASTRUCTURE ONLY.  One instance per process (the
reference keeps its splines in file-scope globals)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libupcref.so")


class RefParams(C.Structure):
    _fields_ = [("Z", C.c_int), ("A", C.c_int), ("R", C.c_double), ("a", C.c_double), ("sqrts", C.c_double),
                ("is_point", C.c_int), ("breakup_mode", C.c_int), ("use_pol", C.c_int), ("nm", C.c_int), ("ny", C.c_int),
                ("mmin", C.c_double), ("mmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
                ("proc_id", C.c_int), ("a_lep", C.c_double), ("alp_mass", C.c_double), ("alp_width", C.c_double),
                ("with_breakup_table", C.c_int)]


def available():
    return os.path.exists(SO)


class Reference:
    def __init__(self, P, with_breakup_table=False):
        self.L = C.CDLL(SO)
        L = self.L
        d = C.c_double
        for name, args in [("upcref_rho0", []), ("upcref_gtot", []), ("upcref_factor", []), ("upcref_formfac", [d]),
                           ("upcref_formfac_knot", [C.c_int]), ("upcref_flux_point", [d, d]), ("upcref_flux_form", [d, d]),
                           ("upcref_breakup_raw", [d, C.c_int]), ("upcref_breakup_spline", [d]), ("upcref_lumi", [d, d]),
                           ("upcref_photon_flux", [d, d]),
                           ("upcref_sigma_m", [d]), ("upcref_sigma_zm", [d, d]), ("upcref_sigma_m_pol", [d, C.c_int]),
                           ("upcref_sigma_zm_pol", [d, d, C.c_int])]:
            getattr(L, name).restype = d
            getattr(L, name).argtypes = args
        L.upcref_breakup_knot.restype = d
        L.upcref_breakup_knot.argtypes = [C.c_int, C.POINTER(d)]
        L.upcref_lumi_pol.argtypes = [d, d, C.POINTER(d), C.POINTER(d)]
        L.upcref_gaa.argtypes = [C.c_void_p, C.c_void_p]
        L.upcref_grid_and_fold.restype = d
        L.upcref_grid_and_fold.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p]
        rp = RefParams(P.Z, P.A, P.R, P.a, P.sqrts, P.is_point, P.breakup_mode, P.use_pol, P.nm, P.ny, P.mmin, P.mmax,
                       P.ymin, P.ymax, P.proc_id, P.a_lep, P.alp_mass, P.alp_width, int(with_breakup_table))
        self.P = P
        assert L.upcref_init(C.byref(rp)) == 0

    def vm_sigma_y(self, pdg, shadowing, dght_pdg, ys):
        """sigma(y) of the reference's own UpcPhotoNuclearVM::calcCrossSectionY (src/UpcPhotoNuclearVM.cpp:340-381)."""
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        out = np.zeros(ys.size)
        mp = C.c_double()
        self.L.upcref_vm_sigma_y.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        rc = self.L.upcref_vm_sigma_y(pdg, shadowing, dght_pdg, ys.ctypes.data, ys.size, out.ctypes.data, C.byref(mp))
        if rc:
            raise RuntimeError(f"upcref_vm_sigma_y: {rc}")
        return out, mp.value

    def gaa(self):
        y, c = np.zeros(200), np.zeros(200)
        self.L.upcref_gaa(y.ctypes.data, c.ctypes.data)
        return y, c

    def lumi_pol(self, M, Y):
        s, p = C.c_double(), C.c_double()
        self.L.upcref_lumi_pol(M, Y, C.byref(s), C.byref(p))
        return s.value, p.value

    def grid_and_fold(self, directory, nthreads=4):
        cs = np.zeros((self.P.ny, self.P.nm))
        ratio = np.zeros((self.P.ny, self.P.nm))
        tot = self.L.upcref_grid_and_fold(nthreads, directory.encode(), cs.ctypes.data, ratio.ctypes.data)
        return cs, ratio, tot


def put_hist(path, name, cells, xaxis, yaxis=None):
    """Places a histogram into the shim's in-memory "ROOT file" `path` (what TFile(path)->Get(name) of the reference's
    light-by-light / pi0 pi0 plug-ins will find): cells = all (nx + 2) [x (ny + 2)] bins, x fastest; axes (n, lo, hi)."""
    L = C.CDLL(SO)
    cells = np.ascontiguousarray(cells, dtype=np.float64).ravel()
    d, i, p = C.c_double, C.c_int, C.c_void_p
    if yaxis is None:
        assert cells.size == xaxis[0] + 2
        L.upcref_put_hist1.argtypes = [C.c_char_p, C.c_char_p, i, d, d, p]
        L.upcref_put_hist1(path.encode(), name.encode(), int(xaxis[0]), float(xaxis[1]), float(xaxis[2]), cells.ctypes.data)
    else:
        assert cells.size == (xaxis[0] + 2) * (yaxis[0] + 2)
        L.upcref_put_hist2.argtypes = [C.c_char_p, C.c_char_p, i, d, d, i, d, d, p]
        L.upcref_put_hist2(path.encode(), name.encode(), int(xaxis[0]), float(xaxis[1]), float(xaxis[2]), int(yaxis[0]),
                           float(yaxis[1]), float(yaxis[2]), cells.ctypes.data)


def cross_sec_dir():
    """CROSS_SEC_DIR the reference's sources were compiled with (the directory its plug-ins look into)."""
    L = C.CDLL(SO)
    L.upcref_cross_sec_dir.restype = C.c_char_p
    return L.upcref_cross_sec_dir().decode()


class RefGenerator:
    """The reference's own UpcGenerator, driven as main.cpp drives it (setParFile, configGeneratorFromFile, init,
    generateEvent), with the two-photon luminosity cache injected: the table is placed into the shim's in-memory
    twoPhotonLumi[Pol].root, so UpcCrossSection::prepareTwoPhotonLumi takes its "found pre-calculated luminosity"
    branch (src/UpcCrossSection.cpp:481-491) and calcNucCrossSectionYM folds THAT table.  Everything after it --
    fillCrossSectionZM, the UpcSampler constructors, generateEvent with getPairMomentum / getPhotonPt, pair / single
    production, the uniform decay, the cuts, the HepMC writer -- is the reference's code with its own MT19937 streams.

    par_text: parameters.in text (BREAKUP_MODE only matters to the lumi table, which is injected: pass 1 to skip the
    minute-long prepareBreakupProb).  lumi / (lumi_s, lumi_p): [nm][ny] tables, already multiplied by dm * dy."""

    def __init__(self, par_text, workdir, lumi=None, lumi_s=None, lumi_p=None, grid=None):
        self.L = L = C.CDLL(SO)
        d, i, p, lg = C.c_double, C.c_int, C.c_void_p, C.c_long
        L.upcrefgen_put_lumi.argtypes = [C.c_char_p, i, i, i, d, d, d, d, p, p, p]
        L.upcrefgen_create.argtypes = [C.c_char_p, C.c_char_p]
        L.upcrefgen_totcs.restype = d
        L.upcrefgen_grid.argtypes = [C.POINTER(i)] * 3
        L.upcrefgen_cs.argtypes = [p, p]
        L.upcrefgen_cdfs.argtypes = [p, p]
        L.upcrefgen_generate.restype = lg
        L.upcrefgen_generate.argtypes = [lg, p, p, p, p, p]
        L.upcrefgen_sample_ym.argtypes = [lg, p, p, p, p]
        L.upcrefgen_sample_z.argtypes = [i, lg, p]
        L.upcrefgen_photon_pt.argtypes = [d, lg, p]
        L.upcrefgen_tape.argtypes = [i]
        L.upcrefgen_tape_read.restype = lg
        L.upcrefgen_tape_read.argtypes = [p, p, lg]
        L.upcrefgen_generate_events_hepmc.argtypes = [lg]
        L.upcrefgen_generate_events_tree.restype = lg
        L.upcrefgen_generate_events_tree.argtypes = [lg, p, lg]
        os.makedirs(workdir, exist_ok=True)
        self.workdir = workdir
        parfile = os.path.join(workdir, "parameters.in")
        with open(parfile, "w") as fh:
            fh.write(par_text)
        nm, ny, mmin, mmax, ymin, ymax = grid
        pol = lumi is None
        tabs = [None if t is None else np.ascontiguousarray(t, float) for t in (lumi, lumi_s, lumi_p)]
        for t in tabs:
            assert t is None or t.shape == (nm, ny)
        ptr = [None if t is None else t.ctypes.data for t in tabs]
        assert L.upcrefgen_put_lumi(workdir.encode(), int(pol), nm, ny, mmin, mmax, ymin, ymax, *ptr) == 0
        assert L.upcrefgen_create(parfile.encode(), workdir.encode()) == 0
        a, b, c = i(), i(), i()
        L.upcrefgen_grid(C.byref(a), C.byref(b), C.byref(c))
        self.nm, self.ny, self.nz = a.value, b.value, c.value
        assert (self.nm, self.ny) == (nm, ny), "grid of the injected table != grid the reference derived"

    def totcs(self):
        return self.L.upcrefgen_totcs()

    def cs(self):
        cs, ratio = np.zeros((self.ny, self.nm)), np.zeros((self.ny, self.nm))
        self.L.upcrefgen_cs(cs.ctypes.data, ratio.ctypes.data)
        return cs, ratio

    def cdfs(self, with_z=True):
        s2 = np.zeros(self.ny * self.nm + 1)
        sz = np.zeros((self.nm, self.nz + 1)) if with_z else None
        self.L.upcrefgen_cdfs(s2.ctypes.data, None if sz is None else sz.ctypes.data)
        return s2, sz

    def generate(self, n):
        npart = np.zeros(n, np.int32)
        pdg = np.zeros((n, 6), np.int32); st = np.zeros((n, 6), np.int32); mo = np.zeros((n, 6), np.int32)
        p4 = np.zeros((n, 6, 4))
        acc = self.L.upcrefgen_generate(n, npart.ctypes.data, pdg.ctypes.data, st.ctypes.data, mo.ctypes.data,
                                        p4.ctypes.data)
        return dict(npart=npart, pdg=pdg, status=st, mother=mo, p4=p4, n_accepted=acc)

    def sample_ym(self, n):
        y, m = np.zeros(n), np.zeros(n)
        yb, mb = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.upcrefgen_sample_ym(n, y.ctypes.data, m.ctypes.data, yb.ctypes.data, mb.ctypes.data)
        return y, m, yb, mb

    def sample_z(self, mbin, n):
        z = np.zeros(n)
        self.L.upcrefgen_sample_z(mbin, n, z.ctypes.data)
        return z

    def photon_pt(self, e, n):
        pt = np.zeros(n)
        self.L.upcrefgen_photon_pt(float(e), n, pt.ctypes.data)
        return pt

    def reseed_z(self, seed0):
        """Decorrelates the z samplers (the reference seeds all of them alike, Q6): see upcrefgen_reseed_z."""
        self.L.upcrefgen_reseed_z.argtypes = [C.c_ulong]
        self.L.upcrefgen_reseed_z(int(seed0))

    def tape(self, on):
        self.L.upcrefgen_tape(int(on))

    def tape_read(self):
        n = self.L.upcrefgen_tape_read(None, None, 0)
        v, t = np.zeros(n), np.zeros(n, np.int32)
        self.L.upcrefgen_tape_read(v.ctypes.data, t.ctypes.data, n)
        return v, t

    def generate_events_hepmc(self, n_events):
        """UpcGenerator::generateEvents with USE_HEPMC_OUTPUT: returns the text of events.hepmc."""
        cwd = os.getcwd()
        os.chdir(self.workdir)
        try:
            self.L.upcrefgen_generate_events_hepmc(n_events)
            with open("events.hepmc") as fh:
                return fh.read()
        finally:
            os.chdir(cwd)

    def generate_events_tree(self, n_events):
        """... with USE_ROOT_OUTPUT: the rows filled into the `particles` tree, [n_rows][9] in branch order
        (eventNumber, pdgCode, particleID, statusID, motherID, px, py, pz, e)."""
        cwd = os.getcwd()
        os.chdir(self.workdir)
        try:
            cap = int(n_events) * 4 + 16
            out = np.zeros((cap, 9))
            rows = self.L.upcrefgen_generate_events_tree(n_events, out.ctypes.data, cap)
            return out[:rows]
        finally:
            os.chdir(cwd)
