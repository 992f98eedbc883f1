"""ctypes loader for oracle/_ref/libupcref.so: the REFERENCE's own src/UpcCrossSection.cpp,
src/UpcTwoPhotonDilep.cpp and src/UpcTwoPhotonALP.cpp compiled unmodified against the GSL/ROOT
shim (oracle/refshim).  TEST INFRASTRUCTURE ONLY.  One instance per process (the reference keeps
its splines in file-scope globals)."""
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
