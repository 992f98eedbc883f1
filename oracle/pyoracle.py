"""ctypes loader for the CPU oracle (oracle/libupcoracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OParams(C.Structure):
    _fields_ = [
        ("Z", C.c_int), ("A", C.c_int),
        ("R", C.c_double), ("a", C.c_double),
        ("sqrts", C.c_double), ("g1", C.c_double), ("g2", C.c_double), ("gtot", C.c_double),
        ("is_point", C.c_int), ("breakup_mode", C.c_int), ("use_pol", C.c_int),
        ("nonzero_gam_pt", C.c_int),
        ("nm", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("mmin", C.c_double), ("mmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("zmin", C.c_double), ("zmax", C.c_double),
        ("nb1", C.c_int), ("nb2", C.c_int),
        ("proc_id", C.c_int), ("a_lep", C.c_double),
        ("alp_mass", C.c_double), ("alp_width", C.c_double),
        ("do_pt_cut", C.c_int), ("do_eta_cut", C.c_int),
        ("pt_min", C.c_double), ("eta_min", C.c_double), ("eta_max", C.c_double),
    ]


def build(force=False):
    so = os.path.join(_HERE, "libupcoracle.so")
    src = os.path.join(_HERE, "upc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        d, i, p = C.c_double, C.c_int, C.c_void_p
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.upco_create.restype = p
        L.upco_create.argtypes = [C.POINTER(OParams), i]
        L.upco_destroy.argtypes = [p]
        L.upco_set_threads.argtypes = [p, i]
        for name in ("upco_bessel_K0", "upco_bessel_K1", "upco_bessel_J1", "upco_tmath_besselK1",
                     "upco_tmath_besselI1"):
            getattr(L, name).restype = d
            getattr(L, name).argtypes = [d]
        L.upco_cspline_init.argtypes = [p, p, i, p]
        L.upco_cspline_eval.restype = d
        L.upco_cspline_eval.argtypes = [p, p, p, i, d]
        L.upco_qags_fluxform.restype = d
        L.upco_qags_fluxform.argtypes = [p, d, d, dp, ip, ip, ip]
        L.upco_qags_test.restype = d
        L.upco_qags_test.argtypes = [i, d, d, d, d, d, dp, ip, ip, ip]
        L.upco_rho0.restype = d; L.upco_rho0.argtypes = [p]
        L.upco_sigma_nn.restype = d; L.upco_sigma_nn.argtypes = [p]
        L.upco_get_gaa.argtypes = [p, p, p, p, p]
        L.upco_formfac.restype = d; L.upco_formfac.argtypes = [p, d]
        L.upco_formfac_spline.restype = d; L.upco_formfac_spline.argtypes = [p, d]
        L.upco_get_formfac_table.argtypes = [p, i, i, p, p]
        L.upco_breakup_raw.restype = d; L.upco_breakup_raw.argtypes = [p, d, i]
        L.upco_breakup_spline.restype = d; L.upco_breakup_spline.argtypes = [p, d]
        L.upco_get_breakup_table.argtypes = [p, i, i, p, p]
        L.upco_breakup_nknots_energy.restype = i; L.upco_breakup_nknots_energy.argtypes = [p]
        L.upco_flux_point.restype = d; L.upco_flux_point.argtypes = [p, d, d]
        L.upco_flux_form.restype = d; L.upco_flux_form.argtypes = [p, d, d]
        L.upco_flux_form_batch.argtypes = [p, p, p, C.c_size_t, p, p]
        L.upco_lumi.restype = d; L.upco_lumi.argtypes = [p, d, d]
        L.upco_lumi_pol.argtypes = [p, d, d, dp, dp]
        L.upco_fill_lumi.argtypes = [p, i, i, i, i, p, p, p, p]
        L.upco_sigma_m.restype = d; L.upco_sigma_m.argtypes = [p, d]
        L.upco_sigma_zm.restype = d; L.upco_sigma_zm.argtypes = [p, d, d]
        L.upco_sigma_m_pol.restype = d; L.upco_sigma_m_pol.argtypes = [p, d, i]
        L.upco_sigma_zm_pol.restype = d; L.upco_sigma_zm_pol.argtypes = [p, d, d, i]
        L.upco_fold.argtypes = [p, p, p, p, p, p, dp]
        L.upco_fill_cs_zm.argtypes = [p, i, p]
        L.upco_pdf_init.argtypes = [p, C.c_size_t, p]
        L.upco_pdf_find.restype = C.c_longlong; L.upco_pdf_find.argtypes = [p, C.c_size_t, d]
        L.upco_sample2d.argtypes = [p, i, i, p, p, d, d, C.POINTER(C.c_longlong), dp, dp]
        L.upco_sample1d.restype = d; L.upco_sample1d.argtypes = [p, i, p, d]
        L.upco_get_bin.restype = i; L.upco_get_bin.argtypes = [i, d, d, d]
        L.upco_photon_pt_cdf.argtypes = [p, d, p]
        L.upco_photon_pt_sample.restype = d; L.upco_photon_pt_sample.argtypes = [p, p, d]
        L.upco_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, dp, dp]
        L.upco_generate_event.restype = i
        L.upco_generate_event.argtypes = [p, C.c_uint64, C.c_uint64, p, p, p, p, ip, p, p, p, p, p]
        L.upco_photon_flux.restype = d; L.upco_photon_flux.argtypes = [p, d, d]
        L.upco_generate_event_u.restype = i
        L.upco_generate_event_u.argtypes = [p, p, p, p, p, p, ip, p, p, p, p, p]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def to_oparams(P):
    """P: any object/dict with the field names of OParams (e.g. upcgen_b200.config.UpcParams)."""
    o = OParams()
    get = (lambda k: P[k]) if isinstance(P, dict) else (lambda k: getattr(P, k))
    for name, ctype in OParams._fields_:
        v = get(name)
        setattr(o, name, int(v) if ctype is C.c_int else float(v))
    return o


class Oracle:
    """Thin OO wrapper over the C oracle."""

    def __init__(self, P, nbc=0, threads=None):
        self.L = lib()
        self.op = to_oparams(P)
        self.h = self.L.upco_create(C.byref(self.op), nbc)
        self.nm, self.ny, self.nz = self.op.nm, self.op.ny, self.op.nz
        if threads:
            self.L.upco_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.upco_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # tables
    def rho0(self):
        return self.L.upco_rho0(self.h)

    def sigma_nn(self):
        return self.L.upco_sigma_nn(self.h)

    def gaa(self):
        b, g, c, ta = (np.zeros(200) for _ in range(4))
        self.L.upco_get_gaa(self.h, _ptr(b), _ptr(g), _ptr(c), _ptr(ta))
        return b, g, c, ta

    def formfac(self, q2):
        return np.array([self.L.upco_formfac(self.h, float(x)) for x in np.atleast_1d(q2)])

    def formfac_spline(self, q2):
        return np.array([self.L.upco_formfac_spline(self.h, float(x)) for x in np.atleast_1d(q2)])

    def formfac_table(self, i0, n):
        y, c = np.zeros(n), np.zeros(n)
        self.L.upco_get_formfac_table(self.h, i0, n, _ptr(y), _ptr(c))
        return y, c

    def breakup_raw(self, b, mode):
        return np.array([self.L.upco_breakup_raw(self.h, float(x), mode) for x in np.atleast_1d(b)])

    def breakup_spline(self, b):
        return np.array([self.L.upco_breakup_spline(self.h, float(x)) for x in np.atleast_1d(b)])

    def breakup_table(self, i0, n):
        y, c = np.zeros(n), np.zeros(n)
        self.L.upco_get_breakup_table(self.h, i0, n, _ptr(y), _ptr(c))
        return y, c

    # fluxes
    def flux_point(self, b, k):
        b, k = np.broadcast_arrays(np.asarray(b, float), np.asarray(k, float))
        return np.array([self.L.upco_flux_point(self.h, float(x), float(y)) for x, y in zip(b.ravel(), k.ravel())]).reshape(b.shape)

    def flux_form(self, b, k, with_neval=False):
        b, k = np.broadcast_arrays(np.asarray(b, float), np.asarray(k, float))
        bb = np.ascontiguousarray(b.ravel()); kk = np.ascontiguousarray(k.ravel())
        out = np.zeros(bb.size); ne = np.zeros(bb.size, dtype=np.int32)
        self.L.upco_flux_form_batch(self.h, _ptr(bb), _ptr(kk), bb.size, _ptr(out), _ptr(ne))
        if with_neval:
            return out.reshape(b.shape), ne.reshape(b.shape)
        return out.reshape(b.shape)

    def qags_fluxform(self, b, k):
        err = C.c_double(); ne = C.c_int(); last = C.c_int(); ier = C.c_int()
        r = self.L.upco_qags_fluxform(self.h, b, k, C.byref(err), C.byref(ne), C.byref(last), C.byref(ier))
        return r, err.value, ne.value, last.value, ier.value

    def photon_flux(self, M, Y):
        """calcPhotonFlux (src/UpcCrossSection.cpp:700-722)."""
        return self.L.upco_photon_flux(self.h, float(M), float(Y))

    # lumi
    def lumi(self, M, Y):
        return self.L.upco_lumi(self.h, M, Y)

    def lumi_pol(self, M, Y):
        s = C.c_double(); p = C.c_double()
        self.L.upco_lumi_pol(self.h, M, Y, C.byref(s), C.byref(p))
        return s.value, p.value

    def fill_lumi(self, im0=0, im1=None, im_step=1, iy_step=1, with_neval=False):
        nm, ny = self.nm, self.ny
        im1 = nm if im1 is None else im1
        pol = bool(self.op.use_pol)
        a = np.full((nm, ny), np.nan)
        b = np.full((nm, ny), np.nan) if pol else None
        ne = np.zeros((nm, ny), dtype=np.int64)
        if pol:
            self.L.upco_fill_lumi(self.h, im0, im1, im_step, iy_step, None, _ptr(a), _ptr(b), _ptr(ne))
            res = (a, b)
        else:
            self.L.upco_fill_lumi(self.h, im0, im1, im_step, iy_step, _ptr(a), None, None, _ptr(ne))
            res = a
        return (res, ne) if with_neval else res

    # sigma
    def sigma_m(self, m):
        return np.array([self.L.upco_sigma_m(self.h, float(x)) for x in np.atleast_1d(m)])

    def sigma_m_pol(self, m, ps):
        return np.array([self.L.upco_sigma_m_pol(self.h, float(x), ps) for x in np.atleast_1d(m)])

    def fold(self, lumi, lumi_s=None, lumi_p=None):
        nm, ny = self.nm, self.ny
        cs = np.zeros((ny, nm)); ratio = np.zeros((ny, nm)); tot = C.c_double()
        lumi = None if lumi is None else np.ascontiguousarray(lumi)
        lumi_s = None if lumi_s is None else np.ascontiguousarray(lumi_s)
        lumi_p = None if lumi_p is None else np.ascontiguousarray(lumi_p)
        self.L.upco_fold(self.h, _ptr(lumi), _ptr(lumi_s), _ptr(lumi_p), _ptr(cs), _ptr(ratio), C.byref(tot))
        return cs, ratio, tot.value

    def cs_zm(self, flag=0):
        out = np.zeros((self.nm, self.nz))
        self.L.upco_fill_cs_zm(self.h, flag, _ptr(out))
        return out

    # photon pt
    def photon_pt_cdf(self, e):
        cdf = np.zeros(5001)
        self.L.upco_photon_pt_cdf(self.h, float(e), _ptr(cdf))
        return cdf

    def photon_pt_sample(self, cdf, r):
        return self.L.upco_photon_pt_sample(self.h, _ptr(cdf), float(r))

    def generate_event(self, seed, cand, cs_sum, z_sum, z_sum_ps=None, ratio=None):
        npart = C.c_int()
        pdg = np.zeros(6, np.int32); st = np.zeros(6, np.int32); mo = np.zeros(6, np.int32)
        p4 = np.zeros((6, 4)); aux = np.zeros(5)
        acc = self.L.upco_generate_event(self.h, seed, cand, _ptr(cs_sum), _ptr(z_sum), _ptr(z_sum_ps),
                                         _ptr(ratio), C.byref(npart), _ptr(pdg), _ptr(st), _ptr(mo),
                                         _ptr(p4), _ptr(aux))
        n = npart.value
        return acc, pdg[:n].copy(), st[:n].copy(), mo[:n].copy(), p4[:n].copy(), aux


    def generate_event_u(self, u, cs_sum, z_sum, z_sum_ps=None, ratio=None):
        """generateEvent with the twelve uniforms of the slot map injected (see upco_generate_event_u)."""
        u = np.ascontiguousarray(u, float)
        assert u.size in (12, 14)
        u = np.concatenate([u, np.zeros(14 - u.size)])
        npart = C.c_int()
        pdg = np.zeros(6, np.int32); st = np.zeros(6, np.int32); mo = np.zeros(6, np.int32)
        p4 = np.zeros((6, 4)); aux = np.zeros(5)
        acc = self.L.upco_generate_event_u(self.h, _ptr(u), _ptr(cs_sum), _ptr(z_sum), _ptr(z_sum_ps), _ptr(ratio),
                                           C.byref(npart), _ptr(pdg), _ptr(st), _ptr(mo), _ptr(p4), _ptr(aux))
        n = npart.value
        return acc, pdg[:n].copy(), st[:n].copy(), mo[:n].copy(), p4[:n].copy(), aux


# free functions -----------------------------------------------------------------------------
def cspline_init(x, y):
    x = np.ascontiguousarray(x, float); y = np.ascontiguousarray(y, float)
    c = np.zeros_like(x)
    lib().upco_cspline_init(_ptr(x), _ptr(y), x.size, _ptr(c))
    return c


def cspline_eval(x, y, c, xv):
    x = np.ascontiguousarray(x, float); y = np.ascontiguousarray(y, float); c = np.ascontiguousarray(c, float)
    return np.array([lib().upco_cspline_eval(_ptr(x), _ptr(y), _ptr(c), x.size, float(v)) for v in np.atleast_1d(xv)])


def qags_test(kind, alpha, a, b, epsabs, epsrel):
    err = C.c_double(); ne = C.c_int(); last = C.c_int(); ier = C.c_int()
    r = lib().upco_qags_test(kind, alpha, a, b, epsabs, epsrel, C.byref(err), C.byref(ne), C.byref(last), C.byref(ier))
    return r, err.value, ne.value, last.value, ier.value


def pdf_init(bins):
    bins = np.ascontiguousarray(bins, float).ravel()
    s = np.zeros(bins.size + 1)
    lib().upco_pdf_init(_ptr(bins), bins.size, _ptr(s))
    return s


def pdf_find(s, r):
    return lib().upco_pdf_find(_ptr(s), s.size - 1, float(r))


def sample2d(s, xe, ye, r1, r2):
    k = C.c_longlong(); x = C.c_double(); y = C.c_double()
    lib().upco_sample2d(_ptr(s), xe.size - 1, ye.size - 1, _ptr(xe), _ptr(ye), float(r1), float(r2),
                        C.byref(k), C.byref(x), C.byref(y))
    return k.value, x.value, y.value


def sample1d(s, edges, r):
    return lib().upco_sample1d(_ptr(s), edges.size - 1, _ptr(edges), float(r))


def get_bin(nbins, x, lo, hi):
    return lib().upco_get_bin(int(nbins), float(x), float(lo), float(hi))


def philox(seed, ctr, block):
    a = C.c_double(); b = C.c_double()
    lib().upco_philox(seed, ctr, block, C.byref(a), C.byref(b))
    return a.value, b.value


def bessel(name, x):
    f = getattr(lib(), {"K0": "upco_bessel_K0", "K1": "upco_bessel_K1", "J1": "upco_bessel_J1",
                        "TK1": "upco_tmath_besselK1", "TI1": "upco_tmath_besselI1"}[name])
    return np.array([f(float(v)) for v in np.atleast_1d(x)])
