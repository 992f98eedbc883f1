"""ctypes binding of libupcgpu.so (include/upcgpu.h) used by the tests, bench.py and the
torch.distributed launcher.  This is plumbing: every numeric result comes from the CUDA library;
there is no Python/NumPy fallback, and loading fails loudly if the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .config import LEPTON_MASS, UpcParams

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libupcgpu.so")
_LIB = None

OK, EINVAL, ECUDA, ENODEV, EQAGS, ERANGE = 0, -1, -2, -3, -4, -5
TABLE_GAA, TABLE_TA, TABLE_FORMFAC, TABLE_BREAKUP = 0, 1, 2, 3
MAX_PART = 6   # UPCGPU_MAX_PART (include/upcgpu.h)


class UpcGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"upcgpu error {code}: {msg}")
        self.code = code


class CParams(C.Structure):
    _fields_ = [
        ("Z", C.c_int), ("A", C.c_int), ("R", C.c_double), ("a", C.c_double),
        ("sqrts", C.c_double), ("g1", C.c_double), ("g2", C.c_double), ("gtot", C.c_double),
        ("is_point", C.c_int), ("breakup_mode", C.c_int), ("use_pol", C.c_int), ("nonzero_gam_pt", C.c_int),
        ("nm", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("mmin", C.c_double), ("mmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("zmin", C.c_double), ("zmax", C.c_double),
        ("nb1", C.c_int), ("nb2", C.c_int),
        ("part_pdg", C.c_int), ("m_part", C.c_double), ("is_charged", C.c_int), ("is_pair", C.c_int),
        ("is_single", C.c_int), ("ignore_csz", C.c_int), ("decay_uniform_pdg", C.c_int),
        ("do_pt_cut", C.c_int), ("do_eta_cut", C.c_int),
        ("pt_min", C.c_double), ("eta_min", C.c_double), ("eta_max", C.c_double),
    ]


class CTableInfo(C.Structure):
    _fields_ = [("rho0", C.c_double), ("sigma_nn", C.c_double), ("factor", C.c_double),
                ("breakup_p20", C.c_double), ("n_breakup_energy_knots", C.c_int),
                ("gaa_zero_below", C.c_double)]


class CFillStats(C.Structure):
    _fields_ = [("qags_integrals", C.c_longlong), ("qags_evals", C.c_longlong), ("qags_overflow", C.c_longlong),
                ("qags_errors", C.c_longlong), ("flux_rows", C.c_longlong), ("band_pairs", C.c_longlong),
                ("ms_tables", C.c_double), ("ms_flux", C.c_double), ("ms_cells", C.c_double), ("ms_total", C.c_double),
                ("ms_qags", C.c_double), ("ms_qags_head", C.c_double), ("qags_head_evals", C.c_longlong),
                ("qags_head_done", C.c_longlong), ("qags_table_evals", C.c_longlong), ("cells_evaluated", C.c_longlong)]


# every symbol include/upcgpu.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "upcgpu_create", "upcgpu_destroy", "upcgpu_last_error", "upcgpu_abi_version", "upcgpu_device_name",
    "upcgpu_prepare_tables", "upcgpu_get_table_info", "upcgpu_get_table", "upcgpu_eval_table", "upcgpu_breakup_raw",
    "upcgpu_flux_point", "upcgpu_flux_form", "upcgpu_fill_lumi", "upcgpu_fill_lumi_shard", "upcgpu_lumi_cells",
    "upcgpu_get_fill_stats", "upcgpu_lumi_shard_buffer", "upcgpu_lumi_gather_buffer", "upcgpu_lumi_unpack",
    "upcgpu_lumi_download", "upcgpu_lumi_upload", "upcgpu_fold_sigma", "upcgpu_sampler_build",
    "upcgpu_sampler_get_cdf", "upcgpu_sampler_spec_stats", "upcgpu_sample_ym", "upcgpu_sample_z", "upcgpu_generate", "upcgpu_generate_packed", "upcgpu_particles_per_event",
    "upcgpu_generate_device",
    "upcgpu_photon_pt_cdf", "upcgpu_philox", "upcgpu_invalidate_tables", "upcgpu_fp64_peak",
    "upcgpu_stream_handle", "upcgpu_launch_count", "upcgpu_elem_sigma_m", "upcgpu_elem_fill_cs_zm",
    "upcgpu_hist_pdf_init", "upcgpu_hist_sample2d", "upcgpu_hist_sample1d", "upcgpu_root_hist_read",
    "upcgpu_create_multi", "upcgpu_group_size", "upcgpu_group_member", "upcgpu_group_set_exchange",
    "upcgpu_group_describe", "upcgpu_root_write_th2d", "upcgpu_root_write_th1d", "upcgpu_root_write_tree", "upcgpu_root_write_sigma_hists", "upcgpu_root_set_compression", "upcgpu_photon_flux",
    "upcgpu_lumi_ipc_export", "upcgpu_lumi_ipc_import", "upcgpu_fill_lumi_shard_peers",
]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `python -m upcgen_b200.build` "
                              "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        p, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
        L.upcgpu_create.argtypes = [C.POINTER(CParams), i, C.POINTER(p)]
        L.upcgpu_destroy.argtypes = [p]
        L.upcgpu_destroy.restype = None
        L.upcgpu_last_error.argtypes = [p]
        L.upcgpu_last_error.restype = C.c_char_p
        L.upcgpu_device_name.argtypes = [p, C.c_char_p, sz]
        L.upcgpu_prepare_tables.argtypes = [p]
        L.upcgpu_get_table_info.argtypes = [p, C.POINTER(CTableInfo)]
        L.upcgpu_get_table.argtypes = [p, i, sz, sz, p, p, p]
        L.upcgpu_eval_table.argtypes = [p, i, p, sz, p]
        L.upcgpu_breakup_raw.argtypes = [p, p, i, sz, p]
        L.upcgpu_flux_point.argtypes = [p, p, p, sz, p]
        L.upcgpu_flux_form.argtypes = [p, p, p, sz, p, p]
        L.upcgpu_fill_lumi.argtypes = [p, p, p, p]
        L.upcgpu_fill_lumi_shard.argtypes = [p, i, i]
        L.upcgpu_lumi_cells.argtypes = [p, p, p, sz, p, p, p]
        L.upcgpu_get_fill_stats.argtypes = [p, C.POINTER(CFillStats)]
        L.upcgpu_lumi_shard_buffer.argtypes = [p, i, C.POINTER(C.c_uint64), C.POINTER(sz)]
        L.upcgpu_lumi_gather_buffer.argtypes = [p, i, i, C.POINTER(C.c_uint64), C.POINTER(sz)]
        L.upcgpu_lumi_unpack.argtypes = [p, i]
        L.upcgpu_lumi_download.argtypes = [p, i, p]
        L.upcgpu_lumi_upload.argtypes = [p, i, p]
        L.upcgpu_fold_sigma.argtypes = [p, p, p, p, p, p, C.POINTER(d)]
        L.upcgpu_sampler_build.argtypes = [p, p, p, p, p]
        L.upcgpu_sampler_get_cdf.argtypes = [p, p, p, p]
        L.upcgpu_sample_ym.argtypes = [p, p, sz, p, p, p, p, p]
        L.upcgpu_sample_z.argtypes = [p, p, p, sz, i, p]
        L.upcgpu_generate.argtypes = [p, C.c_uint64, C.c_uint64, sz, p, p, p, p, p, p, C.POINTER(C.c_uint64)]
        L.upcgpu_generate_device.argtypes = [p, C.c_uint64, C.c_uint64, sz, C.POINTER(C.c_uint64)]
        L.upcgpu_generate_packed.argtypes = [p, C.c_uint64, C.c_uint64, sz, C.c_int, p, p, p, p, p, p, C.POINTER(C.c_uint64)]
        L.upcgpu_particles_per_event.argtypes = [p]
        L.upcgpu_photon_pt_cdf.argtypes = [p, d, p]
        L.upcgpu_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, sz, p]
        L.upcgpu_invalidate_tables.argtypes = [p]
        L.upcgpu_hist_pdf_init.argtypes = [p, p, sz, p]
        L.upcgpu_hist_sample2d.argtypes = [p, p, i, i, p, p, p, sz, p, p, p]
        L.upcgpu_hist_sample1d.argtypes = [p, p, i, p, p, sz, p]
        L.upcgpu_elem_sigma_m.argtypes = [i, d, d, d, i, p, sz, p]
        L.upcgpu_elem_fill_cs_zm.argtypes = [i, d, d, d, i, d, d, i, d, d, i, p]
        L.upcgpu_stream_handle.argtypes = [p, C.POINTER(C.c_uint64)]
        L.upcgpu_launch_count.argtypes = [p]
        L.upcgpu_launch_count.restype = C.c_longlong
        L.upcgpu_fp64_peak.argtypes = [p, i, C.POINTER(d), C.POINTER(d)]
        L.upcgpu_create_multi.argtypes = [C.POINTER(CParams), i, C.POINTER(i), C.POINTER(p)]
        L.upcgpu_group_size.argtypes = [p]
        L.upcgpu_group_member.argtypes = [p, i, C.POINTER(p)]
        L.upcgpu_group_set_exchange.argtypes = [p, i]
        L.upcgpu_group_describe.argtypes = [p, C.c_char_p, sz]
        _LIB = L
    return _LIB


def to_cparams(P: UpcParams) -> CParams:
    """UpcParams -> upcgpu_params, incl. the elementary-process members set by
    UpcCrossSection::setElemProcess / UpcGenerator::init (src/UpcGenerator.cpp:69-140)."""
    c = CParams()
    for name, ctype in CParams._fields_:
        if hasattr(P, name):
            v = getattr(P, name)
            setattr(c, name, int(v) if ctype is C.c_int else float(v))
    pid = P.proc_id
    if pid in (11, 13, 15):
        c.part_pdg, c.m_part, c.is_charged, c.is_pair, c.is_single = pid, LEPTON_MASS[pid], 1, 1, 0
        c.ignore_csz, c.decay_uniform_pdg = 0, 0
    elif pid == 51:
        c.part_pdg, c.m_part, c.is_charged, c.is_pair, c.is_single = 51, P.alp_mass, 0, 0, 1
        c.ignore_csz, c.decay_uniform_pdg = 1, 22
    elif pid == 22:
        c.part_pdg, c.m_part, c.is_charged, c.is_pair, c.is_single = 22, 0.0, 0, 1, 0
        c.ignore_csz, c.decay_uniform_pdg = 0, 0
    elif pid == 111:  # pi0 pi0, each decaying uniformly into two photons (src/UpcTwoPhotonDipion.cpp:37-39, UpcGenerator.cpp:799-803)
        c.part_pdg, c.m_part, c.is_charged, c.is_pair, c.is_single = 111, 0.1349770, 0, 1, 0
        c.ignore_csz, c.decay_uniform_pdg = 0, 22
    else:
        raise ValueError(f"PROC_ID {pid} is outside the GPU path (SURVEY.md section 2)")
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class UpcGpu:
    """One context = one GPU.  Thin, typed wrapper; methods map 1:1 onto include/upcgpu.h."""

    def __init__(self, P: UpcParams, device: int = 0, n_gpus: int = 1, devices=None, _borrowed=None):
        """n_gpus > 1: several devices behind this one handle (upcgpu_create_multi), driven from this process."""
        self.L = lib()
        self.P = P
        self.cp = to_cparams(P)
        self.nm, self.ny, self.nz = P.nm, P.ny, P.nz
        self._owned = _borrowed is None
        if _borrowed is not None:
            self.h = _borrowed
            return
        h = C.c_void_p()
        if n_gpus > 1 or devices is not None:
            devs = list(devices) if devices is not None else list(range(n_gpus))
            arr = (C.c_int * len(devs))(*devs)
            rc = self.L.upcgpu_create_multi(C.byref(self.cp), len(devs), arr, C.byref(h))
        else:
            rc = self.L.upcgpu_create(C.byref(self.cp), device, C.byref(h))
        if rc != OK:
            raise UpcGpuError(rc, self.L.upcgpu_last_error(None).decode())
        self.h = h

    def group_size(self):
        return self.L.upcgpu_group_size(self.h)

    def group_member(self, rank):
        """The member context of `rank` (borrowed: owned by this handle)."""
        m = C.c_void_p()
        self._chk(self.L.upcgpu_group_member(self.h, rank, C.byref(m)))
        return UpcGpu(self.P, _borrowed=m)

    def group_set_exchange(self, mode):
        self._chk(self.L.upcgpu_group_set_exchange(self.h, mode))

    def group_describe(self):
        buf = C.create_string_buffer(256)
        self._chk(self.L.upcgpu_group_describe(self.h, buf, 256))
        return buf.value.decode()

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_owned", True):
                self.L.upcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != OK:
            raise UpcGpuError(rc, self.L.upcgpu_last_error(self.h).decode())

    def device_name(self):
        buf = C.create_string_buffer(256)
        self._chk(self.L.upcgpu_device_name(self.h, buf, 256))
        return buf.value.decode()

    # tables ------------------------------------------------------------------------------
    def prepare_tables(self):
        self._chk(self.L.upcgpu_prepare_tables(self.h))
        return self.table_info()

    def stream_handle(self):
        v = C.c_uint64()
        self._chk(self.L.upcgpu_stream_handle(self.h, C.byref(v)))
        return v.value

    def launch_count(self):
        return self.L.upcgpu_launch_count(self.h)

    def invalidate_tables(self):
        self._chk(self.L.upcgpu_invalidate_tables(self.h))

    def fp64_peak(self, iters=200000):
        tf, ms = C.c_double(), C.c_double()
        self._chk(self.L.upcgpu_fp64_peak(self.h, iters, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    def table_info(self):
        info = CTableInfo()
        self._chk(self.L.upcgpu_get_table_info(self.h, C.byref(info)))
        return info

    def get_table(self, which, i0, n):
        x, y, c = np.zeros(n), np.zeros(n), np.zeros(n)
        self._chk(self.L.upcgpu_get_table(self.h, which, i0, n, _p(x), _p(y), _p(c)))
        return x, y, c

    def eval_table(self, which, x):
        x = _f64(x).ravel()
        out = np.zeros_like(x)
        self._chk(self.L.upcgpu_eval_table(self.h, which, _p(x), x.size, _p(out)))
        return out

    def breakup_raw(self, b, mode):
        b = _f64(b).ravel()
        out = np.zeros_like(b)
        self._chk(self.L.upcgpu_breakup_raw(self.h, _p(b), mode, b.size, _p(out)))
        return out

    # fluxes ------------------------------------------------------------------------------
    def flux_point(self, b, k):
        b, k = np.broadcast_arrays(_f64(b), _f64(k))
        bb, kk = _f64(b.ravel()), _f64(k.ravel())
        out = np.zeros(bb.size)
        self._chk(self.L.upcgpu_flux_point(self.h, _p(bb), _p(kk), bb.size, _p(out)))
        return out.reshape(b.shape)

    def flux_form(self, b, k, with_neval=False):
        b, k = np.broadcast_arrays(_f64(b), _f64(k))
        bb, kk = _f64(b.ravel()), _f64(k.ravel())
        out = np.zeros(bb.size)
        ne = np.zeros(bb.size, dtype=np.int32)
        self._chk(self.L.upcgpu_flux_form(self.h, _p(bb), _p(kk), bb.size, _p(out), _p(ne)))
        if with_neval:
            return out.reshape(b.shape), ne.reshape(b.shape)
        return out.reshape(b.shape)

    # lumi --------------------------------------------------------------------------------
    def fill_lumi(self):
        """Whole grid through the host-buffer entry point (tables are prepared if needed)."""
        if self.P.use_pol:
            s, p = np.zeros((self.nm, self.ny)), np.zeros((self.nm, self.ny))
            self._chk(self.L.upcgpu_fill_lumi(self.h, None, _p(s), _p(p)))
            return s, p
        out = np.zeros((self.nm, self.ny))
        self._chk(self.L.upcgpu_fill_lumi(self.h, _p(out), None, None))
        return out

    def fill_lumi_shard(self, shard=0, nshards=1):
        self._chk(self.L.upcgpu_fill_lumi_shard(self.h, shard, nshards))

    def lumi_cells(self, M, Y):
        M, Y = np.broadcast_arrays(_f64(M), _f64(Y))
        mm, yy = _f64(M.ravel()), _f64(Y.ravel())
        if self.P.use_pol:
            s, p = np.zeros(mm.size), np.zeros(mm.size)
            self._chk(self.L.upcgpu_lumi_cells(self.h, _p(mm), _p(yy), mm.size, None, _p(s), _p(p)))
            return s.reshape(M.shape), p.reshape(M.shape)
        out = np.zeros(mm.size)
        self._chk(self.L.upcgpu_lumi_cells(self.h, _p(mm), _p(yy), mm.size, _p(out), None, None))
        return out.reshape(M.shape)

    def photon_flux(self, M, Y):
        """calcPhotonFlux(M, +Y) and calcPhotonFlux(M, -Y) (src/UpcCrossSection.cpp:700-722)."""
        M, Y = np.broadcast_arrays(_f64(M), _f64(Y))
        mm, yy = _f64(M.ravel()), _f64(Y.ravel())
        fp, fn = np.zeros(mm.size), np.zeros(mm.size)
        self.L.upcgpu_photon_flux.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        self._chk(self.L.upcgpu_photon_flux(self.h, _p(mm), _p(yy), mm.size, _p(fp), _p(fn)))
        return fp.reshape(M.shape), fn.reshape(M.shape)

    def fill_stats(self):
        st = CFillStats()
        self._chk(self.L.upcgpu_get_fill_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in CFillStats._fields_}

    def lumi_shard_buffer(self, which):
        ptr, n = C.c_uint64(), C.c_size_t()
        self._chk(self.L.upcgpu_lumi_shard_buffer(self.h, which, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def lumi_gather_buffer(self, which, nshards):
        ptr, n = C.c_uint64(), C.c_size_t()
        self._chk(self.L.upcgpu_lumi_gather_buffer(self.h, which, nshards, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def lumi_ipc_export(self) -> bytes:
        """CUDA IPC handles of this context's full tables (3 x 64 bytes)."""
        buf = C.create_string_buffer(192)
        self.L.upcgpu_lumi_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
        self._chk(self.L.upcgpu_lumi_ipc_export(self.h, buf))
        return buf.raw

    def lumi_ipc_import(self, nshards, rank, handles: bytes):
        assert len(handles) == nshards * 192
        self.L.upcgpu_lumi_ipc_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        self._chk(self.L.upcgpu_lumi_ipc_import(self.h, nshards, rank, handles))

    def fill_lumi_shard_peers(self):
        self.L.upcgpu_fill_lumi_shard_peers.argtypes = [C.c_void_p]
        self._chk(self.L.upcgpu_fill_lumi_shard_peers(self.h))

    def lumi_unpack(self, nshards):
        self._chk(self.L.upcgpu_lumi_unpack(self.h, nshards))

    def lumi_download(self, which=0):
        out = np.zeros((self.nm, self.ny))
        self._chk(self.L.upcgpu_lumi_download(self.h, which, _p(out)))
        return out

    def lumi_upload(self, which, table):
        t = _f64(table)
        assert t.shape == (self.nm, self.ny)
        self._chk(self.L.upcgpu_lumi_upload(self.h, which, _p(t)))

    # fold / samplers ---------------------------------------------------------------------
    def fold_sigma(self, sig_m=None, sig_s=None, sig_p=None, download=True):
        sig_m = None if sig_m is None else _f64(sig_m)
        sig_s = None if sig_s is None else _f64(sig_s)
        sig_p = None if sig_p is None else _f64(sig_p)
        cs = np.zeros((self.ny, self.nm)) if download else None
        ratio = np.zeros((self.ny, self.nm)) if (download and self.P.use_pol) else None
        tot = C.c_double()
        self._chk(self.L.upcgpu_fold_sigma(self.h, _p(sig_m), _p(sig_s), _p(sig_p), _p(cs), _p(ratio), C.byref(tot)))
        return cs, ratio, tot.value

    def sampler_build(self, cs=None, cszm=None, cszm_s=None, cszm_ps=None):
        arrs = [None if a is None else _f64(a) for a in (cs, cszm, cszm_s, cszm_ps)]
        self._chk(self.L.upcgpu_sampler_build(self.h, *[_p(a) for a in arrs]))

    def sampler_spec_stats(self):
        out = (C.c_ulonglong * 6)()
        self.L.upcgpu_sampler_spec_stats.argtypes = [C.c_void_p, C.c_void_p]
        self._chk(self.L.upcgpu_sampler_spec_stats(self.h, out))
        return dict(blocks=int(out[0] + out[3]), fallback_blocks=int(out[1] + out[4]), rounds=int(out[2] + out[5]),
                    mean=dict(blocks=int(out[0]), fallback_blocks=int(out[1]), rounds=int(out[2])),
                    cumsum=dict(blocks=int(out[3]), fallback_blocks=int(out[4]), rounds=int(out[5])))

    def sampler_cdf(self):
        n = self.nm * self.ny
        s2 = np.zeros(n + 1)
        have_z = not self.cp.ignore_csz
        sz = np.zeros((self.nm, self.nz + 1)) if have_z else None
        sps = np.zeros((self.nm, self.nz + 1)) if (have_z and self.P.use_pol) else None
        self._chk(self.L.upcgpu_sampler_get_cdf(self.h, _p(s2), _p(sz), _p(sps)))
        return s2, sz, sps

    def sample_ym(self, u):
        u = _f64(u).reshape(-1, 2)
        n = u.shape[0]
        k = np.zeros(n, np.int64); yb = np.zeros(n, np.int32); mb = np.zeros(n, np.int32)
        y = np.zeros(n); m = np.zeros(n)
        self._chk(self.L.upcgpu_sample_ym(self.h, _p(u), n, _p(k), _p(yb), _p(mb), _p(y), _p(m)))
        return k, yb, mb, y, m

    def sample_z(self, mbin, u, ps=0):
        mbin = np.ascontiguousarray(mbin, np.int32); u = _f64(u).ravel()
        z = np.zeros(u.size)
        self._chk(self.L.upcgpu_sample_z(self.h, _p(mbin), _p(u), u.size, ps, _p(z)))
        return z

    # generic histogram samplers (UpcSampler1D/2D semantics) ----------------------------------
    def hist_pdf_init(self, bins):
        b = _f64(bins).ravel()
        s = np.zeros(b.size + 1)
        self._chk(self.L.upcgpu_hist_pdf_init(self.h, _p(b), b.size, _p(s)))
        return s

    def hist_sample2d(self, s, xe, ye, u):
        s, xe, ye = _f64(s), _f64(xe), _f64(ye)
        u = _f64(u).reshape(-1, 2)
        n = u.shape[0]
        k = np.zeros(n, np.int64); x = np.zeros(n); y = np.zeros(n)
        self._chk(self.L.upcgpu_hist_sample2d(self.h, _p(s), xe.size - 1, ye.size - 1, _p(xe), _p(ye), _p(u), n,
                                              _p(k), _p(x), _p(y)))
        return k, x, y

    def hist_sample1d(self, s, edges, u):
        s, edges, u = _f64(s), _f64(edges), _f64(u).ravel()
        x = np.zeros(u.size)
        self._chk(self.L.upcgpu_hist_sample1d(self.h, _p(s), edges.size - 1, _p(edges), _p(u), u.size, _p(x)))
        return x

    # events ------------------------------------------------------------------------------
    def generate(self, seed, first, n, with_aux=True):
        npart = np.zeros(n, np.int32)
        pdg = np.zeros((n, MAX_PART), np.int32); st = np.zeros((n, MAX_PART), np.int32); mo = np.zeros((n, MAX_PART), np.int32)
        p4 = np.zeros((n, MAX_PART, 4)); aux = np.zeros((n, 5)) if with_aux else None
        nacc = C.c_uint64()
        self._chk(self.L.upcgpu_generate(self.h, seed, first, n, _p(npart), _p(pdg), _p(st), _p(mo), _p(p4), _p(aux),
                                         C.byref(nacc)))
        return dict(npart=npart, pdg=pdg, status=st, mother=mo, p4=p4, aux=aux, n_accepted=nacc.value)

    def particles_per_event(self):
        return int(self.L.upcgpu_particles_per_event(self.h))

    def generate_packed(self, seed, first, n, part_stride=None, with_aux=True):
        """upcgpu_generate_packed: the particle arrays hold part_stride slots per candidate (default: the number of
        particles an event of this process has) instead of MAX_PART."""
        ps = self.particles_per_event() if part_stride is None else int(part_stride)
        npart = np.zeros(n, np.int32)
        pdg = np.zeros((n, ps), np.int32); st = np.zeros((n, ps), np.int32); mo = np.zeros((n, ps), np.int32)
        p4 = np.zeros((n, ps, 4)); aux = np.zeros((n, 5)) if with_aux else None
        nacc = C.c_uint64()
        self._chk(self.L.upcgpu_generate_packed(self.h, seed, first, n, ps, _p(npart), _p(pdg), _p(st), _p(mo), _p(p4),
                                                _p(aux), C.byref(nacc)))
        return dict(npart=npart, pdg=pdg, status=st, mother=mo, p4=p4, aux=aux, n_accepted=nacc.value)

    def generate_device(self, seed, first, n):
        nacc = C.c_uint64()
        self._chk(self.L.upcgpu_generate_device(self.h, seed, first, n, C.byref(nacc)))
        return nacc.value

    def photon_pt_cdf(self, e):
        cdf = np.zeros(5001)
        self._chk(self.L.upcgpu_photon_pt_cdf(self.h, float(e), _p(cdf)))
        return cdf


def elem_sigma_m(P: UpcParams, which=0, m=None):
    """sigma(m) of the built-in elementary process of P on the grid's lower mass edges (host plug-in)."""
    if m is None:
        m = P.mmin + P.dm * np.arange(P.nm)
    m = _f64(m)
    out = np.zeros_like(m)
    rc = lib().upcgpu_elem_sigma_m(P.proc_id, P.a_lep, P.alp_mass, P.alp_width, which, _p(m), m.size, _p(out))
    if rc != OK:
        raise UpcGpuError(rc, "elem_sigma_m: process not built in")
    return out


def elem_cs_zm(P: UpcParams, flag=0):
    """fillCrossSectionZM for the built-in process of P: [nm][nz]."""
    out = np.zeros((P.nm, P.nz))
    rc = lib().upcgpu_elem_fill_cs_zm(P.proc_id, P.a_lep, P.alp_mass, P.alp_width, flag, P.zmin, P.zmax, P.nz,
                                      P.mmin, P.mmax, P.nm, _p(out))
    if rc != OK:
        raise UpcGpuError(rc, "elem_cs_zm: process not built in")
    return out


def root_hist_read(path: str, name: str):
    """TH1D / TH2D `name` of the ROOT file `path`, read without ROOT (host/UpcRootHist.cpp).  Returns a dict with
    dim, the axes (nx, xlo, xhi[, ny, ylo, yhi]) and `cells`: [ny + 2][nx + 2] (TH2D) or [nx + 2] (TH1D), under- and
    overflow cells included, as ROOT stores them."""
    L = lib()
    L.upcgpu_root_hist_read.argtypes = [C.c_char_p, C.c_char_p] + [C.c_void_p] * 8 + [C.c_size_t, C.c_void_p]
    dim, nx, ny = C.c_int(), C.c_int(), C.c_int()
    xlo, xhi, ylo, yhi = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    n = C.c_size_t()
    args = [path.encode(), name.encode(), C.byref(dim), C.byref(nx), C.byref(xlo), C.byref(xhi), C.byref(ny), C.byref(ylo),
            C.byref(yhi)]
    if L.upcgpu_root_hist_read(*args, None, 0, C.byref(n)) != OK:
        raise UpcGpuError(EINVAL, f"root_hist_read: cannot read {name} from {path}")
    cells = np.zeros(n.value)
    if L.upcgpu_root_hist_read(*args, _p(cells), cells.size, C.byref(n)) != OK:
        raise UpcGpuError(EINVAL, f"root_hist_read: cannot read {name} from {path}")
    out = dict(dim=dim.value, nx=nx.value, xlo=xlo.value, xhi=xhi.value, ny=ny.value, ylo=ylo.value, yhi=yhi.value)
    out["cells"] = cells.reshape(ny.value + 2, nx.value + 2) if dim.value == 2 else cells
    return out


def root_write_th1d(path: str, name: str, values, xlo, xhi, cells=None, entries=None):
    """Writes one TH1D (bin i + 1 = values[i]; or all nx + 2 cells) into a ROOT file, without ROOT."""
    L = lib()
    if cells is None:
        v = _f64(values).ravel()
        cells = np.concatenate([[0.0], v, [0.0]])
    cells = _f64(cells).ravel()
    nx = cells.size - 2
    L.upcgpu_root_write_th1d.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_double]
    rc = L.upcgpu_root_write_th1d(path.encode(), name.encode(), nx, float(xlo), float(xhi), _p(cells),
                                  float(nx if entries is None else entries))
    if rc != OK:
        raise UpcGpuError(rc, "root_write_th1d failed")


def root_write_th2d(path: str, hists: dict, nx, xlo, xhi, ny, ylo, yhi):
    """Writes TH2D objects {name: table[nx][ny]} (bin (ix + 1, iy + 1) = table[ix][iy], as the reference fills
    hD2LDMDY: src/UpcCrossSection.cpp:564-571) into a ROOT file, without ROOT."""
    L = lib()
    names = list(hists)
    cells = []
    for nme in names:
        t = _f64(hists[nme])
        assert t.shape == (nx, ny)
        c = np.zeros((ny + 2, nx + 2))
        c[1:-1, 1:-1] = t.T
        cells.append(np.ascontiguousarray(c))
    arr_n = (C.c_char_p * len(names))(*[n.encode() for n in names])
    arr_c = (C.c_void_p * len(names))(*[c.ctypes.data for c in cells])
    L.upcgpu_root_write_th2d.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                         C.c_double, C.c_double, C.c_void_p]
    rc = L.upcgpu_root_write_th2d(path.encode(), len(names), arr_n, nx, xlo, xhi, ny, ylo, yhi, arr_c)
    if rc != OK:
        raise UpcGpuError(rc, f"root_write_th2d: cannot write {path}")


def root_write_tree(path: str, tree: str, title: str, columns: dict):
    """Writes a TTree of flat branches {name: (type 'I' | 'D', values)} into a ROOT file, without ROOT."""
    L = lib()
    names = list(columns)
    vals = [_f64(columns[n][1]) for n in names]
    types = "".join(columns[n][0] for n in names).encode()
    arr_n = (C.c_char_p * len(names))(*[n.encode() for n in names])
    arr_c = (C.c_void_p * len(names))(*[v.ctypes.data for v in vals])
    L.upcgpu_root_write_tree.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_char_p, C.c_void_p,
                                         C.c_size_t]
    rc = L.upcgpu_root_write_tree(path.encode(), tree.encode(), title.encode(), len(names), arr_n, types, arr_c,
                                  vals[0].size)
    if rc != OK:
        raise UpcGpuError(rc, f"root_write_tree: cannot write {path}")


def root_set_compression(setting: int) -> int:
    """ROOT compression setting of the files written through this library (0: none, 4xx: LZ4); returns the previous one."""
    L = lib()
    prev = C.c_int()
    L.upcgpu_root_set_compression.argtypes = [C.c_int, C.POINTER(C.c_int)]
    rc = L.upcgpu_root_set_compression(int(setting), C.byref(prev))
    if rc:
        raise UpcGpuError(rc, f"root_set_compression: setting {setting} is not supported (0, 1xx or 4xx)")
    return prev.value


def root_write_sigma_hists(path: str, y_edges, m_edges, cs):
    """hNucCSYM and its two projections as the reference writes them at debug level > 0 (src/UpcGenerator.cpp:900-917)."""
    L = lib()
    ye, me = _f64(y_edges), _f64(m_edges)
    cs = _f64(cs)
    assert cs.shape == (ye.size - 1, me.size - 1)
    L.upcgpu_root_write_sigma_hists.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rc = L.upcgpu_root_write_sigma_hists(path.encode(), ye.size - 1, _p(ye), me.size - 1, _p(me), _p(cs))
    if rc:
        raise UpcGpuError(rc, f"root_write_sigma_hists: cannot write {path}")


def philox(seed, ctr0, block, n):
    out = np.zeros((n, 2))
    rc = lib().upcgpu_philox(seed, ctr0, block, n, _p(out))
    if rc != OK:
        raise UpcGpuError(rc, "philox")
    return out
