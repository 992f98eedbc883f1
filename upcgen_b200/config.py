"""parameters.in semantics and the generator's parameter block (host logic, no GPU).

Mirrors the reference's configuration path:
  * `UpcGenerator::setParameterValue`        src/UpcGenerator.cpp:183-305  (KEY value pairs)
  * `UpcGenerator::configGeneratorFromFile`  src/UpcGenerator.cpp:318-346  ('#' comments, order-free,
                                              unknown keys ignored, missing file -> defaults)
  * `UpcGenerator::init` grid overrides      src/UpcGenerator.cpp:69-140
  * defaults                                 include/UpcCrossSection.h:73-164, include/UpcGenerator.h
The resulting `UpcParams` is what crosses the C-ABI (include/upcgpu.h: `upcgpu_params`).
"""
from __future__ import annotations

import dataclasses
import math
import os

# include/UpcPhysConstants.h:26-32
ALPHA = 1.0 / 137.035999074
HC = 0.1973269718
M_PROT = 0.9382720813
M_NEUT = 0.939565346
M_EL = 0.000510998946
M_MU = 0.1056583745
M_TAU = 1.77686

LEPTON_MASS = {11: M_EL, 13: M_MU, 15: M_TAU}


@dataclasses.dataclass
class UpcParams:
    # nucleus / beam (UpcCrossSection.h:73-85)
    Z: int = 82
    A: int = 208
    R: float = 6.68
    a: float = 0.447
    sqrts: float = 5020.0
    g1: float = 5020.0 / (2.0 * M_PROT)
    g2: float = 5020.0 / (2.0 * M_PROT)
    # Q1: gtot is fixed in the UpcCrossSection constructor from the DEFAULT sqrts
    # (src/UpcCrossSection.cpp:51-54), before parameters.in is read
    gtot: float = math.cosh((math.acosh(5020.0 / (2.0 * M_PROT)) + math.acosh(5020.0 / (2.0 * M_PROT))) / 2.0)
    is_point: int = 1
    breakup_mode: int = 1
    use_pol: int = 0          # nucProcessCS->usePolarizedCS (lumi + fold)
    nonzero_gam_pt: int = 1
    nm: int = 1001
    ny: int = 121
    nz: int = 100
    mmin: float = 3.56
    mmax: float = 50.0
    ymin: float = -6.0
    ymax: float = 6.0
    zmin: float = -1.0
    zmax: float = 1.0
    nb1: int = 120
    nb2: int = 120
    proc_id: int = 11         # uninitialised in the reference (UpcGenerator.h:64); parameters.in sets it
    a_lep: float = 0.0
    alp_mass: float = 1.0
    alp_width: float = 0.010
    do_pt_cut: int = 0
    do_eta_cut: int = 0
    pt_min: float = 0.0
    eta_min: float = 0.0
    eta_max: float = 0.0
    # generator-level fields that do not cross into the table kernels
    gen_use_pol: int = 0      # UpcGenerator::usePolarizedCS (may be cleared by init(), Q5)
    n_events: int = 1000
    seed: int = 0
    use_root_out: int = 1
    use_hepmc_out: int = 0
    pythia_version: int = -1
    do_fsr: int = 0
    do_decays: int = 0
    do_mass_cut: int = 0
    low_m_cut: float = 0.0
    hi_m_cut: float = 9999.0
    shadowing: int = 0
    decay_pdg: int = 0

    # ------------------------------------------------------------------------------------
    def set_parameter(self, key: str, val: str) -> None:
        """`UpcGenerator::setParameterValue`, src/UpcGenerator.cpp:183-305."""
        f, i = float, (lambda s: int(float(s)) if ("." in s or "e" in s.lower()) else int(s))
        if key == "NEVENTS": self.n_events = i(val)
        elif key == "SQRTS":
            self.sqrts = f(val)
            self.g1 = self.sqrts / (2.0 * M_PROT)
            self.g2 = self.sqrts / (2.0 * M_PROT)
        elif key == "PROC_ID": self.proc_id = i(val)
        elif key == "LEP_A": self.a_lep = f(val)
        elif key == "ALP_MASS": self.alp_mass = f(val)
        elif key == "ALP_WIDTH": self.alp_width = f(val)
        elif key == "DO_PT_CUT": self.do_pt_cut = i(val)
        elif key == "PT_MIN": self.pt_min = f(val)
        elif key == "DO_ETA_CUT": self.do_eta_cut = i(val)
        elif key == "ETA_MIN": self.eta_min = f(val)
        elif key == "ETA_MAX": self.eta_max = f(val)
        elif key == "ZMIN": self.zmin = f(val)
        elif key == "ZMAX": self.zmax = f(val)
        elif key == "MMIN": self.mmin = f(val)
        elif key == "MMAX": self.mmax = f(val)
        elif key == "YMIN": self.ymin = f(val)
        elif key == "YMAX": self.ymax = f(val)
        elif key == "BINS_Z": self.nz = i(val)
        elif key == "BINS_M": self.nm = i(val)
        elif key == "BINS_Y": self.ny = i(val)
        elif key == "WS_R": self.R = f(val)
        elif key == "WS_A": self.a = f(val)
        elif key == "NUCLEUS_Z": self.Z = i(val)
        elif key == "NUCLEUS_A": self.A = i(val)
        elif key == "FLUX_POINT": self.is_point = 1 if i(val) else 0
        elif key == "BREAKUP_MODE": self.breakup_mode = i(val)
        elif key == "PYTHIA_VERSION": self.pythia_version = i(val)
        elif key == "PYTHIA8_FSR": self.do_fsr = i(val)
        elif key == "PYTHIA8_DECAYS": self.do_decays = i(val)
        elif key == "NON_ZERO_GAM_PT": self.nonzero_gam_pt = 1 if i(val) else 0
        elif key == "USE_POLARIZED_CS":
            self.use_pol = 1 if i(val) else 0
            self.gen_use_pol = 1 if i(val) else 0
        elif key == "SEED": self.seed = i(val)
        elif key == "USE_ROOT_OUTPUT": self.use_root_out = i(val)
        elif key == "USE_HEPMC_OUTPUT": self.use_hepmc_out = i(val)
        elif key == "DO_M_CUT": self.do_mass_cut = i(val)
        elif key == "LOW_M_CUT": self.low_m_cut = f(val)
        elif key == "HIGH_M_CUT": self.hi_m_cut = f(val)
        elif key == "SHADOWING": self.shadowing = i(val)
        elif key == "DECAY_PDG": self.decay_pdg = i(val)
        # unknown keys are silently ignored, like the reference

    @classmethod
    def from_text(cls, text: str) -> "UpcParams":
        """`configGeneratorFromFile`: lines starting with '#' skipped, first two whitespace
        separated tokens are KEY VALUE; a line with fewer tokens re-applies the previous pair
        (the reference does not reset its variables, src/UpcGenerator.cpp:325-336)."""
        p = cls()
        key, val = "", ""
        for line in text.splitlines():
            if line[:1] == "#":
                continue
            toks = line.split()
            if len(toks) >= 1:
                key = toks[0]
            if len(toks) >= 2:
                val = toks[1]
            if key:
                try:
                    p.set_parameter(key, val)
                except ValueError:
                    raise ValueError(f"parameters.in: bad value {val!r} for {key}")
        return p

    @classmethod
    def from_file(cls, path: str) -> "UpcParams":
        if not os.path.exists(path):
            return cls()  # "Input file not found! Using default parameters..."
        with open(path) as fh:
            return cls.from_text(fh.read())

    def init(self) -> "UpcParams":
        """Grid/flag overrides of `UpcGenerator::init`, src/UpcGenerator.cpp:69-140."""
        if self.proc_id == 22:  # light-by-light: fixed grid, generator-level pol flag cleared (Q5)
            self.zmin, self.zmax, self.nz = -0.99, 0.99, 198
            self.mmin, self.mmax, self.nm = 0.05, 50.0, 1000
            self.gen_use_pol = 0
        if self.proc_id == 111:
            self.zmin, self.zmax, self.nz = -1.0, 1.0, 100
            self.mmin, self.mmax, self.nm = 0.275, 5.0, 91
            self.gen_use_pol = 0
        if self.proc_id == 51:
            self.mmin = self.alp_mass - 4.0 * self.alp_width
            self.mmax = self.alp_mass + 4.0 * self.alp_width
        if 11 <= self.proc_id <= 15:
            m_part = LEPTON_MASS.get(self.proc_id, 0.0)
            if self.mmin < m_part * 2.0:
                self.mmin = m_part * 2.0
        return self

    # derived ----------------------------------------------------------------------------
    @property
    def dm(self): return (self.mmax - self.mmin) / self.nm
    @property
    def dy(self): return (self.ymax - self.ymin) / self.ny
    @property
    def dz(self): return (self.zmax - self.zmin) / self.nz
    @property
    def ignore_csz(self): return self.proc_id == 51
    @property
    def n_cells(self): return self.nm * self.ny


# ----------------------------------------------------------------------------------------
# BASELINE.json configs.  parameters.in text, so that the same
# parser path is exercised.
_BASE = """NUCLEUS_Z 82
NUCLEUS_A 208
WS_R 6.68
WS_A 0.447
SQRTS 5020
PROC_ID 15
LEP_A 0
NEVENTS 1000
DO_PT_CUT 0
PT_MIN 0
DO_ETA_CUT 0
ETA_MIN -1.0
ETA_MAX 1.0
ZMIN -1
ZMAX 1
MMIN 3.56
MMAX 50
YMIN -6
YMAX 6
BINS_Z 100
BINS_M 1001
BINS_Y 121
FLUX_POINT 1
BREAKUP_MODE 1
NON_ZERO_GAM_PT 1
USE_POLARIZED_CS 0
PYTHIA_VERSION 8
PYTHIA8_FSR 1
PYTHIA8_DECAYS 0
SEED 12345
USE_ROOT_OUTPUT 1
USE_HEPMC_OUTPUT 0
SHADOWING 4
DECAY_PDG 13
"""

CONFIG_OVERRIDES = {
    # cfg1: the repo's parameters.in (ditau, point flux, no breakup), SEED fixed
    "cfg1": "",
    # cfg2: dimuon, Woods-Saxon form-factor flux, XNXN, 1e6 events  (the bench workload)
    "cfg2": "PROC_ID 13\nFLUX_POINT 0\nBREAKUP_MODE 2\nNEVENTS 1000000\n",
    # cfg3: light-by-light grid with polarised lumi tables, 0NXN (lumi-level only, Q5)
    "cfg3": "PROC_ID 22\nUSE_POLARIZED_CS 1\nBREAKUP_MODE 4\n",
    # cfg4: dielectron fine grid, form-factor flux, 8 GPUs
    "cfg4": "PROC_ID 11\nMMIN 1\nMMAX 100\nBINS_M 10001\nBINS_Y 1201\nFLUX_POINT 0\n",
    # cfg5: Xe-Xe 5.44 TeV ALP, photon pT, 0N0N, 1e7 events.  WS_R/WS_A for 129Xe are not in the
    # reference; fixed here to 5.36 / 0.59 fm.
    "cfg5": ("NUCLEUS_Z 54\nNUCLEUS_A 129\nSQRTS 5440\nWS_R 5.36\nWS_A 0.59\nPROC_ID 51\n"
             "ALP_MASS 1\nALP_WIDTH 0.01\nNON_ZERO_GAM_PT 1\nBREAKUP_MODE 3\nNEVENTS 10000000\n"),
}


def config_text(name: str, extra: str = "") -> str:
    """The parameters.in text of cfg1..cfg5 (+ extra lines): what `named_config` parses."""
    return _BASE + CONFIG_OVERRIDES[name] + extra


def named_config(name: str, extra: str = "") -> UpcParams:
    """One of cfg1..cfg5, optionally with extra 'KEY value' lines appended (later wins)."""
    return UpcParams.from_text(_BASE + CONFIG_OVERRIDES[name] + extra).init()
