#!/usr/bin/env python
"""tests/golden/lbyl_elem.npz: the light-by-light elementary cross sections on the grid the reference forces for
PROC_ID 22 (src/UpcGenerator.cpp:74-79), read from the reference's cross_sections/lbyl/*.root files through the
product's own reader (UpcRootHist via upcgpu_elem_sigma_m / upcgpu_elem_fill_cs_zm; tests/test_root_hist.py pins
that reader).  The data files themselves are not copied: sigma(m) in full (1000 values) and every 25th row of the
dsigma/dz table.

    UPCGEN_CROSS_SEC_DIR=/root/reference/cross_sections python tools/gen_lbyl_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("UPCGEN_CROSS_SEC_DIR", "/root/reference/cross_sections")
from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

P = named_config("cfg3", "USE_POLARIZED_CS 0\n")
sig = capi.elem_sigma_m(P, 0)
cszm = capi.elem_cs_zm(P, 0)
rows = np.arange(0, P.nm, 25)
out = os.path.join(ROOT, "tests", "golden", "lbyl_elem.npz")
np.savez_compressed(out, sig_m=sig, im_rows=rows, cszm_rows=cszm[rows], nm=P.nm, nz=P.nz, mmin=P.mmin, mmax=P.mmax,
                    zmin=P.zmin, zmax=P.zmax)
print(out, os.path.getsize(out), "bytes; sigma(m) sum", sig.sum(), "cszm rows", cszm[rows].shape)
