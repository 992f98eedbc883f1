#!/usr/bin/env python
"""A small pass over every stage (tables, form-factor fill with a starved hand-over pool, polarised fill, fold, sampler
build, events with photon pT, photon flux) -- the command compute-sanitizer wraps:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py cells
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"

if what in ("all", "ff"):
    P = named_config("cfg2", "BINS_M 24\nBINS_Y 8\n")
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    t = g.fill_lumi()
    cs, _, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
    g.sampler_build(cszm=capi.elem_cs_zm(P))
    ev = g.generate(5, 0, 3000)
    print("ff:", float(t.sum()), tot, ev["n_accepted"], g.fill_stats()["qags_evals"])
    print("flux1d:", g.photon_flux(np.array([3.1, 9.4]), np.array([0.5, -2.0])))
    g.close()
if what in ("all", "cells", "pol"):
    P = named_config("cfg1", "PROC_ID 11\nUSE_POLARIZED_CS 1\nBINS_M 20\nBINS_Y 9\nMMIN 1\nMMAX 20\n")
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    s, p = g.fill_lumi()
    cs, ratio, tot = g.fold_sigma(sig_s=capi.elem_sigma_m(P, 1), sig_p=capi.elem_sigma_m(P, 2))
    g.sampler_build(cszm_s=capi.elem_cs_zm(P, 1), cszm_ps=capi.elem_cs_zm(P, 2))
    ev = g.generate(7, 0, 3000)
    print("pol:", float(s.sum()), float(p.sum()), tot, ev["n_accepted"])
    g.close()
if what in ("all", "cells", "alp"):
    P = named_config("cfg5", "BINS_M 20\nBINS_Y 10\n")
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    t = g.fill_lumi()
    cs, _, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
    g.sampler_build(cszm=None)
    ev = g.generate(9, 0, 20000)
    print("alp:", float(t.sum()), tot, ev["n_accepted"])
    g.close()
if what in ("all", "pdf"):
    # the speculate-and-verify CDF build (tables beyond 4096 bins): a smooth table, one with a sign change of regime
    P = named_config("cfg1", "BINS_M 16\nBINS_Y 8\n")
    g = capi.UpcGpu(P, 0)
    rng = np.random.default_rng(3)
    i = np.arange(40000)
    for t in (np.exp(-((i / 40000.0 - 0.4) ** 2) * 20) * 1e-3, np.exp(rng.normal(0, 9, 23000)), (i[:9000] + 1.0) * 1e-3):
        s2 = g.hist_pdf_init(t)
        print("pdf:", t.size, float(s2[-1]), g.sampler_spec_stats()["blocks"])
    g.close()
