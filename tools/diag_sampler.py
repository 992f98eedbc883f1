#!/usr/bin/env python
"""Times upcgpu_sampler_build on tables of different content (wall clock, best of 5)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upcgen_b200 import capi
from upcgen_b200.config import named_config

P = named_config("cfg2")
g = capi.UpcGpu(P, 0)
g.prepare_tables()
g.fill_lumi()
cs, _, _ = g.fold_sigma(sig_m=capi.elem_sigma_m(P, 0))
cszm = capi.elem_cs_zm(P, 0)
rng = np.random.default_rng(1)
tables = {"cfg2 sigma table": cs, "uniform(1,2)": rng.uniform(1, 2, cs.shape), "ones": np.ones(cs.shape),
          "cfg2 table + 1e-30": cs + 1e-30}
print("cs: zeros", int((cs == 0).sum()), "of", cs.size, "min positive", cs[cs > 0].min(), "first", cs.ravel()[:4])
for name, t in tables.items():
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        g.sampler_build(cs=t, cszm=cszm)
        best = min(best, time.perf_counter() - t0)
    a = g.sampler_spec_stats()
    g.sampler_build(cs=t, cszm=cszm)
    b = g.sampler_spec_stats()
    print(f"{name:24s} {best * 1e3:7.3f} ms", {k: {q: b[k][q] - a[k][q] for q in b[k]} for k in ("mean", "cumsum")})
g.close()
# a table of cfg4's size (10001 x 1201 bins) through upcgpu_hist_pdf_init (includes the 96 MB up and down)
P4 = named_config("cfg1", "BINS_M 16\nBINS_Y 8\n")
g4 = capi.UpcGpu(P4, 0)
yy = np.exp(-np.linspace(-6, 6, 1201) ** 2 / 3)
mm = 1.0 / np.linspace(1, 100, 10001) ** 3
big = {"cfg4-size sigma-like": np.outer(yy, mm).ravel(), "cfg4-size ones": np.ones(1201 * 10001)}
for name, t in big.items():
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); g4.hist_pdf_init(t); best = min(best, time.perf_counter() - t0)
    a = g4.sampler_spec_stats(); g4.hist_pdf_init(t); b = g4.sampler_spec_stats()
    print(f"{name:24s} {best * 1e3:8.2f} ms", {k: {q: b[k][q] - a[k][q] for q in b[k]} for k in ("mean", "cumsum")})
g4.close()

# the real cfg4 (10001 x 1201 cells): fill, fold, then the sampler build on the device-resident sigma table
P4 = named_config("cfg4")
g = capi.UpcGpu(P4, 0)
g.prepare_tables()
g.fill_lumi_shard(0, 1)
g.fold_sigma(sig_m=capi.elem_sigma_m(P4, 0), download=False)
cszm4 = capi.elem_cs_zm(P4, 0)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); g.sampler_build(cszm=cszm4); best = min(best, time.perf_counter() - t0)
a = g.sampler_spec_stats(); g.sampler_build(cszm=cszm4); b = g.sampler_spec_stats()
print(f"the real cfg4 sampler build {best * 1e3:8.2f} ms", {k: {q: b[k][q] - a[k][q] for q in b[k]} for k in ("mean", "cumsum")})
g.close()
