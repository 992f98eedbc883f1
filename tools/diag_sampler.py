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
    print(f"{name:24s} {best * 1e3:7.3f} ms")
g.close()
