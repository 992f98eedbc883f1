#!/usr/bin/env python
"""A small multi-device pass (upcgpu_create_multi, both exchanges) for compute-sanitizer on a box with >= 2 GPUs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

P = named_config("cfg2", "BINS_M 70\nBINS_Y 8\n")
one = capi.UpcGpu(P, 0)
one.prepare_tables()
ref = one.fill_lumi()
one.close()
for exchange in (1, 0):
    g = capi.UpcGpu(P, n_gpus=2)
    g.group_set_exchange(exchange)
    g.prepare_tables()
    t = g.fill_lumi()
    cs, _, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
    g.sampler_build(cszm=capi.elem_cs_zm(P))
    ev = g.generate(5, 0, 4000)
    print(g.group_describe(), "equal:", bool(np.array_equal(t, ref)), tot, ev["n_accepted"])
    g.close()
