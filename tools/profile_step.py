#!/usr/bin/env python
"""Runs W + K table steps of a workload and nothing else -- the command ncu wraps.
    python tools/profile_step.py cfg2 1 2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
events = int(sys.argv[4]) if len(sys.argv) > 4 else 0
P = named_config(cfg)
g = capi.UpcGpu(P, 0)
fold = not (P.use_pol and P.proc_id in (22, 111))  # cfg3: lumi tables only
sig = {} if not fold else dict(sig_s=capi.elem_sigma_m(P, 1), sig_p=capi.elem_sigma_m(P, 2)) if P.use_pol else dict(sig_m=capi.elem_sigma_m(P, 0))
for i in range(warm + steps):
    g.invalidate_tables()
    g.prepare_tables()
    g.fill_lumi_shard(0, 1)
    if fold:
        g.fold_sigma(download=False, **sig)
    print("step", i, g.fill_stats())
if events:
    g.sampler_build(cszm=None if P.ignore_csz else capi.elem_cs_zm(P, 0))
    print("accepted", g.generate_device(12345, 0, events))
g.close()
