#!/usr/bin/env python
"""One rank's share of a sharded fill on one GPU (what a rank of an N-GPU run executes, minus the all-gather):
    python tools/profile_shard.py cfg2 8 [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
P = named_config(cfg)
g = capi.UpcGpu(P, 0)
for i in range(steps):
    g.invalidate_tables()
    g.prepare_tables()
    g.fill_lumi_shard(int(os.environ.get("SHARD", "0")), n)
    print("step", i, g.fill_stats())
g.close()
