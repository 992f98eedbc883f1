"""Diagnostic: is the fill deterministic run to run, and how far apart are the tables with a starved head pool?"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from upcgen_b200 import capi
from upcgen_b200.config import named_config
P = named_config("cfg2", "BINS_M 40\nBINS_Y 12\n")
g = capi.UpcGpu(P, 0)
g.prepare_tables()
t = g.fill_lumi()
t2 = g.fill_lumi()
print("same process twice equal:", np.array_equal(t, t2), np.max(np.abs(t - t2) / t))
np.save(sys.argv[1], t)
""" % ROOT
with tempfile.TemporaryDirectory() as d:
    out = {}
    for tag, env in (("full", {}), ("full2", {}), ("tiny", {"UPCGPU_TEST_HEAD_POOL": "3"}), ("tiny2", {"UPCGPU_TEST_HEAD_POOL": "3"})):
        f = os.path.join(d, tag + ".npy")
        r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, env={**os.environ, **env})
        print(tag, r.stdout.strip(), r.stderr[-300:])
        out[tag] = np.load(f)
    for a, b in (("full", "full2"), ("full", "tiny"), ("tiny", "tiny2")):
        x, y = out[a], out[b]
        rel = np.abs(x - y) / x
        print(a, b, "equal", np.array_equal(x, y), "max rel", rel.max(), "n differing", int((x != y).sum()), "of", x.size)
