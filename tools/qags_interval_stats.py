#!/usr/bin/env python
"""Which dyadic sub-intervals of [0, 10] does gsl_integration_qags bisect for the form-factor flux
integrals of a grid?  (oracle trace; design input for the head kernel's candidate set)
    python tools/qags_interval_stats.py cfg2 [row stride]
"""
import ctypes as C
import os
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 37
P = named_config(cfg)
o = pyoracle.Oracle(P)
L = o.L
L.upco_qags_fluxform_trace.restype = C.c_int
L.upco_qags_fluxform_trace.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_uint), C.c_int, C.POINTER(C.c_int)]
R, hc, g1, nb = P.R, 0.1973269718, P.g1, 120
dm, dy = (P.mmax - P.mmin) / P.nm, (P.ymax - P.ymin) / P.ny
buf = (C.c_uint * 64)()
ne = C.c_int()
seqs = Counter()
per_round = [Counter() for _ in range(40)]
n_int = 0
tot_rules = 0
rows = [(im, iy) for im in range(P.nm) for iy in range(P.ny + 1)][::stride]
shared_cnt = Counter()
for im, iy in rows:
    k = (P.mmin + dm * im) / 2 * np.exp(P.ymin + dy * iy)
    bmax = max(5 * g1 * hc / k, 5 * R)
    ld = (np.log(bmax) - np.log(0.05 * R)) / nb
    for i in range(nb):
        b = (0.05 * R * np.exp(i * ld) + 0.05 * R * np.exp((i + 1) * ld)) / 2
        if b > 2 * R:
            break
        n = L.upco_qags_fluxform_trace(o.h, b, k, buf, 64, C.byref(ne))
        tr = [buf[j] for j in range(n)]
        n_int += 1
        tot_rules += 1 + 2 * n
        for r, t in enumerate(tr):
            per_round[r][t] += 1
        seqs[tuple(tr[5:])] += 1
print(f"{cfg}: {len(rows)} rows, {n_int} integrals, {tot_rules} GK21 rules")
allc = Counter()
for r in range(40):
    if not per_round[r]:
        break
    tot = sum(per_round[r].values())
    top = per_round[r].most_common(6)
    print(f"bisection {r + 1}: {tot} integrals ({tot / n_int:.3f}); " + ", ".join(f"({t >> 24},{t & 0xffffff}) {c / tot:.3f}" for t, c in top))
    if r >= 5:
        allc.update(per_round[r])
tail = sum(allc.values())
print("tail bisections:", tail, "distinct intervals:", len(allc))
cum = 0
for n, (t, c) in enumerate(allc.most_common(60)):
    cum += c
    print(f"  {n + 1:3d} ({t >> 24},{t & 0xffffff}) {c / tail:.4f} cum {cum / tail:.4f}")
