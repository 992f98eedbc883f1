#!/usr/bin/env python
"""Oracle-side count of the form-factor flux integrals of a whole grid (deterministic):
number of QAGS integrals and of integrand evaluations over all distinct photon-energy rows
(im, r), r = 0..ny (y-symmetric grid).  Written to tests/golden/<cfg>_qags_counts.json; the GPU
test asserts that the device QAGS took exactly as many evaluations over the full grid, and
bench.py uses the count for the roofline arithmetic (100 flop per evaluation).

    python tools/gen_qags_counts.py cfg2
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from upcgen_b200.config import HC, named_config  # noqa: E402


def main(cfg):
    P = named_config(cfg)
    assert not P.is_point and P.ymin == -P.ymax
    o = pyoracle.Oracle(P)
    nb = P.nb1
    tot_int = 0
    tot_ev = 0
    per_m = []
    t0 = time.time()
    for im in range(P.nm):
        M = P.mmin + P.dm * im
        Y = P.ymin + P.dy * np.arange(P.ny + 1)
        k = M / 2.0 * np.exp(Y)
        bmin = 0.05 * P.R
        bmax = np.maximum(5.0 * P.g1 * HC / k, 5.0 * P.R)
        ld = (np.log(bmax) - np.log(bmin)) / nb
        i = np.arange(nb)
        bl = bmin * np.exp(i[None, :] * ld[:, None])
        bh = bmin * np.exp((i[None, :] + 1.0) * ld[:, None])
        b = (bh + bl) / 2.0
        sel = ~(b > 2.0 * P.R)
        kk = np.broadcast_to(k[:, None], b.shape)
        _, ne = o.flux_form(b[sel], kk[sel], with_neval=True)
        tot_int += int(sel.sum())
        tot_ev += int(ne.sum())
        per_m.append(int(ne.sum()))
        if im % 50 == 0:
            print(im, tot_int, tot_ev, "%.0fs" % (time.time() - t0), flush=True)
    out = dict(config=cfg, nm=P.nm, ny=P.ny, rows=P.nm * (P.ny + 1), qags_integrals=tot_int, qags_evals=tot_ev,
               evals_per_m_first8=per_m[:8], generated_by="tools/gen_qags_counts.py (CPU oracle)")
    path = os.path.join(ROOT, "tests", "golden", f"{cfg}_qags_counts.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
