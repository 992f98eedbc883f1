#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/upc_oracle.c).

The reference ships no golden vectors and cannot be built here as-is; the
fixtures therefore come from the oracle, which is itself pinned against scipy/QUADPACK, mpmath,
the survey's probe anchors and the reference's own translation units compiled against a
GSL/ROOT shim (oracle/refshim).  Fixtures are small sub-grids of every BASELINE config including
the corner cells, plus sub-function vectors.

    python tools/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from upcgen_b200.config import named_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def subgrid(P, n_m, n_y):
    im = np.unique(np.round(np.linspace(0, P.nm - 1, n_m)).astype(int))
    iy = np.unique(np.concatenate([np.round(np.linspace(0, P.ny - 1, n_y)).astype(int), [P.ny // 2]]))
    return im, iy


def main():
    os.makedirs(OUT, exist_ok=True)
    for cfg, (n_m, n_y) in {"cfg1": (16, 9), "cfg2": (12, 7), "cfg3": (10, 7), "cfg4": (6, 5), "cfg5": (12, 7)}.items():
        P = named_config(cfg)
        o = pyoracle.Oracle(P)
        im, iy = subgrid(P, n_m, n_y)
        M = P.mmin + P.dm * im
        Y = P.ymin + P.dy * iy
        out = dict(im=im, iy=iy, M=M, Y=Y, dm=P.dm, dy=P.dy)
        if P.use_pol:
            v = np.array([[o.lumi_pol(float(m), float(y)) for y in Y] for m in M])
            out["lumi_s"], out["lumi_p"] = v[..., 0], v[..., 1]
        else:
            out["lumi"] = np.array([[o.lumi(float(m), float(y)) for y in Y] for m in M])
        out["rho0"] = o.rho0()
        out["sigma_nn"] = o.sigma_nn()
        b, g, c, ta = o.gaa()
        out["gaa_y"], out["gaa_c"], out["ta_y"] = g, c, ta
        if P.breakup_mode > 1:
            bb = np.array([1e-6, 0.334, 2.0, 6.68, 10.0, 13.36, 15.0, 17.5, 19.999, 20.0])
            out["bk_b"] = bb
            out["bk_spline"] = o.breakup_spline(bb)
            out["bk_raw"] = o.breakup_raw(bb, P.breakup_mode)
        if not P.is_point:
            bl = np.exp(np.linspace(np.log(0.05 * P.R), np.log(2 * P.R), 9))
            kl = np.exp(np.linspace(np.log(5e-3), np.log(5e3), 7))
            fl, ne = o.flux_form(bl[:, None], kl[None, :], with_neval=True)
            out["ff_b"], out["ff_k"], out["ff_flux"], out["ff_neval"] = bl, kl, fl, ne
        bp = np.exp(np.linspace(np.log(P.R), np.log(1e4), 9))
        kp = np.exp(np.linspace(np.log(5e-3), np.log(5e3), 7))
        out["pt_b"], out["pt_k"], out["pt_flux"] = bp, kp, o.flux_point(bp[:, None], kp[None, :])
        if P.proc_id in (11, 13, 15, 51):
            out["sigma_m"] = o.sigma_m(M)
        np.savez_compressed(os.path.join(OUT, f"{cfg}_subgrid.npz"), **out)
        print(cfg, "cells", im.size * iy.size)
    # sampler vectors on a small synthetic histogram
    rng = np.random.default_rng(2024)
    bins = rng.uniform(0, 1, (7, 11))
    bins[2, 3:6] = 0.0
    s = pyoracle.pdf_init(bins)
    xe = np.linspace(-1.5, 2.0, 8)
    ye = np.linspace(0.3, 4.7, 12)
    r1 = np.concatenate([rng.uniform(0, s[-1] * (1 - 1e-12), 40), s[[1, 5, 24, 30]], np.nextafter(s[[5, 24]], 0), [0.0]])
    r1 = r1[r1 < s[-1]]
    r2 = rng.uniform(0, 1, r1.size)
    res = np.array([pyoracle.sample2d(s, xe, ye, a, b) for a, b in zip(r1, r2)])
    np.savez_compressed(os.path.join(OUT, "sampler_vectors.npz"), bins=bins, sum=s, xe=xe, ye=ye, r1=r1, r2=r2,
                        k=res[:, 0].astype(np.int64), x=res[:, 1], y=res[:, 2])
    print("sampler vectors", r1.size)


if __name__ == "__main__":
    main()
