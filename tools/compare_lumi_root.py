#!/usr/bin/env python
"""Compare a two-photon-luminosity cache written by the reference (twoPhotonLumi.root / twoPhotonLumiPol.root,
src/UpcCrossSection.cpp:481-507, :564-585) with the table this library fills for the same parameters.in.

    python tools/compare_lumi_root.py parameters.in twoPhotonLumi.root

The file is read without ROOT (upcgpu_root_hist_read); histogram bin (im + 1, iy + 1) holds table[im][iy].
Needs a GPU for the fill; with --read-only it just prints what the file holds.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upcgen_b200 import capi  # noqa: E402
from upcgen_b200.config import UpcParams  # noqa: E402


def table_from_hist(h, nm, ny):
    if h["dim"] != 2 or (h["nx"], h["ny"]) != (nm, ny):
        raise SystemExit(f"histogram is {h['nx']} x {h['ny']}, the parameters ask for {nm} x {ny}")
    return np.ascontiguousarray(h["cells"][1:ny + 1, 1:nm + 1].T)   # cells[y bin][x bin] -> table[im][iy]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("parameters")
    ap.add_argument("rootfile")
    ap.add_argument("--read-only", action="store_true")
    a = ap.parse_args()
    P = UpcParams.from_file(a.parameters).init()
    names = ("hD2LDMDY_s", "hD2LDMDY_p") if P.use_pol else ("hD2LDMDY",)
    ref = [table_from_hist(capi.root_hist_read(a.rootfile, n), P.nm, P.ny) for n in names]
    for n, t in zip(names, ref):
        print(f"{n}: {t.shape[0]} x {t.shape[1]}, sum {t.sum():.12e}, max {t.max():.6e}")
    if a.read_only:
        return
    g = capi.UpcGpu(P, 0)
    got = g.fill_lumi()
    got = list(got) if P.use_pol else [got]
    for n, r, t in zip(names, ref, got):
        sel = r > 0
        e = np.abs(t[sel] - r[sel]) / r[sel]
        print(f"{n}: max rel diff {e.max():.3e}, median {np.median(e):.3e} over {sel.sum()} cells")
    g.close()


if __name__ == "__main__":
    main()
