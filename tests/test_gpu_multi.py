"""Several GPUs behind one handle (upcgpu_create_multi, csrc/upc_group.cu) and the `-ngpus` option of the C++
drop-in: the (y, m) grid sharded by m rows over the devices of ONE process, exchanged with NCCL (all-gather) or with
peer stores from inside the cell kernel, events split by candidate ranges.  Everything must equal the single-GPU
result bit for bit -- cells are independent and the Philox counters do not know how many devices there are.
Skipped on boxes with fewer than two GPUs; the head-pool fallback test at the end needs one."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HOST = os.path.join(ROOT, "upcgen_b200", "host")

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        rt = C.CDLL("libcudart.so.12")
        n = C.c_int(0)
        return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0
    except OSError:
        return 0


NDEV = _ndev()
two = pytest.mark.skipif(NDEV < 2, reason="needs two GPUs")


@pytest.fixture(scope="module")
def capi():
    from upcgen_b200 import capi as c
    c.lib()
    return c


def _single(capi, P):
    g = capi.UpcGpu(P, 0)
    g.prepare_tables()
    if P.use_pol:
        lumi = g.fill_lumi()
        cs, ratio, tot = g.fold_sigma(sig_s=capi.elem_sigma_m(P, 1), sig_p=capi.elem_sigma_m(P, 2))
        g.sampler_build(cszm_s=capi.elem_cs_zm(P, 1), cszm_ps=capi.elem_cs_zm(P, 2))
    else:
        lumi = g.fill_lumi()
        cs, ratio, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
        g.sampler_build(cszm=None if P.ignore_csz else capi.elem_cs_zm(P))
    ev = g.generate(99, 1000, 50001)
    st = g.fill_stats()
    g.close()
    return lumi, cs, tot, ev, st


@two
@pytest.mark.parametrize("exchange", [0, 1])
@pytest.mark.parametrize("cfg,extra", [
    ("cfg2", "BINS_M 150\nBINS_Y 20\n"),                                   # form-factor flux (device QAGS) + XNXN
    ("cfg1", "PROC_ID 11\nUSE_POLARIZED_CS 1\nBINS_M 70\nBINS_Y 11\nMMIN 1\nMMAX 20\n"),  # polarised: two tables
    ("cfg5", "BINS_M 90\nBINS_Y 24\n"),                                    # Xe-Xe ALP: single production + decay
])
def test_group_equals_single_device(capi, cfg, extra, exchange):
    from upcgen_b200.config import named_config
    P = named_config(cfg, extra)
    lumi1, cs1, tot1, ev1, st1 = _single(capi, P)
    n = min(NDEV, 4)
    g = capi.UpcGpu(P, n_gpus=n)
    assert g.group_size() == n
    g.group_set_exchange(exchange)
    print(g.group_describe())
    g.prepare_tables()
    lumi = g.fill_lumi()
    st = g.fill_stats()
    if P.use_pol:
        assert np.array_equal(lumi[0], lumi1[0]) and np.array_equal(lumi[1], lumi1[1])
        cs, ratio, tot = g.fold_sigma(sig_s=capi.elem_sigma_m(P, 1), sig_p=capi.elem_sigma_m(P, 2))
        g.sampler_build(cszm_s=capi.elem_cs_zm(P, 1), cszm_ps=capi.elem_cs_zm(P, 2))
    else:
        assert np.array_equal(lumi, lumi1)
        cs, ratio, tot = g.fold_sigma(sig_m=capi.elem_sigma_m(P))
        g.sampler_build(cszm=None if P.ignore_csz else capi.elem_cs_zm(P))
    assert np.array_equal(cs, cs1) and tot == tot1
    # the work counters add up to the single-device fill's
    for k in ("qags_integrals", "qags_evals", "flux_rows", "band_pairs", "cells_evaluated"):
        assert st[k] == st1[k], k
    # every member holds the full table (it samples events from it)
    for r in range(n):
        m = g.group_member(r)
        kinds = (1, 2) if P.use_pol else (0,)
        for j, which in enumerate(kinds):
            ref = lumi1[j] if P.use_pol else lumi1
            assert np.array_equal(m.lumi_download(which), ref), (r, which)
    # events: candidate ranges split over the devices, same Philox counters -> the same events
    ev = g.generate(99, 1000, 50001)
    assert ev["n_accepted"] == ev1["n_accepted"]
    for k in ("npart", "pdg", "status", "mother", "p4", "aux"):
        assert np.array_equal(ev[k], ev1[k]), k
    g.close()


@two
def test_group_rejects_bad_requests(capi):
    from upcgen_b200.config import named_config
    P = named_config("cfg1", "BINS_M 16\nBINS_Y 8\n")
    with pytest.raises(capi.UpcGpuError):
        capi.UpcGpu(P, devices=[0, 0])
    with pytest.raises(capi.UpcGpuError):
        capi.UpcGpu(P, devices=[0, NDEV])
    g = capi.UpcGpu(P, 0)
    assert g.group_size() == 1
    with pytest.raises(capi.UpcGpuError):
        g.group_set_exchange(1)
    g.close()


PAR = """NUCLEUS_Z 82
NUCLEUS_A 208
SQRTS 5020
PROC_ID 13
NEVENTS 5000
DO_PT_CUT 1
PT_MIN 0.5
MMIN 4
MMAX 30
BINS_M 96
BINS_Y 14
BINS_Z 50
FLUX_POINT 0
BREAKUP_MODE 2
NON_ZERO_GAM_PT 1
SEED 4242
USE_ROOT_OUTPUT 0
USE_HEPMC_OUTPUT 1
"""


@two
def test_upcgen_cli_ngpus_equals_one_gpu(tmp_path):
    """`upcgen -ngpus N` (the C++ drop-in: UpcGenerator / UpcCrossSection over upcgpu_create_multi) writes the same
    events.hepmc, byte for byte, and reports the same cross sections as on one GPU."""
    subprocess.check_call(["make", "-s", "-C", HOST])
    outs = {}
    for n in (1, min(NDEV, 8)):
        d = tmp_path / f"n{n}"
        d.mkdir()
        (d / "my.in").write_text(PAR)
        r = subprocess.run([os.path.join(HOST, "upcgen"), "-parfile", "my.in", "-ngpus", str(n)], cwd=d,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        if n > 1:
            assert "GPU group" in r.stderr, r.stderr
        outs[n] = ((d / "events.hepmc").read_bytes(), [l for l in r.stdout.splitlines() if "cross section" in l])
    (a, xa), (b, xb) = outs.values()
    assert a == b and len(a) > 100000
    assert xa == xb


@two
def test_cfg4_sharded_over_the_group_reproduces_the_golden_subgrid(capi):
    """BASELINE config 4 at its full size (10001 x 1201 cells) filled by every device of the box through one handle:
    the oracle's golden sub-grid is reproduced, every member holds the same table bit for bit, and the work counters
    are those of the grid (one integral per (y row, b index, m row) of the form-factor flux)."""
    from upcgen_b200.config import named_config
    P = named_config("cfg4")
    g = capi.UpcGpu(P, n_gpus=min(NDEV, 8))
    g.prepare_tables()
    table = g.fill_lumi()
    st = g.fill_stats()
    assert st["qags_errors"] == 0 and st["qags_overflow"] == 0
    gold = np.load(os.path.join(ROOT, "tests", "golden", "cfg4_subgrid.npz"))
    im, iy = gold["im"], gold["iy"]
    ref = gold["lumi"] * float(gold["dm"]) * float(gold["dy"])
    sel = ref > 1e-250
    eg = np.max(np.abs(table[np.ix_(im, iy)][sel] - ref[sel]) / ref[sel])
    print("cfg4 on", g.group_size(), "GPUs: golden sub-grid max rel", eg, "integrals", st["qags_integrals"])
    assert eg < 1e-7
    for r in range(1, g.group_size()):
        assert np.array_equal(g.group_member(r).lumi_download(0), table), r
    g.close()


_PEER_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from upcgen_b200 import capi, dist as udist
from upcgen_b200.config import named_config
rank, local, world = udist.init_from_env("nccl")
dev = torch.device("cuda", local)
for cfg, extra in (("cfg2", "BINS_M 150\nBINS_Y 20\n"), ("cfg1", "PROC_ID 11\nUSE_POLARIZED_CS 1\nBINS_M 70\nBINS_Y 11\nMMIN 1\nMMAX 20\n")):
    P = named_config(cfg, extra)
    g = capi.UpcGpu(P, local)
    g.prepare_tables()
    ref = g.fill_lumi()                                  # the whole grid on this rank's GPU
    assert udist.setup_peer_exchange(g, rank, world)     # CUDA IPC mappings of every rank's tables
    for w in ((1, 2) if P.use_pol else (0,)):
        g.lumi_upload(w, np.zeros((P.nm, P.ny)))         # wipe: what is read back below was written by the peers
    dist.barrier()
    udist.fill_lumi_peers(g, rank, world, dev)
    got = [g.lumi_download(w) for w in ((1, 2) if P.use_pol else (0,))]
    exp = list(ref) if P.use_pol else [ref]
    assert all(np.array_equal(a, b) for a, b in zip(got, exp)), (cfg, rank)
    # ... and the NCCL all-gather path gives the same table
    for w in ((1, 2) if P.use_pol else (0,)):
        g.lumi_upload(w, np.zeros((P.nm, P.ny)))
    dist.barrier()
    udist.fill_lumi_distributed(g, rank, world, dev)
    got = [g.lumi_download(w) for w in ((1, 2) if P.use_pol else (0,))]
    assert all(np.array_equal(a, b) for a, b in zip(got, exp)), (cfg, rank, "nccl")
    dist.barrier()
    g.close()
if rank == 0:
    print("PEER_OK")
dist.destroy_process_group()
"""


@two
def test_torchrun_peer_store_exchange(tmp_path):
    """One process per GPU (the bench's launch mode): the cell kernel stores every finished cell into every rank's
    table through CUDA IPC mappings (upcgpu_lumi_ipc_export / _import, upcgpu_fill_lumi_shard_peers); every rank ends
    up with the single-GPU table, bit for bit, as it does through the NCCL all-gather."""
    script = tmp_path / "peer_worker.py"
    script.write_text(_PEER_WORKER.format(root=ROOT))
    n = min(NDEV, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29577", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "PEER_OK" in r.stdout


def test_head_state_pool_fallback():
    """The pool of QAGS hand-over states is a quarter of the integrals; when it runs dry the integral restarts in the
    second head pass and, if there is still no slot, in the large-workspace pass.  Forced here with a 3-slot pool
    (UPCGPU_TEST_HEAD_POOL): the table is the same bit for bit, and the overflow pass did run."""
    code = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np
from upcgen_b200 import capi
from upcgen_b200.config import named_config
P = named_config("cfg2", "BINS_M 40\nBINS_Y 12\n")
g = capi.UpcGpu(P, 0)
g.prepare_tables()
t = g.fill_lumi()
st = g.fill_stats()
np.save(sys.argv[1], t)
print("RESULT " + json.dumps({k: st[k] for k in ("qags_integrals", "qags_evals", "qags_overflow", "qags_errors")}))
""" % ROOT
    import json
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, env in (("full", {}), ("tiny", {"UPCGPU_TEST_HEAD_POOL": "3"})):
            f = os.path.join(d, tag + ".npy")
            r = subprocess.run([sys.executable, "-c", code, f], capture_output=True, text=True, timeout=600,
                               env={**os.environ, **env})
            assert r.returncode == 0, r.stderr[-2000:]
            res[tag] = (np.load(f), json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:]))
    (t0, s0), (t1, s1) = res["full"], res["tiny"]
    print(s0, s1)
    assert s0["qags_overflow"] == 0 and s1["qags_overflow"] > 0
    assert s0["qags_errors"] == 0 and s1["qags_errors"] == 0
    assert s0["qags_integrals"] == s1["qags_integrals"]
    # the restarted integrals take the same decisions: same evaluation count, same table
    assert s0["qags_evals"] == s1["qags_evals"]
    assert np.array_equal(t0, t1)
