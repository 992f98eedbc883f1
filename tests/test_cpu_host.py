"""CPU-side tests (no GPU): parameters.in semantics, the C-ABI library's exported surface, the
oracle against the committed golden fixtures, and the host-side sharding logic with gloo."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- parameters.in ---------------------------------------------------------------------------
def test_parameters_in_semantics():
    from upcgen_b200.config import M_PROT, UpcParams
    txt = """# a comment line
NUCLEUS_Z 54
NUCLEUS_A 129   # trailing comment
SQRTS 5440
UNKNOWN_KEY 17
BINS_M 77

PROC_ID 13
FLUX_POINT 0
USE_POLARIZED_CS 1
"""
    P = UpcParams.from_text(txt)
    assert (P.Z, P.A, P.nm, P.proc_id, P.is_point, P.use_pol, P.gen_use_pol) == (54, 129, 77, 13, 0, 1, 1)
    assert P.g1 == 5440.0 / (2.0 * M_PROT) == P.g2
    # Q1: gtot stays at the value of the DEFAULT sqrts (constructor), not 5440
    assert P.gtot == pytest.approx(5020.0 / (2.0 * M_PROT), rel=1e-12)
    assert P.ny == 121 and P.mmax == 50.0          # defaults kept, unknown key ignored
    assert UpcParams.from_file("/nonexistent/parameters.in").nm == 1001   # missing file -> defaults


def test_init_overrides():
    from upcgen_b200.config import M_TAU, named_config
    P1 = named_config("cfg1")
    assert P1.mmin == 3.56 and P1.proc_id == 15   # 2 m_tau = 3.55372 < 3.56: MMIN is kept
    P3 = named_config("cfg3")
    assert (P3.nm, P3.mmin, P3.mmax, P3.nz, P3.zmin) == (1000, 0.05, 50.0, 198, -0.99)
    assert P3.use_pol == 1 and P3.gen_use_pol == 0      # Q5: only the generator's flag is cleared
    P5 = named_config("cfg5")
    assert P5.mmin == pytest.approx(0.96) and P5.mmax == pytest.approx(1.04) and P5.ignore_csz
    P4 = named_config("cfg4")
    assert P4.nm * P4.ny == 12011201


def test_mmin_rule():
    from upcgen_b200.config import M_TAU, named_config
    # the repo file has MMIN 3.56 with tau pairs: 2 m_tau = 3.55372 < 3.56, so MMIN stays
    assert 2 * M_TAU < 3.56
    assert named_config("cfg1").mmin == 3.56
    assert named_config("cfg1", "MMIN 1.0\n").mmin == pytest.approx(2 * M_TAU)


# ---- C-ABI surface -----------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from upcgen_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "upcgpu.h")).read()
    declared = set(re.findall(r"\b(upcgpu_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"upcgpu_ctx", "upcgpu_params"}
    lib = ctypes.CDLL(capi.SO_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert lib.upcgpu_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    from upcgen_b200 import capi
    from upcgen_b200.config import named_config
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.UpcGpuError) as e:
        capi.UpcGpu(named_config("cfg1"), 0)
    assert e.value.code == capi.ENODEV and "no CPU fallback" in str(e.value)


def test_cparams_layout_matches_header():
    """The ctypes mirror lists the fields of upcgpu_params in header order."""
    from upcgen_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "upcgpu.h")).read()
    body = hdr[hdr.index("typedef struct {", hdr.index("Parameter block")):hdr.index("} upcgpu_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        parts = decl.replace(",", " ").split()
        names += parts[1:]
    assert names == [n for n, _ in capi.CParams._fields_]


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pkg = os.path.join(ROOT, "upcgen_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "upc_oracle" not in txt and "libupcoracle" not in txt, f


# ---- oracle vs committed golden fixtures -------------------------------------------------------
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "cfg5"])
def test_oracle_reproduces_golden_subgrid(get_oracle, cfg):
    P, o = get_oracle(cfg)
    g = np.load(os.path.join(GOLD, f"{cfg}_subgrid.npz"))
    sel_m, sel_y = [0, len(g["M"]) // 2, -1], [0, len(g["Y"]) // 2, -1]
    for a in sel_m:
        for b in sel_y:
            if P.use_pol:
                s, p = o.lumi_pol(float(g["M"][a]), float(g["Y"][b]))
                assert s == g["lumi_s"][a, b] and p == g["lumi_p"][a, b]
            else:
                assert o.lumi(float(g["M"][a]), float(g["Y"][b])) == g["lumi"][a, b]
    assert o.rho0() == float(g["rho0"])
    assert np.array_equal(o.gaa()[1], g["gaa_y"])
    if "bk_spline" in g:
        assert np.array_equal(o.breakup_spline(g["bk_b"]), g["bk_spline"])
    if "ff_flux" in g:
        fl, ne = o.flux_form(g["ff_b"][:, None], g["ff_k"][None, :], with_neval=True)
        assert np.array_equal(fl, g["ff_flux"]) and np.array_equal(ne, g["ff_neval"])


def test_oracle_sampler_golden(oracle_mod):
    g = np.load(os.path.join(GOLD, "sampler_vectors.npz"))
    s = oracle_mod.pdf_init(g["bins"])
    assert np.array_equal(s, g["sum"])
    for i in range(g["r1"].size):
        k, x, y = oracle_mod.sample2d(s, g["xe"], g["ye"], g["r1"][i], g["r2"][i])
        assert (k, x, y) == (g["k"][i], g["x"][i], g["y"][i])


def test_oracle_symmetry_used_by_the_gpu_rows(get_oracle):
    """k2(im, iy) = k1(im, ny-iy) up to an ulp on a y grid symmetric about 0: the identity the GPU
    path uses to integrate every flux once ('flux rows')."""
    P, o = get_oracle("cfg2")
    iy = np.arange(1, P.ny)
    y = P.ymin + P.dy * iy
    y_m = P.ymin + P.dy * (P.ny - iy)
    assert np.max(np.abs(y_m + y)) < 4e-15
    # and flux(b, k) is insensitive to that perturbation at the 1e-13 level
    k = 3.56 / 2 * np.exp(-y[:40]); k2 = 3.56 / 2 * np.exp(y_m[:40])
    b = np.full_like(k, 5.0)
    f1, n1 = o.flux_form(b, k, with_neval=True); f2, n2 = o.flux_form(b, k2, with_neval=True)
    # (relative to the row's scale: the integral changes sign, so single fluxes pass through 0)
    assert np.array_equal(n1, n2) and np.max(np.abs(f1 - f2)) < 1e-13 * np.max(f1)


def test_qags_count_fixture_is_consistent(get_oracle):
    """tests/golden/cfg2_qags_counts.json (used by bench.py's roofline) vs a fresh oracle count of
    the first mass row."""
    from upcgen_b200.config import HC
    P, o = get_oracle("cfg2")
    fx = json.load(open(os.path.join(GOLD, "cfg2_qags_counts.json")))
    Y = P.ymin + P.dy * np.arange(P.ny + 1)
    k = P.mmin / 2.0 * np.exp(Y)
    bmin = 0.05 * P.R
    bmax = np.maximum(5.0 * P.g1 * HC / k, 5.0 * P.R)
    ld = (np.log(bmax) - np.log(bmin)) / P.nb1
    i = np.arange(P.nb1)
    b = (bmin * np.exp((i[None, :] + 1.0) * ld[:, None]) + bmin * np.exp(i[None, :] * ld[:, None])) / 2.0
    sel = ~(b > 2.0 * P.R)
    _, ne = o.flux_form(b[sel], np.broadcast_to(k[:, None], b.shape)[sel], with_neval=True)
    assert int(ne.sum()) == fx["evals_per_m_first8"][0]
    assert fx["rows"] == P.nm * (P.ny + 1)


# ---- sharding logic, world_size 2 on gloo ------------------------------------------------------
_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from upcgen_b200 import dist as udist
rank, local, world = udist.init_from_env("gloo")
nm, ny = 37, 5
full = np.arange(nm * ny, dtype=np.float64).reshape(nm, ny) + 0.25
rows = udist.cyclic_rows(nm, rank, world)
rps = udist.rows_per_shard(nm, world)
shard = np.zeros((rps, ny)); shard[: len(rows)] = full[rows]
src = torch.from_numpy(shard.ravel().copy())
dst = torch.empty(world * rps * ny, dtype=torch.float64)
dist.all_gather_into_tensor(dst, src)
got = udist.unpack_host(dst.numpy(), nm, ny, world)
assert np.array_equal(got, full), "unpack mismatch"
# event sharding: Philox counter ranges are disjoint and cover [0, n)
n = 1000
per = n // world
mine = set(range(rank * per, (rank + 1) * per))
t = torch.tensor([len(mine)]); dist.all_reduce(t)
assert t.item() == n
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_sharding_layout_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout


def test_cyclic_rows_cover_grid():
    from upcgen_b200 import dist as udist
    for nm, w in [(1001, 8), (1000, 3), (5, 8), (10001, 4)]:
        seen = sorted(sum((udist.cyclic_rows(nm, r, w) for r in range(w)), []))
        assert seen == list(range(nm))
        assert max(len(udist.cyclic_rows(nm, r, w)) for r in range(w)) == udist.rows_per_shard(nm, w)
        blk = udist.shard_block(nm, w)
        for r in range(w):                                                          # local index -> m row, as the kernels do
            rows = udist.cyclic_rows(nm, r, w)

            def to_im(li):
                cyc = li // blk
                pos = w - 1 - r if cyc & 1 else r
                return (cyc * w + pos) * blk + li % blk
            assert rows == [to_im(li) for li in range(len(rows))]
    # the bench grid on 8 GPUs: whole warps of neighbouring rows, shards within 3 % of the mean (the last one is short)
    sizes = [len(udist.cyclic_rows(1001, r, 8)) for r in range(8)]
    assert udist.shard_block(1001, 8) == 32 and max(sizes) == 128 and sorted(sizes)[1] >= 96
    # the snake: rank 0 holds the first block of even cycles and the last block of odd ones
    r0 = udist.cyclic_rows(1001, 0, 8)
    assert r0[:32] == list(range(32)) and r0[32:64] == list(range(15 * 32, 16 * 32)) and r0[64:96] == list(range(16 * 32, 17 * 32))


def test_head_interval_table_covers_the_oracles_bisections(get_oracle):
    """The CUDA head kernel (csrc/upc_qags_head.cuh) runs an integral for as long as QAGS bisects one of 15 tabulated
    intervals.  The oracle's trace of a sample of cfg2 integrals must stay inside that set almost always, and the
    first four bisections must be the leftmost interval every time (the statistic the kernel's table was chosen from;
    tools/qags_interval_stats.py prints the full version)."""
    import ctypes as C
    P, o = get_oracle("cfg2")
    L = o.L
    L.upco_qags_fluxform_trace.restype = C.c_int
    L.upco_qags_fluxform_trace.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_uint), C.c_int,
                                           C.POINTER(C.c_int)]
    # heap index (1 << level) + position of the tabulated parents, as in kHdParent
    parents = {1, 2, 4, 8, 16, 17, 32, 64, 128, 33, 256, 34, 3, 512, 1024}
    hc, nb = 0.1973269718, 120
    buf = (C.c_uint * 64)()
    ne = C.c_int()
    total = inside = 0
    rows = [(im, iy) for im in range(0, P.nm, 97) for iy in range(0, P.ny + 1, 13)]
    for im, iy in rows:
        k = (P.mmin + P.dm * im) / 2 * np.exp(P.ymin + P.dy * iy)
        bmax = max(5 * P.g1 * hc / k, 5 * P.R)
        ld = (np.log(bmax) - np.log(0.05 * P.R)) / nb
        for i in range(0, nb, 7):
            b = (0.05 * P.R * np.exp(i * ld) + 0.05 * P.R * np.exp((i + 1) * ld)) / 2
            if b > 2 * P.R:
                break
            n = L.upco_qags_fluxform_trace(o.h, b, k, buf, 64, C.byref(ne))
            assert ne.value == 21 * (1 + 2 * n)
            heaps = [(1 << (buf[j] >> 24)) + (buf[j] & 0xFFFFFF) for j in range(n)]
            assert heaps[:4] == [1, 2, 4, 8][:min(n, 4)]
            total += n
            inside += sum(h in parents for h in heaps)
    assert total > 2000 and inside / total > 0.995, (inside, total)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
    keys, the same metric / unit / config object as the GPU arm, and a cpu_baseline describing the run."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "lumi_cells_per_s" and line["unit"] == "cells/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0
    assert set(line["config"]) >= {"workload", "cells", "grid"} and line["config"]["grid"] == [1001, 121]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_hepmc_writer_bytes_equal_stream_formatting(tmp_path):
    """WriterHepMC formats its records with std::to_chars (three times the pace of operator<<); the file must be, byte
    for byte, what the reference's stream insertions with setprecision(9) give (include/UpcGenerator.h:214-241) -- random
    momenta over twelve decades, zeros, denormals, 1e300, infinities and NaN."""
    import filecmp
    exe = str(tmp_path / "hepmc_check")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "upcgen_b200", "host"),
                           "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "cpp", "hepmc_check.cpp")])
    r = subprocess.run([exe, "20000", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    print(r.stdout.strip())
    assert filecmp.cmp(tmp_path / "fast.hepmc", tmp_path / "ref.hepmc", shallow=False)
    lines = (tmp_path / "fast.hepmc").read_text().splitlines()
    assert lines[0] == "HepMC::Version 3.02.04" and lines[-1] == "HepMC::Asciiv3-END_EVENT_LISTING"
    assert any(" inf -inf nan " in l for l in lines) and any(" 1e+300 4.94065646e-324 " in l for l in lines)
