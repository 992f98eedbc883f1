#!/usr/bin/env python
"""bench.py -- the hot path of upcgen on B200: luminosity/sigma table fill (cells/s) + events/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2] [--no-legs]

One STEP = one pass of the table path over the whole (y, m) grid of the workload:
    prepare the lookup tables (T1-T4)  ->  fill the two-photon luminosity table (F1-F3, L1-L3); with N > 1 the m rows
    are dealt to the ranks and the cell kernel stores every finished cell into every rank's table over NVLink (CUDA
    IPC peer memory; UPCGPU_EXCHANGE=nccl: an all-gather instead)  ->  fold with sigma(m) (X1).
`value` is cells/s with everything resident in HBM, timed with CUDA events on the library's stream (max over ranks);
`e2e` is the same step through the host-buffer C-ABI calls (upcgpu_fill_lumi / upcgpu_fold_sigma; per rank for N > 1:
upcgpu_fill_lumi_shard_peers + upcgpu_lumi_download + upcgpu_fold_sigma) with pinned host buffers and the copies inside
the timed region.  The event stage (S1-S3, E1-E5) is timed separately and reported under "events".

The headline workload is cfg2 (BASELINE configs[1]).  The same JSON line carries the north-star target as two more
objects, measured by the same code at the same N: "cfg4" (10001 x 1201 cells, table step and e2e) and "events_cfg5"
(10^7 candidates of the Xe-Xe ALP config); --no-legs skips them.

--impl reference times the reference's own CPU implementation of the path on all host threads, on a bounded sample of
the same grid: oracle/_ref (the reference's sources compiled against the GSL/ROOT shim; kind "reference") when it was
built, else the oracle port (kind "port").  The reference binary itself needs ROOT + GSL and cannot be built here.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_TEXT = {
    "cfg1": "cfg1: Pb-Pb 5.02 TeV ditau PROC_ID 15, point flux, no breakup, 1001x121 grid",
    "cfg2": "cfg2: Pb-Pb 5.02 TeV dimuon PROC_ID 13, FLUX_POINT 0 (Woods-Saxon form-factor flux, QAGS), "
            "BREAKUP_MODE 2 XNXN, 1001x121 (m,y) grid, 1e6 events",
    "cfg3": "cfg3: Pb-Pb light-by-light grid, USE_POLARIZED_CS 1, BREAKUP_MODE 4, 1000x121 grid (lumi tables only)",
    "cfg4": "cfg4: Pb-Pb dielectron PROC_ID 11, MMIN 1 MMAX 100, 10001x1201 grid, form-factor flux",
    "cfg5": "cfg5: Xe-Xe 5.44 TeV ALP PROC_ID 51, NON_ZERO_GAM_PT 1, 0N0N, 1001x121 grid, 1e7 events",
}

# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, and its FP64-pipe activity, from
# the committed `ncu --set full` captures (profiles/r02_ncu_qags_head_pass1_cfg2.txt, r01_v13_ncu_qags_rows_cfg2.txt, r02_ncu_k_cells_cfg{1,2}.txt)
NCU_QUOTED = {
    ("cfg2", "k_flux_qags_head"): {"traffic": 672.9e6 + 496.7e6, "fp64_pipe_pct": 59.3,
                                   "source": "profiles/r02_ncu_qags_head_pass1_cfg2.txt"},
    ("cfg2", "k_flux_qags_rows"): {"traffic": 28.9e6 + 0.06e6, "fp64_pipe_pct": 3.8,
                                   "source": "profiles/r01_v13_ncu_qags_rows_cfg2.txt"},
    ("cfg2", "k_cells"): {"traffic": 234.8e6 + 6.6e6, "fp64_pipe_pct": 44.4,
                          "source": "profiles/r02_ncu_k_cells_cfg2.txt"},
    ("cfg1", "k_cells"): {"traffic": 234.5e6 + 4.4e6, "fp64_pipe_pct": 50.3,
                          "source": "profiles/r02_ncu_k_cells_cfg1.txt"},
}

# algorithmic work per unit
FLOP_PER_QAGS_EVAL = 100.0          # one integrand evaluation of fluxFormIntegrand
FLOP_CELL = {(0, 0): 1.19e6, (0, 1): 1.91e6, (1, 0): 1.30e6, (1, 1): 2.05e6}  # (pol, breakup) per cell


def make_config(workload, P, world, peer=None):
    """The `config` object of the JSON line (the same for both arms)."""
    how = ("the cell kernel stores every finished cell into the table of every rank over NVLink (CUDA IPC peer memory); "
           "a one-element NCCL all-reduce orders the fold behind it" if peer else "NCCL all-gather of the table")
    return {"workload": WORKLOAD_TEXT[workload], "cells": P.nm * P.ny, "grid": [P.nm, P.ny],
            "l2": "flushed between timed steps (256 MiB device write)",
            "reflection": "columns iy > ny/2 of a y grid symmetric about 0 are written from their mirror images",
            "parallelism": f"m rows in blocks of 32 dealt in a snake over {world} GPU(s); {how}"
            if world > 1 else "single GPU"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.index = ",".join(str(i) for i in indices)   # the GPUs of the job (one node, ranks = device ordinals)
        self.n_gpus = len(indices)
        self.rows = []          # (arrival time, csv line)
        self.proc = None
        self.t_begin = 0.0

    def mark_begin(self):
        """Start of the timed region: only samples that arrive after this moment are used.  (The sampler itself is
        started BEFORE the warm-up steps, so that spawning nvidia-smi does not fall into a timed step.)"""
        self.t_begin = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def count(self):
        """samples that have arrived since mark_begin"""
        return sum(1 for t_arr, _ in list(self.rows) if t_arr >= self.t_begin) // self.n_gpus

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t_arr, r in self.rows:
            if t_arr < self.t_begin:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus": self.index}


def make_cpu_leg(workload, cores):
    """Returns (step_fn, n_cells, kind, sample_text).  Preferred: the REFERENCE's own grid driver
    (src/UpcCrossSection.cpp prepareTwoPhotonLumi, OpenMP static m-slabs, -nthreads = cores) from
    oracle/_ref -- the reference's sources compiled against the GSL/ROOT shim -- on a coarser grid
    over the SAME (m, y) ranges (bounded sample: the cost of a cell depends on (m, y) only).
    Fallback: the oracle port on a strided sub-grid."""
    import tempfile

    from oracle import pyoracle, pyref
    from upcgen_b200.config import named_config
    P = named_config(workload)
    if pyref.available() and P.proc_id in (11, 13, 15, 51):
        nm_s, ny_s = 64, 11
        Ps = named_config(workload, f"BINS_M {nm_s}\nBINS_Y {ny_s}\n")
        ref = pyref.Reference(Ps, with_breakup_table=True)

        def step():
            with tempfile.TemporaryDirectory() as d:   # fresh directory: no cached lumi file
                ref.grid_and_fold(d, nthreads=cores)

        sample = (f"{nm_s}x{ny_s} = {nm_s * ny_s} cells over the same m and y ranges as the {P.nm}x{P.ny} grid; "
                  f"the reference's prepareTwoPhotonLumi + calcNucCrossSectionYM (sources compiled against the "
                  f"GSL/ROOT shim, oracle/_ref), {cores} OpenMP threads; table set-up excluded")
        return step, nm_s * ny_s, "reference", sample
    o = pyoracle.Oracle(P, threads=cores)
    im_step = max(1, P.nm // 64)
    iy_step = max(1, P.ny // 11)
    n_cells = len(range(0, P.nm, im_step)) * len(range(0, P.ny, iy_step))

    def step():
        o.fill_lumi(im_step=im_step, iy_step=iy_step)

    sample = (f"{n_cells} cells = every {im_step}th m row x every {iy_step}th y column of the {P.nm}x{P.ny} grid; "
              f"oracle/upc_oracle.c (port), {cores} OpenMP threads; lumi fill only")
    return step, n_cells, "port", sample


def reference_arm(args):
    """CPU arm: the reference's own sources from oracle/_ref (kind 'reference'; else the oracle port, kind 'port') on
    all host threads, bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from upcgen_b200.config import named_config
    P = named_config(args.workload)
    cores = os.cpu_count() or 1
    step, n_cells, kind, sample = make_cpu_leg(args.workload, cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_cells / dt
    line = {
        "impl": "reference", "metric": "lumi_cells_per_s", "value": val, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.workload, P, args.gpus,
                              args.gpus > 1 and os.environ.get("UPCGPU_EXCHANGE", "") != "nccl"),  # = the GPU arm's
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference binary needs ROOT+GSL and cannot be built in this image; kind=reference runs the "
                "reference's own sources against a GSL/ROOT shim, kind=port the oracle restatement",
    }
    print(json.dumps(line))
    return 0


class Bench:
    """One workload on this rank's GPU: context, plug-in inputs, and the timed legs."""

    def __init__(self, workload, rank, local, world, dev):
        import torch

        from upcgen_b200 import capi
        from upcgen_b200.config import named_config
        self.torch, self.capi = torch, capi
        self.workload, self.rank, self.local, self.world, self.dev = workload, rank, local, world, dev
        self.P = P = named_config(workload)
        self.gpu = capi.UpcGpu(P, local)
        self.ext_stream = torch.cuda.ExternalStream(self.gpu.stream_handle(), device=dev)
        self.n_cells = P.nm * P.ny
        self.pol, self.bk = int(P.use_pol), int(P.breakup_mode > 1)
        # host plug-in values (elementary sigma(m)), computed once: they are inputs of the step
        # (cfg3, light-by-light with USE_POLARIZED_CS 1, is defined at the lumi-table level only -- SURVEY Q5: the
        # reference's fold multiplies by LbyL's identically-zero polarised sigma -- so its step has no fold)
        self.fold = not (self.pol and P.proc_id in (22, 111))
        if not self.fold:
            self.sig = {}
        elif self.pol:
            self.sig = dict(sig_s=capi.elem_sigma_m(P, 1), sig_p=capi.elem_sigma_m(P, 2))
        else:
            self.sig = dict(sig_m=capi.elem_sigma_m(P, 0))
        # N > 1: the exchange is folded into the cell kernel (stores into every rank's table through CUDA IPC
        # mappings) unless the devices cannot map each other or UPCGPU_EXCHANGE=nccl asks for the all-gather
        self.peer = False
        if world > 1 and os.environ.get("UPCGPU_EXCHANGE", "peer") != "nccl":
            from upcgen_b200 import dist as udist
            self.peer = udist.setup_peer_exchange(self.gpu, rank, world)

    def fill(self):
        from upcgen_b200 import dist as udist
        if self.peer:
            udist.fill_lumi_peers(self.gpu, self.rank, self.world, self.dev)
        else:
            udist.fill_lumi_distributed(self.gpu, self.rank, self.world, self.dev)

    def close(self):
        self.gpu.close()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        import torch.distributed as dist
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device(self):
        from upcgen_b200 import dist as udist
        gpu = self.gpu
        gpu.invalidate_tables()
        gpu.prepare_tables()
        self.fill()
        if self.fold:
            gpu.fold_sigma(download=False, **self.sig)   # the step's one host wait
        else:
            gpu.fill_stats()                             # no fold: collect the queued fill

    def time_device(self, steps, warmup, l2_flush, clocks=None, want_clocks=False):
        """Device-resident step: CUDA events on the library's stream, L2 flushed between steps, max over ranks."""
        import numpy as np
        torch, gpu = self.torch, self.gpu
        if clocks:
            clocks.start()
        for _ in range(warmup):
            self.step_device()
        stage = {"ms_tables": 0.0, "ms_flux": 0.0, "ms_qags": 0.0, "ms_cells": 0.0}
        ms_steps = []
        launches0 = gpu.launch_count()
        self.barrier()
        if clocks:
            clocks.mark_begin()
        for _ in range(steps):
            l2_flush.fill_(1)                 # flush L2 between timed iterations (126 MB L2 < 256 MiB)
            self.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.ext_stream)
            self.step_device()
            e1.record(self.ext_stream)
            self.barrier()
            ms_steps.append(e0.elapsed_time(e1))
            st = gpu.fill_stats()
            for k in stage:
                stage[k] += st[k] / steps
        launches = gpu.launch_count() - launches0
        ms = self.max_over_ranks(float(np.mean(ms_steps)))
        self.ms_steps = [round(x, 4) for x in ms_steps]   # this rank's timed steps, one by one
        stats = gpu.fill_stats()
        clk = None
        if want_clocks:
            # nvidia-smi needs a few hundred ms to deliver its first line and samples every 100 ms: when the timed steps
            # are over before two samples have arrived (cfg2 on 8 GPUs: 5 x 2.7 ms) the same step is repeated, untimed,
            # until they have (at most 4 s).  ONE sampler per job (rank 0, all GPUs of the job in one nvidia-smi
            # process: eight pollers beside eight ranks perturbed the steps they were watching); every rank takes the
            # same decision -- the fill is collective --, so the "still waiting" flag is reduced over the ranks.
            n_extra, t_wait = 0, time.perf_counter()
            while True:
                waiting = clocks is not None and clocks.count() < 2 and time.perf_counter() - t_wait < 4.0
                if self.max_over_ranks(1.0 if waiting else 0.0) == 0.0:
                    break
                for _ in range(10):
                    self.step_device()
                n_extra += 10
            if clocks is not None:
                clk = clocks.stop()
                clk["window"] = ("the timed steps" if n_extra == 0 else
                                 f"the timed steps + {n_extra} untimed repetitions of the same step")
        return ms, stage, clk, launches, stats

    def time_e2e(self, steps, l2_flush):
        """The same step through the host-buffer C-ABI calls, pinned host buffers, copies inside the timed region."""
        import ctypes as C

        import numpy as np

        from upcgen_b200 import dist as udist
        torch, gpu, P, pol, fold, sig = self.torch, self.gpu, self.P, self.pol, self.fold, self.sig
        n_cells, world = self.n_cells, self.world
        n_tab = 2 if pol else 1
        d2h = int(n_tab * n_cells * 8 + (n_cells * 8 * (2 if pol else 1) + 8 if fold else 0))
        h2d = int(sum(np.asarray(v).size * 8 for v in sig.values()))
        host_lumi = [torch.empty((P.nm, P.ny), dtype=torch.float64).pin_memory() for _ in range(n_tab)]
        host_cs = torch.empty((P.ny, P.nm), dtype=torch.float64).pin_memory()
        host_ratio = torch.empty((P.ny, P.nm), dtype=torch.float64).pin_memory() if pol else None
        host_sig = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in sig.items()}
        tot = C.c_double()

        def vp(t):
            return None if t is None else C.c_void_p(t.data_ptr())

        def fold_host():
            if pol:
                gpu._chk(gpu.L.upcgpu_fold_sigma(gpu.h, None, vp(host_sig["sig_s"]), vp(host_sig["sig_p"]),
                                                 vp(host_cs), vp(host_ratio), C.byref(tot)))
            else:
                gpu._chk(gpu.L.upcgpu_fold_sigma(gpu.h, vp(host_sig["sig_m"]), None, None, vp(host_cs), None,
                                                 C.byref(tot)))
            return tot.value

        if world == 1:
            def step():
                gpu.invalidate_tables()
                if pol:
                    gpu._chk(gpu.L.upcgpu_fill_lumi(gpu.h, None, vp(host_lumi[0]), vp(host_lumi[1])))
                else:
                    gpu._chk(gpu.L.upcgpu_fill_lumi(gpu.h, vp(host_lumi[0]), None, None))
                return fold_host() if fold else 0.0
            api = ("upcgpu_fill_lumi" + (" + upcgpu_fold_sigma" if fold else "")
                   + " (include/upcgpu.h), pinned host buffers")
        else:
            # N > 1: the sharded fill, the NCCL all-gather, then on EVERY rank the read-back of the lumi table(s)
            # (upcgpu_lumi_download) and the fold with its sigma table read-back (upcgpu_fold_sigma), host buffers
            kinds = (1, 2) if pol else (0,)

            def step():
                gpu.invalidate_tables()
                gpu.prepare_tables()
                self.fill()
                for j, which in enumerate(kinds):
                    gpu._chk(gpu.L.upcgpu_lumi_download(gpu.h, which, vp(host_lumi[j])))
                return fold_host() if fold else 0.0
            api = (("per rank: upcgpu_fill_lumi_shard_peers (cell kernel stores into every rank's table) + upcgpu_lumi_download"
                    if self.peer else
                    "per rank: upcgpu_fill_lumi_shard + NCCL all-gather + upcgpu_lumi_unpack + upcgpu_lumi_download")
                   + (" + upcgpu_fold_sigma" if fold else "") + " (pinned host buffers; bytes are per rank; max over ranks)")
        tot_mb = step()
        dts = []
        for _ in range(steps):
            l2_flush.fill_(1)
            self.barrier()
            t0 = time.perf_counter()
            tot_mb = step()
            torch.cuda.synchronize(self.dev)
            dts.append(time.perf_counter() - t0)
        dt = self.max_over_ranks(float(np.mean(dts)))
        return {"value": n_cells / dt, "unit": "cells/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "api": api, "total_cross_section_mb": tot_mb}

    def time_events(self, n_total, e2e_cap=1 << 18):
        """Event stage: S1 (sampler build) and E1-E5 on n_total candidates split over the ranks by Philox counter
        ranges (no collective); needs the folded sigma table of this context on the device."""
        import ctypes as C
        torch, gpu, P, capi, world, rank = self.torch, self.gpu, self.P, self.capi, self.world, self.rank
        cszm = None if P.ignore_csz else capi.elem_cs_zm(P, 0)
        gpu.sampler_build(cszm=cszm)          # warm-up (allocations)
        torch.cuda.synchronize(self.dev)
        ms_sampler = None
        for _ in range(3):                    # wall clock around a host call: the best of three
            t0 = time.perf_counter()
            gpu.sampler_build(cszm=cszm)      # S1: the CDFs of the (y, m) table and of the nm z tables
            torch.cuda.synchronize(self.dev)
            dt = (time.perf_counter() - t0) * 1e3
            ms_sampler = dt if ms_sampler is None else min(ms_sampler, dt)
        n_ev = max(n_total // world, 1 << 14)
        first = rank * n_ev
        gpu.generate_device(12345, first, n_ev)  # warm-up at full size: the scratch buffers grow on demand
        self.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(self.ext_stream)
        acc = gpu.generate_device(12345, first, n_ev)
        e1.record(self.ext_stream)
        self.barrier()
        ms_ev = self.max_over_ranks(e0.elapsed_time(e1))
        acc_all = acc
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([float(acc)], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t)
            acc_all = int(t.item())
        ev = {"events_per_s": world * n_ev / (ms_ev * 1e-3), "candidates": world * n_ev, "accepted": int(acc_all),
              "ms": ms_ev, "sharding": "Philox counter ranges, no collective", "sampler_build_ms": ms_sampler}
        # the same through upcgpu_generate with pinned host buffers for every output array (per rank; max over ranks)
        n_small = min(n_ev, e2e_cap)
        mp = gpu.particles_per_event()     # particle slots per candidate: upcgpu_generate_packed (2 pairs, 3 ALP + decay)
        hb = {"npart": torch.empty(n_small, dtype=torch.int32).pin_memory(),
              "pdg": torch.empty((n_small, mp), dtype=torch.int32).pin_memory(),
              "status": torch.empty((n_small, mp), dtype=torch.int32).pin_memory(),
              "mother": torch.empty((n_small, mp), dtype=torch.int32).pin_memory(),
              "p4": torch.empty((n_small, mp, 4), dtype=torch.float64).pin_memory()}
        nacc = C.c_uint64()

        def gen_e2e():
            gpu._chk(gpu.L.upcgpu_generate_packed(gpu.h, 12345, first, n_small, mp, C.c_void_p(hb["npart"].data_ptr()),
                                                  C.c_void_p(hb["pdg"].data_ptr()), C.c_void_p(hb["status"].data_ptr()),
                                                  C.c_void_p(hb["mother"].data_ptr()), C.c_void_p(hb["p4"].data_ptr()),
                                                  None, C.byref(nacc)))

        gen_e2e()
        self.barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            gen_e2e()
        dt = self.max_over_ranks((time.perf_counter() - t0) / reps)
        ev["e2e_events_per_s"] = world * n_small / dt
        ev["e2e_candidates_per_rank"] = n_small
        ev["e2e_d2h_bytes"] = int(sum(t.numel() * t.element_size() for t in hb.values()))
        ev["e2e_api"] = (f"upcgpu_generate_packed, {mp} particle slots per candidate, pinned host buffers; chunks of 2^21 "
                         "candidates, the copies of one chunk beside the kernels of the next")
        return ev

    def roofline(self, stage, st, peak_tf):
        """Roofline of the dominant kernel.  `achieved` is ALGORITHMIC work (what the reference's
        algorithm does per unit, no shortcut deducted) over the kernel's measured time.  The kernels do less than that:
        J1 of the rows on the common b grid comes from a table, cells above ny/2 are mirror images, far (b >= 20 fm)
        pairs are a closed sum.  `executed` repeats the figure with only the work the kernel really performed (from its
        own counters), which is what the FP64 pipe sees; the ncu pipe-activity figure of the committed capture stands
        beside it (quoted from profiles/, not measured in this run)."""
        from upcgen_b200 import dist as udist
        P, pol, bk = self.P, self.pol, self.bk
        ms_head = st.get("ms_qags_head", 0.0)
        executed = None
        if st["qags_evals"] > 0 and stage["ms_qags"] >= stage["ms_cells"]:
            if ms_head > 0.5 * stage["ms_qags"]:
                work = FLOP_PER_QAGS_EVAL * st["qags_head_evals"]
                t_k = ms_head * 1e-3
                kern = "k_flux_qags_head"
                units = f"{st['qags_head_evals']} integrand evaluations x {FLOP_PER_QAGS_EVAL:.0f} flop"
                ex = st["qags_head_evals"] - st["qags_table_evals"]
                executed = {"flop": FLOP_PER_QAGS_EVAL * ex, "what": f"{ex} evaluations with J1 computed "
                            f"({st['qags_table_evals']} took J1 from the common-grid table: a multiplication each)"}
            else:
                work = FLOP_PER_QAGS_EVAL * (st["qags_evals"] - st["qags_head_evals"])
                t_k = (stage["ms_qags"] - ms_head) * 1e-3
                kern = "k_flux_qags_rows"
                units = f"{st['qags_evals'] - st['qags_head_evals']} integrand evaluations x {FLOP_PER_QAGS_EVAL:.0f} flop"
        else:
            cells_rank = len(udist.cyclic_rows(P.nm, self.rank, self.world)) * P.ny
            work = FLOP_CELL[(pol, bk)] * cells_rank
            t_k = stage["ms_cells"] * 1e-3
            kern = "k_cells"
            units = f"{cells_rank} cells x {FLOP_CELL[(pol, bk)]:.3g} flop"
            # per evaluated (b1,b2) pair: 5 phi x (15, +10 with breakup, +2 polarised) + 5; per evaluated cell: rows + fluxes
            per_triplet = 15 + (10 if bk else 0) + (2 if pol else 0)
            ex_flop = st["band_pairs"] * (5 * per_triplet + 5) + st["cells_evaluated"] * (360 + 120 * 120 * 2)
            executed = {"flop": float(ex_flop), "what": f"{st['band_pairs']} (b1,b2) pairs evaluated point by point in "
                        f"{st['cells_evaluated']} cells (the other pairs are a closed sum, the other cells mirror images)"}
        achieved = work / t_k / 1e12 if t_k > 0 else 0.0
        if executed is not None and t_k > 0:
            executed["tflops"] = executed["flop"] / t_k / 1e12
            executed["frac"] = executed["tflops"] / peak_tf if peak_tf else None
        ncu = NCU_QUOTED.get((self.workload, kern)) if self.world == 1 else None
        return {"bound": "fp64", "kernel": kern, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None,
                "traffic": ncu["traffic"] if ncu else None,
                "traffic_source": ncu["source"] + " (ncu --set full capture of this kernel, quoted; not measured in this run)"
                if ncu else None,
                "peak_source": "measured live: upcgpu_fp64_peak DFMA loop (MEASURED_PEAKS.json has no FP64 figure; "
                               "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)",
                "algorithmic_work": units, "kernel_ms": t_k * 1e3, "executed": executed,
                "ncu_fp64_pipe_active_pct": ncu["fp64_pipe_pct"] if ncu else None,
                "qags_stage": {"ms": stage["ms_qags"], "ms_head": ms_head, "evals": st["qags_evals"],
                               "tflops": FLOP_PER_QAGS_EVAL * st["qags_evals"] / (stage["ms_qags"] * 1e-3) / 1e12
                               if stage["ms_qags"] > 0 else None}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOAD_TEXT))
    ap.add_argument("--events", type=int, default=None, help="total candidates of the event-stage timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the cfg4 fill leg and the cfg5 10^7-event leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch

    from upcgen_b200 import dist as udist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    rank, local, world = udist.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ---- the headline workload (cfg2 unless --workload says otherwise) -------------------------------
    B = Bench(args.workload, rank, local, world, dev)
    P, gpu = B.P, B.gpu
    for _ in range(2):
        B.step_device()                       # first touch: allocations, the integral-count cache
    peak_tf, _ = gpu.fp64_peak(400000)
    ms, stage, clk, launches, st = B.time_device(args.steps, args.warmup, l2_flush,
                                                  ClockSampler(list(range(world))) if rank == 0 else None, want_clocks=True)
    e2e = B.time_e2e(args.steps, l2_flush)
    events = None
    if args.workload in ("cfg1", "cfg2", "cfg5"):
        events = B.time_events(args.events or min(P.n_events, 1 << 20))
    roofline = B.roofline(stage, st, peak_tf)
    device_name = gpu.device_name()
    peer_exchange = B.peer
    work = {"qags_integrals": st["qags_integrals"], "qags_evals": st["qags_evals"],
            "band_pairs": st["band_pairs"], "flux_rows": st["flux_rows"]}
    B.close()

    # ---- CPU baseline beside it (rank 0, N = 1, bounded sample) --------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cstep, nc, kind, sample = make_cpu_leg(args.workload, cores)
        cstep()  # warm caches/threads
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0 and reps < 50:
            cstep()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": nc / dt, "unit": "cells/s", "cores": cores, "kind": kind,
               "sample": sample + f"; {reps} repetitions"}

    # ---- the north-star target beside it: cfg4 (10001 x 1201 cells) fill and cfg5 (10^7 events) ------
    cfg4 = ev5 = None
    if args.workload == "cfg2" and not args.no_legs:
        B4 = Bench("cfg4", rank, local, world, dev)
        B4.step_device()                      # first touch
        k4 = max(1, min(args.steps, 2))
        ms4, stage4, _, launches4, st4 = B4.time_device(k4, 1, l2_flush)
        e2e4 = B4.time_e2e(1, l2_flush)
        cfg4 = {"workload": WORKLOAD_TEXT["cfg4"], "cells": B4.n_cells, "grid": [B4.P.nm, B4.P.ny], "steps": k4,
                "warmup": 2, "ms_per_step": ms4, "cells_per_s": B4.n_cells / (ms4 * 1e-3), "stage_ms": stage4,
                "e2e": e2e4, "gpu_launches": int(launches4), "roofline": B4.roofline(stage4, st4, peak_tf),
                "work": {"qags_integrals": st4["qags_integrals"], "qags_evals": st4["qags_evals"],
                         "band_pairs": st4["band_pairs"], "flux_rows": st4["flux_rows"]}}
        B4.close()
        B5 = Bench("cfg5", rank, local, world, dev)
        B5.step_device()                      # tables, lumi table, fold: the sigma table the samplers are built from
        ev5 = B5.time_events(10_000_000, e2e_cap=1 << 23)
        ev5["workload"] = WORKLOAD_TEXT["cfg5"]
        B5.close()

    if rank == 0:
        line = {
            "metric": "lumi_cells_per_s", "value": B.n_cells / (ms * 1e-3), "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.workload, P, world, peer_exchange),
            "sigma_table_ms": ms, "ms_steps_rank0": B.ms_steps,
            "stage_ms": stage, "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "events": events,
            "work": work, "cfg4": cfg4, "events_cfg5": ev5,
            "device": device_name,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
