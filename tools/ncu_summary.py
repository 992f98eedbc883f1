#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers the design discussion uses.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.avg",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:110])
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k]:>20s} {u[k]}")
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[h]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  warp stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    sass = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    if len(rows) > 2:
        ix = {h: i for i, h in enumerate(rows[1])}
        c, tot, n = Counter(), 0, 0
        for r in rows[2:]:
            try:
                inst = int(r[ix["Instructions Executed"]])
            except (ValueError, IndexError):
                continue
            src = r[ix["Source"]].strip()
            op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
            c[op] += inst
            tot += inst
            n += 1
        print(f"  SASS: {n} instructions (~{n * 16 / 1024:.0f} KB); executed mix: " +
              " ".join(f"{op}={100 * v / tot:.1f}%" for op, v in c.most_common(12)))


if __name__ == "__main__":
    main(sys.argv[1])
