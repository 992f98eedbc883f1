#!/usr/bin/env python
"""Prints the numbers of bench.py JSON lines side by side:  python tools/bench_table.py profiles/r02_bench_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        txt = open(f).read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    s = d["stage_ms"]
    ev = d.get("events") or {"events_per_s": 0, "e2e_events_per_s": 0, "sampler_build_ms": 0}
    print(f"{f}: N={d['n_gpus']} step {d['ms_per_step']:.3f} ms (tables {s['ms_tables']:.2f} flux {s['ms_flux']:.2f} "
          f"cells {s['ms_cells']:.2f}) e2e {d['e2e']['ms_per_step']:.3f} ms; events {ev['events_per_s'] / 1e6:.0f} M/s "
          f"(e2e {ev['e2e_events_per_s'] / 1e6:.0f}), sampler {ev['sampler_build_ms']:.1f} ms")
    if d.get("cfg4"):
        c = d["cfg4"]
        e = d["events_cfg5"]
        print(f"    cfg4 {c['ms_per_step']:.1f} ms (flux {c['stage_ms']['ms_flux']:.1f} cells {c['stage_ms']['ms_cells']:.1f}) "
              f"e2e {c['e2e']['ms_per_step']:.1f} ms; cfg5 1e7 events {e['events_per_s'] / 1e6:.0f} M/s (e2e {e['e2e_events_per_s'] / 1e6:.0f})")
