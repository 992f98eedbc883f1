#!/bin/bash
# round-end check: smoke(), GPU tests with their printed parity figures, the bench line, launch list, ncu captures
mkdir -p gpurun_out
( python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s.log
( timeout 400 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py cfg2 1 1 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_flux_qags_head -s 2 -c 1 -f -o gpurun_out/prof_head_pass1 python tools/profile_step.py cfg2 1 1 > gpurun_out/prof_head_pass1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cells -s 1 -c 1 -f -o gpurun_out/prof_k_cells python tools/profile_step.py cfg2 1 1 > gpurun_out/prof_k_cells.log 2>&1
tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/pytest_s.log; cut -c1-300 gpurun_out/bench.json
