#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the launch list and one ncu --set full capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [kernel-regex]'
K=${1:-k_flux_qags_head}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( time timeout 400 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py cfg2 1 1 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof \
    python tools/profile_step.py cfg2 1 1 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench.json
