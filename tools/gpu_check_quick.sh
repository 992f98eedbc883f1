#!/bin/bash
# GPU tests + one shard's launch list (what a rank of an 8-GPU run executes)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_shard8.csv python tools/profile_shard.py cfg2 8 2 > gpurun_out/launches_shard8.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 3 gpurun_out/pytest.log; cut -c1-330 gpurun_out/bench.json
