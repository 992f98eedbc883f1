mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ev.csv python tools/profile_step.py cfg2 0 1 1048576 > gpurun_out/launches_ev.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ev5.csv python tools/profile_step.py cfg5 0 1 1048576 > gpurun_out/launches_ev5.log 2>&1
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
