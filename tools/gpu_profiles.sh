#!/bin/bash
# launch list + ncu --set full captures of the three main kernels of a cfg2 step (profiles/ are summarised from these)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_step.py cfg2 1 1 > gpurun_out/launches.log 2>&1
for K in "k_flux_qags_head<11" "k_flux_qags_head<16" k_flux_qags_rows k_cells; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 1 -c 1 -f -o "gpurun_out/prof_${K//[<]/_}" \
      python tools/profile_step.py cfg2 1 1 > "gpurun_out/prof_${K//[<]/_}.log" 2>&1
done
timeout 300 python bench.py --workload cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 300 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 300 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
ls -la gpurun_out | tail -20
