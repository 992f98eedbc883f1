#!/bin/bash
# launch lists + ncu --set full captures of the main kernels (profiles/ are summarised from these with tools/ncu_summary.py)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python tools/profile_step.py cfg2 1 1 1000000 > gpurun_out/launches_cfg2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5.csv \
    python tools/profile_step.py cfg5 0 1 8000000 > gpurun_out/launches_cfg5.log 2>&1
for K in "k_flux_qags_head<11" k_cells; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 1 -c 1 -f -o "gpurun_out/prof_${K//[<]/_}_cfg2" \
      python tools/profile_step.py cfg2 1 1 > "gpurun_out/prof_${K//[<]/_}_cfg2.log" 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_cells" -s 1 -c 1 -f -o gpurun_out/prof_k_cells_cfg1 \
    python tools/profile_step.py cfg1 1 1 > gpurun_out/prof_k_cells_cfg1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_pt_serve_keys" -c 1 -f -o gpurun_out/prof_k_pt_serve_keys_cfg2 \
    python tools/profile_step.py cfg2 0 1 1000000 > gpurun_out/prof_k_pt_serve_keys_cfg2.log 2>&1
ls -la gpurun_out | tail -20
