mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_flux_qags_head -s 2 -c 1 -f -o gpurun_out/prof_head_pass1 python tools/profile_step.py cfg2 1 1 > gpurun_out/prof_head_pass1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_flux_qags_head -s 3 -c 1 -f -o gpurun_out/prof_head_pass2 python tools/profile_step.py cfg2 1 1 > gpurun_out/prof_head_pass2.log 2>&1
tail -n 2 gpurun_out/prof_head_pass1.log
