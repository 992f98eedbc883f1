#!/bin/bash
# round-end check on ONE GPU: the GPU suite, the default bench (both arms), the head kernel's ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_flux_qags_head" -s 2 -c 1 -f -o gpurun_out/prof_head_pass1_cfg2 \
    python tools/profile_step.py cfg2 1 1 > gpurun_out/prof_head_pass1_cfg2.log 2>&1; echo "ncu head rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_cells" -s 1 -c 1 -f -o gpurun_out/prof_k_cells8_cfg1 \
    python tools/profile_step.py cfg1 1 1 > gpurun_out/prof_k_cells8_cfg1.log 2>&1; echo "ncu cells rc=$?"
ls -la gpurun_out | tail -8
