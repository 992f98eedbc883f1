#!/bin/bash
# ncu --set full capture of one kernel of a cfg step:  bash tools/gpu_prof.sh <kernel-regex> [cfg] [out-name]
K=${1:-k_flux_qags_head}; CFG=${2:-cfg2}; OUT=${3:-prof}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/$OUT \
    python tools/profile_step.py $CFG 1 1 > gpurun_out/$OUT.log 2>&1
tail -2 gpurun_out/$OUT.log
