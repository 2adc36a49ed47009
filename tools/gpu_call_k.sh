#!/bin/bash
mkdir -p gpurun_out
timeout 150 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 2 -o gpurun_out/r02b_fused_full -f python tools/profile_step.py c2 32 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/r02b_fused_full.ncu-rep
