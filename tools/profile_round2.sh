#!/bin/bash
# Round-2 single-GPU evidence: bench lines of every config, the reference arm, the ncu launch list of a step, a full
# ncu capture of the one-pass kernels and a memcheck run.  Writes into gpurun_out/ (copied to profiles/ by hand).
set -x
timeout 600 python bench.py > gpurun_out/r02_bench_c2_M32.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/r02_bench_c2_M32.json | head -c 120; echo
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r02_bench_c2_M32_200steps.json 2>/dev/null
timeout 300 python bench.py --queries 8 --no-cpu-baseline > gpurun_out/r02_bench_c2_M8.json 2>/dev/null
bash tools/bench_multi.sh 1 "c3 c4 c5"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 2 -c 2 -o gpurun_out/r02_fused_full -f python tools/profile_step.py c2 32 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 compute-sanitizer --tool memcheck python tools/dev_memcheck_step.py > gpurun_out/r02_compute_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer_memcheck.log
