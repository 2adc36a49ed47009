set -x
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "saved_pooled or fused_and" 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in c3 c4 c5; do timeout 400 python bench.py --config $c --no-cpu-baseline --steps 20 > gpurun_out/bench_${c}_M32.json 2>/dev/null; tail -c 200 gpurun_out/bench_${c}_M32.json | head -c 60; echo; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2>/dev/null; head -c 300 gpurun_out/bench_reference_arm.json; echo
timeout 300 python bench.py --kernel-mode 1 --no-cpu-baseline --steps 5 > gpurun_out/bench_v0_general_kernels.json 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ks_kernel<1>|kp_kernel<1>' -c 4 -o gpurun_out/streaming_bwd -f python tools/stage_times.py --only ep_fwd,bwd_proj,bwd_pool --iters 1 > gpurun_out/ncu_full_bwd.log 2>&1; tail -2 gpurun_out/ncu_full_bwd.log
