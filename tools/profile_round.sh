set -x
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_c2_M32.json 2> gpurun_out/bench_c2_M32.err; tail -c 300 gpurun_out/bench_c2_M32.json
timeout 300 python bench.py --queries 8 --no-cpu-baseline > gpurun_out/bench_c2_M8.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ks_kernel|kp_kernel' -c 8 -o gpurun_out/streaming -f python tools/stage_times.py --only ep_fwd,bwd_proj,bwd_pool --iters 1 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
