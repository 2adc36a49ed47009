#!/bin/bash
# Round-2 final single-GPU evidence (after the operand-copy fusion, the side-branch weight gradient and PDL):
# bench lines, the ncu launch list of the bench command, a memcheck run.  Writes into gpurun_out/.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02b_bench_c2_M32.json 2> gpurun_out/bench_c2.err; tail -c 200 gpurun_out/r02b_bench_c2_M32.json; echo
timeout 300 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r02b_bench_c2_M32_200steps.json 2>/dev/null
timeout 300 python bench.py --queries 8 --no-cpu-baseline > gpurun_out/r02b_bench_c2_M8.json 2>/dev/null
for c in c3 c4 c5; do timeout 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/r02b_bench_${c}_M32_1gpu.json 2>/dev/null; done
EP_FUSE_OPERANDS=0 EP_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_c2_M32_selfcontained_calls.json 2>/dev/null
for f in gpurun_out/r02b_bench_*.json; do python - $f <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f launches %s step_frac %.3f roofline %s %.3f" % (j["ms_per_step"], j["launches_per_step"], j["step_roofline_frac"], j["roofline"].get("kernel"), j["roofline"]["frac"] or 0))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_ncu_launch_list.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
python tools/summarize_launch_list.py gpurun_out/r02b_ncu_launch_list.csv 17 auto > gpurun_out/r02b_ncu_launch_list_summary.txt; tail -3 gpurun_out/r02b_ncu_launch_list_summary.txt
timeout 600 compute-sanitizer --tool memcheck python tools/dev_memcheck_step.py > gpurun_out/r02b_compute_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02b_compute_sanitizer_memcheck.log
