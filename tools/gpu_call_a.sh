#!/bin/bash
# one GPU call: the GPU test suite, then A/B of the operand-copy fusion on the bench step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 --durations=8 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_fused.json 2> gpurun_out/bench_c2_fused.err; echo "bench rc=$?"
EP_FUSE_OPERANDS=0 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_unfused.json 2> gpurun_out/bench_c2_unfused.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --debug-bits 1073741824 > gpurun_out/bench_c2_fused_dWtn.json 2> gpurun_out/bench_c2_fused_dWtn.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --queries 8 > gpurun_out/bench_c2_M8_fused.json 2> gpurun_out/bench_c2_M8.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --config c3 > gpurun_out/bench_c3_fused.json 2> gpurun_out/bench_c3.err
for f in gpurun_out/bench_c2_fused.json gpurun_out/bench_c2_unfused.json gpurun_out/bench_c2_fused_dWtn.json gpurun_out/bench_c2_M8_fused.json gpurun_out/bench_c3_fused.json; do
  python - "$f" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f launches %s loss %.5f" % (j["ms_per_step"], j["launches_per_step"], j["mean_loss"]), j["kernel_us"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
