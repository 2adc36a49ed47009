#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
run() { tag=$1; shift; "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - gpurun_out/bench_$tag.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f launches %s loss %.5f" % (j["ms_per_step"], j["launches_per_step"], j["mean_loss"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run c2_pdl python bench.py --steps 100 --warmup 5 --no-cpu-baseline
EP_PDL=0 run c2_nopdl python bench.py --steps 100 --warmup 5 --no-cpu-baseline
run c2_pdl2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline
EP_PDL=0 run c2_nopdl2 python bench.py --steps 100 --warmup 5 --no-cpu-baseline
run c2_pdl_nograph python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-graph
run c2_M8 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --queries 8
run c3 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --config c3
