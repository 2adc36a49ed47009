#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r02b_bench_c2_M32.json 2> gpurun_out/bench_c2.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02b_bench_c2_M32.json").read().strip().splitlines()[-1])
print("c2: ms/step %.4f launches %s step_frac %.3f e2e %.4g cpu %.4g" % (j["ms_per_step"], j["launches_per_step"], j["step_roofline_frac"], j["e2e"]["value"], j["cpu_baseline"]["value"]))
PY
