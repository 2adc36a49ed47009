#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/b2.err | grep '^{' > gpurun_out/r02b_bench_c2_M32_2gpu.json
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02b_bench_c2_M32_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs: ms/step %.4f launches %s replica diff %s comm_sms %s e2e %.4g" % (j["ms_per_step"], j["launches_per_step"], j["run"]["replica_max_abs_diff"], j["run"]["comm_sms"], j["e2e"]["value"]))
PY
