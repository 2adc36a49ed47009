#!/bin/bash
# 2-GPU call: multi-GPU parity tests, then A/B of the two-group update under torchrun
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -3
b2() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/b2_$tag.err | grep '^{' > gpurun_out/bench_2gpu_$tag.json
python - gpurun_out/bench_2gpu_$tag.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f launches %s replica diff %s comm_sms %s" % (j["ms_per_step"], j["launches_per_step"], j["run"]["replica_max_abs_diff"], j["run"]["comm_sms"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
b2 two
EP_TWO_GROUP_UPDATE=0 b2 one
b2 two_b
b2 two_k20 --comm-sms 20
tail -3 gpurun_out/b2_two.err
