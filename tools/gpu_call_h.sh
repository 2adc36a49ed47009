#!/bin/bash
mkdir -p gpurun_out
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
b2 pdl
EP_PDL=0 b2 nopdl
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1 GPU (GPU 0 of the pair) pdl late-trigger: ms/step %.4f' % j['ms_per_step'])"
