#!/bin/bash
# bench.py for a list of configs on N GPUs of this box: tools/bench_multi.sh N "c2 c3 ..." [extra bench args]
N=$1; CFGS=$2; shift 2
for c in $CFGS; do
  out=gpurun_out/r02_bench_${c}_M32_${N}gpu.json
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --config $c --steps 50 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/bm_${c}_${N}.err | grep '^{' > $out
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --config $c --steps 50 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/bm_${c}_${N}.err | grep '^{' > $out
  fi
  python - <<PY
import json
try:
    d = json.load(open("$out"))
    print("$c", "N=$N", "ms/step %.4f" % d["ms_per_step"], "tok/s %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["run"].get("replica_max_abs_diff"), "comm_sms", d["run"]["comm_sms"])
except Exception as e:
    print("$c N=$N failed:", e)
PY
done
