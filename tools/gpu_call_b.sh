#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py -m gpu -q -k operand_copies 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv python tools/profile_step.py c2 32 3 > gpurun_out/ncu_ll.log 2>&1
python tools/summarize_launch_list.py gpurun_out/launches_fused.csv 17
