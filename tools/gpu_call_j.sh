#!/bin/bash
# ncu metrics of the head's kernels around the one-pass pooling kernels (GEMMs: tensor-pipe activity, L2 / DRAM bytes)
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size,launch__registers_per_thread"
timeout 200 ncu --metrics $M --clock-control none -k regex:'gemm_tf32_kernel|g_ops_kernel|refresh_kernel|bn_fwd_kernel|bn_bwd_kernel|ce_kernel|lars_|reduce_partials|colsum' -s 16 -c 15 --csv --log-file gpurun_out/r02b_ncu_head_kernels_raw.csv python tools/profile_step.py c2 32 3 > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log; wc -l gpurun_out/r02b_ncu_head_kernels_raw.csv
