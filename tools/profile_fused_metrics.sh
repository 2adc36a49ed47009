#!/bin/bash
# DRAM bytes / L2 hit rate / duration of the fused kernels for a few settings (ncu metrics pass, no full set)
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second"
IFS=";" read -ra LIST <<< "${FM_ARGS:-c2 32 3 0 0;c2 32 3 0 120;c2 32 3 0 96;c2 32 3 0 74}"; for args in "${LIST[@]}"; do
  tag=$(echo $args | tr ' ' '_')
  ncu --metrics $M --clock-control none -k regex:fused_kernel -s 2 -c 2 --csv --log-file gpurun_out/fm_$tag.csv python tools/profile_step.py $(echo $args) > /dev/null 2>&1
  echo "== $args"; grep -v "^==" gpurun_out/fm_$tag.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(' ', d['Kernel Name'][:24], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
"
done
