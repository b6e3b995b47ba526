#!/bin/bash
# DRAM traffic / L2 hit rate of the sweep kernel vs panel size.  Run on the GPU box.
for P in 12 24 32 48 96; do
  for H in 0 1; do
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
        --clock-control none -k regex:sweep_major -s 4 -c 2 --csv \
        python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --option panel_mb=$P --option hint=$H --option lpg=4 --option unroll=1 --option minb=3 \
        2>/dev/null | grep -E "sweep_major" | awk -F'","' -v p=$P -v h=$H '{print "panel",p,"hint",h,$(NF-2),$(NF)}'
  done
done
