#!/bin/bash
tag=${1:-r2e}
out=gpurun_out/$tag
mkdir -p $out
for mode in onebatch default; do
  if [ $mode = onebatch ]; then export GSPH_GRAV_LIST_MB=100000; else unset GSPH_GRAV_LIST_MB; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gravity' --csv --log-file $out/launch_$mode.csv python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/launch_$mode.log 2>&1
  grep -E "gravity" $out/launch_$mode.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-160 | tail -12
done
