#!/bin/bash
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p $out
export GSPH_GRAV_LIST_MB=100000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gravity_walk_kernel|gravity_consume_kernel' --launch-skip 2 -c 2 -o $out/ncu_grav -f python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/ncu_grav.log 2>&1
tail -2 $out/ncu_grav.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu_grav.ncu-rep > $out/ncu_grav_summary.txt 2>&1
cat $out/ncu_grav_summary.txt
