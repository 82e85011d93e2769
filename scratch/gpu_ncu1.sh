#!/bin/bash
# ncu --set full of one kernel regex at 16M (one gravity batch); usage: gpu_ncu1.sh tag regex
tag=$1; rx=$2
out=gpurun_out/$tag
mkdir -p $out
export GSPH_GRAV_LIST_MB=100000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 1 -c 1 -o $out/ncu -f python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/ncu.log 2>&1
tail -1 $out/ncu.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu.ncu-rep > $out/ncu_summary.txt 2>&1
cat $out/ncu_summary.txt
