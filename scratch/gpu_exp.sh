#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
for r in 0 2 5; do
timeout 900 python scratch/rank_phases.py 67108864 8 $r > $out/phases_$r.txt 2>&1; grep -E "rank|density_pass|forces_begin|forces_finish|sum of" $out/phases_$r.txt
done
