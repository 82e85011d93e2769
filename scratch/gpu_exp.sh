#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python scratch/rank_phases.py 67108864 8 > $out/phases.txt 2>&1; tail -40 $out/phases.txt
