#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== launch list (default command, 64M)"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches64m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/launches64m.log 2>&1; wc -l $out/launches64m.csv
