#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_X=0" base
run "GSPH_X=1" base2
