#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python scratch/block_steps_bench.py 16777216 5 32 > $out/block16m.json 2> $out/block16m.err; tail -1 $out/block16m.json | cut -c1-1200; tail -3 $out/block16m.err
timeout 900 python scratch/block_steps_bench.py 67108864 5 16 > $out/block64m.json 2> $out/block64m.err; tail -1 $out/block64m.json | cut -c1-1200; tail -3 $out/block64m.err
