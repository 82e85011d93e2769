#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 3 --warmup 3 > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -1 $out/bench_$2.err | cut -c1-200; }
run "GSPH_X=0" default
echo "== pytest gravity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mainloop.py -m gpu -x -q -k "grav or sphere or cloud or stack or overflow or stars" 2>&1 | tail -3
