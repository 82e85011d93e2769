#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_X=0" base
echo "== pytest gravity"; timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mainloop.py -m gpu -x -q -k "grav or sphere or cloud or stack or overflow" 2>&1 | tail -4 | tee $out/pytest.txt
