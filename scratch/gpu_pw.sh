#!/bin/bash
tag=${1:-pw}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/pytest.txt
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_PREWALK=0" nopre
run "GSPH_PREWALK=1" pre
