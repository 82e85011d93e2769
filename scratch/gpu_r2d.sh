#!/bin/bash
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_X=0" default
run "GSPH_GRAV_FUSED=1" fused
run "GSPH_GRAV_WALK_BPS=2 GSPH_GRAV_CONS_BPS=2" w2c2
run "GSPH_GRAV_WALK_BPS=2 GSPH_GRAV_CONS_BPS=1" w2c1
run "GSPH_GRAV_WALK_BPS=1 GSPH_GRAV_CONS_BPS=2" w1c2
run "GSPH_GRAV_LIST_MB=100000" onebatch
