#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
for w in khi bossbodenheimer glass; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > $out/bench_$w.json 2> $out/bench_$w.err; python scratch/show_bench.py $out/bench_$w.json | head -2 | cut -c1-400; tail -2 $out/bench_$w.err | cut -c1-300
done
