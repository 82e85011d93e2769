#!/bin/bash
tag=${1:-x1}
out=gpurun_out/$tag
mkdir -p $out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/pytest.txt
echo "== bench"; timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; python scratch/show_bench.py $out/bench_default.json | head -1; tail -1 $out/bench_default.err | cut -c1-300
