#!/bin/bash
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest.txt
echo "== rank phases 64M world 8"; timeout 600 python scratch/rank_phases.py 67108864 8 2>&1 | tail -25 | tee $out/phases8.txt
echo "== bench 64M"; timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > $out/bench64m.json 2> $out/bench64m.err; python scratch/show_bench.py $out/bench64m.json; tail -3 $out/bench64m.err
