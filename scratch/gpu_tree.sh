#!/bin/bash
tag=${1:-t1}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/pytest.txt
echo "== racecheck gravity"; timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gravity_forces_match_reference and (fastmono or quad)" > $out/racecheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" $out/racecheck.log | tail -3
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_TREE_BOTTOM=0" nobottom
run "GSPH_TREE_BOTTOM=1" bottom
echo "== rank phases"; timeout 600 python scratch/rank_phases.py 67108864 8 2>&1 | grep -E "shard_begin|shard_finish|sum of|t_sort|t_levels"
