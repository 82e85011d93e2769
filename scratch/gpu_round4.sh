#!/bin/bash
tag=${1:-r}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest.txt
echo "== bench 16M"; timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m.json 2> $out/bench16m.err; python scratch/show_bench.py $out/bench16m.json; tail -3 $out/bench16m.err
echo "== bench 64M"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/bench64m.json 2> $out/bench64m.err; python scratch/show_bench.py $out/bench64m.json; tail -3 $out/bench64m.err
echo "== ncu tree kernels 16M"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'partition_v2_kernel|count_right_kernel' --launch-skip 20 -c 2 -o $out/ncu_tree -f python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu_tree.log 2>&1; tail -2 $out/ncu_tree.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu_tree.ncu-rep > $out/ncu_tree_summary.txt 2>&1
