#!/bin/bash
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest.txt
echo "== bench 16M"; timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m.json 2> $out/bench16m.err; python scratch/show_bench.py $out/bench16m.json; tail -3 $out/bench16m.err
if [ "$2" != "no64" ]; then
echo "== bench 64M"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/bench64m.json 2> $out/bench64m.err; python scratch/show_bench.py $out/bench64m.json; tail -3 $out/bench64m.err
fi
if [ "$3" != "" ]; then
echo "== ncu full 4M $3"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" --launch-skip 1 -c 1 -o $out/ncu4m -f python bench.py --particles 4194304 --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu4m.log 2>&1; tail -2 $out/ncu4m.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu4m.ncu-rep > $out/ncu4m_summary.txt 2>&1
fi
