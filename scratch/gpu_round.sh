#!/bin/bash
# one GPU-box round: parity tests, bench lines, ncu capture. Outputs under gpurun_out/<tag>/
tag=${1:-r}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest.txt
if grep -q failed $out/pytest.txt; then
  echo "== pytest (tree v1)"; GSPH_TREE_V1=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $out/pytest_v1.txt
fi
echo "== bench 16M"; timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m.json 2> $out/bench16m.err; python scratch/show_bench.py $out/bench16m.json; tail -3 $out/bench16m.err
echo "== bench 16M tree v1"; GSPH_TREE_V1=1 timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m_v1.json 2> $out/bench16m_v1.err; python scratch/show_bench.py $out/bench16m_v1.json
if [ "$2" != "nolong" ]; then
echo "== ncu full 4M"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'density_kernel|hydro_kernel|gravity_kernel' --launch-skip 1 -c 3 -o $out/ncu4m -f python bench.py --particles 4194304 --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu4m.log 2>&1; tail -2 $out/ncu4m.log | cut -c1-300
python profiles/ncu_summary.py $out/ncu4m.ncu-rep > $out/ncu4m_summary.txt 2>&1
fi
