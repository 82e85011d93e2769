#!/bin/bash
# round-2 profile set: default bench line (64M, cpu baseline), reference arm, launch list of the same command,
# ncu --set full of the hot kernels at 16M
tag=${1:-r2p}
out=gpurun_out/$tag
mkdir -p $out
echo "== default bench (64M)"; timeout 1500 python bench.py > $out/bench_default.json 2> $out/bench_default.err; python scratch/show_bench.py $out/bench_default.json; tail -3 $out/bench_default.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 400 $out/bench_reference.json
echo "== launch list (default command, 64M)"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches64m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/launches64m.log 2>&1; wc -l $out/launches64m.csv
echo "== ncu full 16M"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'density_kernel|hydro_kernel|gravity_walk_kernel|gravity_cells_kernel|gravity_direct_kernel|density_walk_kernel|hydro_walk_kernel' --launch-skip 2 -c 16 -o $out/ncu16m -f python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/ncu16m.log 2>&1; tail -2 $out/ncu16m.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu16m.ncu-rep > $out/ncu16m_summary.txt 2>&1
grep -E "=====|gpu__time_duration|dram__bytes|fp64|issue_active" $out/ncu16m_summary.txt
