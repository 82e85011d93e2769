#!/bin/bash
tag=${1:-r}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/pytest.txt
echo "== bench 16M"; timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m.json 2> $out/bench16m.err; python scratch/show_bench.py $out/bench16m.json; tail -3 $out/bench16m.err
echo "== bench 16M no pretest"; GSPH_GRAV_PRETEST=0 timeout 600 python bench.py --particles 16777216 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench16m_nopre.json 2> $out/bench16m_nopre.err; python scratch/show_bench.py $out/bench16m_nopre.json
echo "== launch list 16M"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches16m.csv python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline > $out/launches16m.log 2>&1; wc -l $out/launches16m.csv
echo "== ncu full 4M gravity"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gravity_kernel' -c 1 -o $out/ncu4m_grav -f python bench.py --particles 4194304 --steps 1 --warmup 0 --no-cpu-baseline > $out/ncu4m.log 2>&1; tail -2 $out/ncu4m.log | cut -c1-200
python profiles/ncu_summary.py $out/ncu4m_grav.ncu-rep > $out/ncu4m_grav_summary.txt 2>&1
