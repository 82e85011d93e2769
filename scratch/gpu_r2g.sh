#!/bin/bash
tag=${1:-r2g}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $out/pytest.txt
run() { echo "== bench 64M $1"; env $1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-resident > $out/bench_$2.json 2> $out/bench_$2.err; python scratch/show_bench.py $out/bench_$2.json | head -1; tail -2 $out/bench_$2.err; }
run "GSPH_X=0" default
run "GSPH_GRAV_LIST_MB=100000" onebatch
export GSPH_GRAV_LIST_MB=100000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gravity' --csv --log-file $out/launch.csv python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/launch.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('$out/launch.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ix={h:i for i,h in enumerate(hdr)}
for r in rows[-12:]:
    if len(r)==len(hdr) and r[ix['Metric Name']]=='gpu__time_duration.sum':
        print('   %-44s %10.3f ms'%(r[ix['Kernel Name']][:44], float(r[ix['Metric Value']].replace(',',''))/1e6))
PY
