#!/bin/bash
# quick gravity iteration: parity subset + per-kernel durations at 16M (one batch)
tag=${1:-g}
out=gpurun_out/$tag
mkdir -p $out
echo "== pytest (gravity subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gravity or sphere or star or halves" 2>&1 | tail -5 | tee $out/pytest.txt
export GSPH_GRAV_LIST_MB=100000 GSPH_DEBUG=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gravity' --csv --log-file $out/launch.csv python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/launch.log 2>&1
grep "gsph:" $out/launch.log | tail -2
python - <<PY
import csv
rows=list(csv.reader(open('$out/launch.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ix={h:i for i,h in enumerate(hdr)}
tot=0
for r in rows[-4:]:
    if len(r)==len(hdr) and r[ix['Metric Name']]=='gpu__time_duration.sum':
        v=float(r[ix['Metric Value']].replace(',',''))/1e6; tot+=v
        print('   %-44s %10.3f ms'%(r[ix['Kernel Name']][:44], v))
print('   total %.3f ms'%tot)
PY
