#!/bin/bash
# launch list (per-kernel durations) at 16M for kernels matching regex $2
tag=$1; rx=$2
out=gpurun_out/$tag
mkdir -p $out
export GSPH_GRAV_LIST_MB=100000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$rx" --csv --log-file $out/launch.csv python bench.py --particles 16777216 --steps 1 --warmup 0 --no-cpu-baseline --no-resident > $out/launch.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open('$out/launch.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ix={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows:
    if len(r)==len(hdr) and r[ix['Metric Name']]=='gpu__time_duration.sum':
        k=r[ix['Kernel Name']][:60]; v=float(r[ix['Metric Value']].replace(',',''))/1e6
        agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print('   %-62s n=%3d last %9.3f ms  min %9.3f'%(k,len(v),v[-1],min(v)))
PY
