#!/bin/bash
export EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v8.so
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2r_plain.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2r_launches.csv')))
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[rows.index(hdr)+1:]:
    if len(r)>vi:
        try: d[r[ki][:70]].append(float(r[vi].replace(',','')))
        except: pass
for k,v in d.items(): print('%-72s n=%3d avg %.1f us'%(k,len(v),sum(v)/len(v)/1e3))
P
python tools/bench_configs.py cfg3_capture_cx deep150 2>&1 | grep config | cut -c1-30,200-420
