#!/bin/bash
# usage: ab_env.sh "ENV1=a ENV2=b" lib1 lib2 ...   - like ab.sh, with extra environment for every run
envs=$1; shift
for v in "$@"; do
  env $envs EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/$v.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', '$envs', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'rows', d['config']['report_rows'], 'pass', d['config']['reads_passing'], d['roofline']['kernel'][:30])"
done
