#!/bin/bash
# usage: abcfg.sh "cfgA cfgB" lib1 lib2 ...
cfgs=$1; shift
for v in "$@"; do
  echo "== $v"; EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/$v.so python tools/bench_configs.py $cfgs 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print(d['config'], d['kernel'][:40], 'window_ms', d['window_ms'], 'frac', d['frac_window'], 'rows', d['rows'], 'pass', d['n_pass'])"
done
