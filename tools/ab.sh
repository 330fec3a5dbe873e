#!/bin/bash
# run the bench once per A/B library and print the window-kernel time
for v in "$@"; do
  EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/$v.so python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'rows', d['config']['report_rows'], 'pass', d['config']['reads_passing'])"
done
