#!/bin/bash
# A/B of the stand-alone threshold pass: bench's two-pass variant times it on the resident 50 M reads
for v in "$@"; do
  EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/$v.so python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d['two_pass_variant']; print('$v', 'threshold_ms', t['threshold_ms'], 'window_ms', t['window_ms'], 'n_pass', t['n_pass'], 'fused n_pass', d['config']['reads_passing'])"
done
