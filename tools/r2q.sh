#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
(python tools/bench_configs.py cfg2_wgbs150 cfg2b_pe_merged cfg3_capture_cx deep150 cfg4_long_nothr cfg4_long_thr 2>&1 | grep config) > gpurun_out/r2q.txt
cut -c1-100 gpurun_out/r2q.txt; python - <<'P'
import json
for l in open('gpurun_out/r2q.txt'):
    d=json.loads(l); print(d['config'], d['kernel'][:24], 'total', d['total_ms'], 'thr', d['threshold_ms'], 'win', d['window_ms'], 'gather', d['gather_ms'], 'frac_win', d['frac_window'], 'rows', d['rows'], 'pass', d['n_pass'])
P
