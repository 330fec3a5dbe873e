#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -6 gpurun_out/r2v_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2v_bench.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], d['e2e']['h2d_bytes_per_step'], 'cpu ok', d['cpu_baseline'].get('gpu_matches_bit_exact'))
P
EA_TRACE=1 python tools/e2e_breakdown.py 2>&1 | tail -22 > gpurun_out/r2v_trace.txt; cat gpurun_out/r2v_trace.txt
