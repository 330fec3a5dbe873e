#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -6 gpurun_out/r2x_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], d['e2e']['h2d_bytes_per_step'], 'cpu ok', d['cpu_baseline'].get('gpu_matches_bit_exact'))
P
