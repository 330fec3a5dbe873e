#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
for e in "X=1" "EA_PL_UNORDERED=1"; do
env $e python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$e', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,3), 'frac', round(d['roofline']['frac'],4), 'rows', d['config']['report_rows'], 'pass', d['config']['reads_passing'], 'launches', d['gpu_launches'])"
done > gpurun_out/r2h_ord.txt 2>&1
cat gpurun_out/r2h_ord.txt
