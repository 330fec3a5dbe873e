#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vertical" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --policy 6 2>gpurun_out/r2e_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('vs cfg2', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'rows', d['config']['report_rows'], 'pass', d['config']['reads_passing'], d['roofline']['kernel'][:40])" > gpurun_out/r2e_vs.txt 2>&1
cat gpurun_out/r2e_vs.txt; tail -3 gpurun_out/r2e_bench.err
