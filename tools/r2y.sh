#!/bin/bash
EA_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; grep -B1 -A4 "ea_pack_blob_calls" gpurun_out/r2y_bench.err | tail -24
python - <<'P'
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print('e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], d['e2e']['h2d_bytes_per_step'])
P
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "calls_only or vcf or mhl_report_golden" > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
