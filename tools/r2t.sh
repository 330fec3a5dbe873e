#!/bin/bash
python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2t_bench.json'))
print('value', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('gpu_matches_bit_exact'))
print(d['two_pass_variant']); print(d.get('widened_paths'))
P
EA_TRACE=1 python tools/e2e_breakdown.py 2>&1 | tail -40 > gpurun_out/r2t_trace.txt; tail -30 gpurun_out/r2t_trace.txt
