#!/bin/bash
uptime
python tools/e2e_breakdown.py 2>&1 | tail -2
EA_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; grep -A2 "ea_pack_blob_calls" gpurun_out/r2z_bench.err | tail -6
python tools/e2e_breakdown.py 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], d['e2e']['h2d_bytes_per_step'])"
uptime
