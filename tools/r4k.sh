#!/bin/bash
cd "$GRAFT_REPO_ROOT"
EA_TRACE=1 python bench.py --steps 3 --warmup 3 --no-other --no-cpu --no-widened > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err
grep -B1 -A14 'ea_pack_calls' gpurun_out/r4k_bench.err | head -120
python -c "
import json; d=json.load(open('gpurun_out/r4k_bench.json')); e=d['e2e']; print(e['seconds_per_step'], e['cold_first_call_seconds'])"
