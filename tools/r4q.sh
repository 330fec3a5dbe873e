#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/alloc_stall_probe.py 2>&1 | tail -3
EA_NO_POOL_RESERVE=1 python tools/alloc_stall_probe.py 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-other --no-cpu --no-widened > gpurun_out/r4q_bench.json 2> gpurun_out/r4q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r4q_bench.json')); e=d['e2e']; print(e['seconds_per_step'], e['cold_first_call_seconds'], e['warmup_seconds'], e['step_seconds'], e['from_pinned_blob']['seconds_per_step'])"
