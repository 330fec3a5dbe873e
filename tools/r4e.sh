#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/shim_breakdown.py > gpurun_out/r4e_shim.txt 2>&1
grep -v '^\[ea\]' gpurun_out/r4e_shim.txt | tail -20
