#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py -x -q > gpurun_out/r4b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4b_pytest.log
tail -5 gpurun_out/r4b_pytest.log
EA_TRACE=1 python tools/shim_breakdown.py > gpurun_out/r4b_shim.txt 2>&1
tail -80 gpurun_out/r4b_shim.txt
