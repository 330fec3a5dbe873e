#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py tests/test_gpu_parity.py -x -q -k "pack or calls or string or widen" > gpurun_out/r4v_pytest.log 2>&1; tail -2 gpurun_out/r4v_pytest.log
EA_TRACE=1 python tools/shim_breakdown.py 2>&1 | grep -E "transposition|call [0-9]|fresh" | head -24
