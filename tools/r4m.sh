#!/bin/bash
cd "$GRAFT_REPO_ROOT"
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Flags" | cut -c1-400 | sed 's/Flags.*avx512/Flags: ... avx512 present .../' | head -6
python -m pytest tests/test_gpu_widen.py -x -q -k "pack_calls or pageable" > gpurun_out/r4m_pytest.log 2>&1; tail -2 gpurun_out/r4m_pytest.log
EA_TRACE=1 python tools/shim_breakdown.py 2>&1 | grep -E "transposition|^call|fresh|pre-touched" | head -30
