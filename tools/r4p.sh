#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py tests/test_gpu_parity.py -x -q -k "widen or pack or string or templid or calls" > gpurun_out/r4p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4p_pytest.log
tail -4 gpurun_out/r4p_pytest.log
python tools/pack_strings_run.py > gpurun_out/r4p_pack.txt 2>&1; cat gpurun_out/r4p_pack.txt
