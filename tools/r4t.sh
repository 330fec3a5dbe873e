#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py -x -q -k "multi or shards or plan" > gpurun_out/r4t_pytest.log 2>&1; tail -3 gpurun_out/r4t_pytest.log
