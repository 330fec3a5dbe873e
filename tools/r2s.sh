#!/bin/bash
tools/ab_env.sh "X=1" v8 > gpurun_out/r2s.txt 2>&1; cat gpurun_out/r2s.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto or byte" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -3 gpurun_out/r2s_pytest.log
