#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
