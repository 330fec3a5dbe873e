#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
(
tools/ab_env.sh "X=1" v2_minb4 v2_minb5 v2_merged v2_oldfin
for S in 2048 2560 3072; do tools/ab_env.sh "EA_PL_S=$S" v2_minb4 v2_minb5; done
) > gpurun_out/r2c_ab.txt 2>&1
cat gpurun_out/r2c_ab.txt
