#!/bin/bash
EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v6.so python tools/bench_configs.py cfg4_long_nothr 2>&1 | grep config > gpurun_out/r2n.txt
for S in 2048 3072; do EA_PL_S=$S EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v6.so python tools/bench_configs.py cfg4_small 2>&1 | grep config >> gpurun_out/r2n.txt; done
cut -c1-420 gpurun_out/r2n.txt
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
