#!/bin/bash
tools/ab_env.sh "X=1" v10 v10_noallin v10 v10_noallin > gpurun_out/r3a.txt 2>&1; cat gpurun_out/r3a.txt
for v in v10; do EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/$v.so python tools/bench_configs.py cfg2b_pe_merged cfg3_capture_cx deep150 2>&1 | grep config | cut -c1-30,150-420; done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto" > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
