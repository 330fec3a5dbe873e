#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vertical" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
for pol in 6 0; do for c in cfg2_wgbs150 cfg2b_pe_merged cfg3_capture_cx deep150; do EA_POLICY=$pol python tools/bench_configs.py $c 2>&1 | tail -1; done; done > gpurun_out/r2g_cfgs.txt
cat gpurun_out/r2g_cfgs.txt
