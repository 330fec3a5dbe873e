#!/bin/bash
tools/ab_env.sh "X=1" v3_s0_nopf v3_s0 v3_s1 v3_s2 v3_s3 v3_s2_nopf > gpurun_out/r2d_ab.txt 2>&1
cat gpurun_out/r2d_ab.txt
for c in cfg2b_pe_merged cfg3_capture_cx deep150 cfg4_long_nothr; do EA_LIB_PATH=$PWD/epialleler_b200/lib/libepialleler_b200.so python tools/bench_configs.py $c 2>&1 | tail -1; done > gpurun_out/r2d_cfgs.txt
cat gpurun_out/r2d_cfgs.txt
