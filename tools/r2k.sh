#!/bin/bash
export EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v5.so
(
python tools/bench_configs.py cfg2_wgbs150 cfg2b_pe_merged deep150 2>&1 | grep config
EA_PL_DEEP=0 python tools/bench_configs.py cfg3_capture_cx 2>&1 | grep config
EA_PL_DEEP=1 python tools/bench_configs.py cfg3_capture_cx 2>&1 | grep config
) > gpurun_out/r2k.txt
cut -c1-400 gpurun_out/r2k.txt
