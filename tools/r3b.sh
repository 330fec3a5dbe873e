#!/bin/bash
tools/ab_env.sh "X=1" v10 v11 v10 v11 > gpurun_out/r3b.txt 2>&1; cat gpurun_out/r3b.txt
EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v11.so bash tools/ncu_cfg.sh cfg4_small k_window_vsr r3b_vsr
