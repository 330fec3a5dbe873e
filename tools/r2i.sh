#!/bin/bash
for e in "X=1" "EA_PL_UNORDERED=1"; do tools/ab_env.sh "$e" v4; done > gpurun_out/r2i.txt 2>&1
cat gpurun_out/r2i.txt
