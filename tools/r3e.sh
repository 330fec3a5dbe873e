#!/bin/bash
(for S in 0 2112 2176 2240 2272 2304 2336 2368 2432 2496 2528; do tools/ab_env.sh "EA_PL_S=$S" v12; done) > gpurun_out/r3e.txt 2>&1
cat gpurun_out/r3e.txt
