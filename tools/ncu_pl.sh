#!/bin/bash
# ncu --set full capture of the dominant window kernel on configs[1] (run after the plain command exited 0)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_window_pl -s 3 -c 1 -o gpurun_out/r2b_window_pl -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
