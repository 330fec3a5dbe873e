#!/bin/bash
bash tools/ncu_kernel.sh 0 k_window_pl r2_window_pl
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/r2_plain.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > /dev/null 2>&1
echo "launch list rc=$?"
