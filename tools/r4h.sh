#!/bin/bash
# round-2 final profile of the dominant kernel: ncu --set full summary + per-line profile + SASS, launch list of the step
cd "$GRAFT_REPO_ROOT"
bash tools/ncu_kernel.sh 0 k_window_pl r2b_window_pl_final
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > gpurun_out/r2b_plain.json 2>/dev/null &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-other > /dev/null 2>&1
echo "launch list rc=$?"
grep -c k_window_pl gpurun_out/r2b_launches.csv
head -30 gpurun_out/r2b_window_pl_final_ncu.txt
