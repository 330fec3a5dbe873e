#!/bin/bash
# usage: ncu_vs.sh <policy> <kernel regex> <tag>
POL=${1:-6}; KR=${2:-k_window_vs}; TAG=${3:-r2f_window_vs}
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-other --policy $POL > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:$KR -s 3 -c 1 -o /tmp/$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-other --policy $POL > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"
ls -la /tmp/$TAG.ncu-rep
python profiles/ncu_summary.py /tmp/$TAG.ncu-rep > gpurun_out/${TAG}_ncu.txt 2>&1
python profiles/src_lines.py /tmp/$TAG.ncu-rep 400 > gpurun_out/${TAG}_lines.txt 2>&1
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null; gzip -f gpurun_out/${TAG}_sass.csv
