#!/bin/bash
# usage: ncu_cfg.sh <config name> <kernel regex> <tag> [policy]
CFG=$1; KR=$2; TAG=$3; export EA_POLICY=${4:-0}
python tools/bench_configs.py $CFG > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:$KR -s 2 -c 1 -o /tmp/$TAG -f python tools/bench_configs.py $CFG > gpurun_out/${TAG}_run.log 2>&1
echo "ncu rc=$?"
python profiles/ncu_summary.py /tmp/$TAG.ncu-rep > gpurun_out/${TAG}_ncu.txt 2>&1
python profiles/src_lines.py /tmp/$TAG.ncu-rep 400 > gpurun_out/${TAG}_lines.txt 2>&1
