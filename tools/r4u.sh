#!/bin/bash
cd "$GRAFT_REPO_ROOT"
bash tools/ncu_cfg.sh cfg4_long_nothr k_window_vsr r2b_cfg4_vsr
cat gpurun_out/r2b_cfg4_vsr_plain.json | cut -c1-420
grep -E "time_duration|thread_inst_executed_per|inst_executed.sum|warps_active|issue_active|dram__bytes|registers_per" gpurun_out/r2b_cfg4_vsr_ncu.txt
head -12 gpurun_out/r2b_cfg4_vsr_lines.txt | cut -c1-150
