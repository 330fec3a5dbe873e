#!/bin/bash
cd "$GRAFT_REPO_ROOT"
bash tools/ncu_cfg.sh cfg3_capture_cx k_window_pl r2b_cfg3
bash tools/ncu_cfg.sh cfg2b_pe_merged k_window_pl r2b_cfg2b
cat gpurun_out/r2b_cfg3_plain.json gpurun_out/r2b_cfg2b_plain.json | cut -c1-400
grep -E "time_duration|thread_inst_executed_per|inst_executed.sum|warps_active|issue_active|dram__bytes_read|shared_atom|bank_conflicts|stalled_(barrier|long|short|wait|not_sel|math)" gpurun_out/r2b_cfg3_ncu.txt
