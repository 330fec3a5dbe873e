#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py tests/test_gpu_parity.py -x -q -k "widen or mhl or pageable or vcf or text" > gpurun_out/r4i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4i_pytest.log
tail -4 gpurun_out/r4i_pytest.log
python tools/mhl_run.py 50000000 > gpurun_out/r4i_mhl.txt 2>&1; cat gpurun_out/r4i_mhl.txt
python tools/shim_breakdown.py > gpurun_out/r4i_shim.txt 2>&1
grep -v '^\[ea\]' gpurun_out/r4i_shim.txt | tail -12
ncu --set full --clock-control none --import-source on -k regex:k_window_mhl_pl -s 1 -c 1 -o /tmp/r2b_mhl_pl -f python tools/mhl_run.py 50000000 > gpurun_out/r4i_ncu_run.log 2>&1
python profiles/ncu_summary.py /tmp/r2b_mhl_pl.ncu-rep > gpurun_out/r2b_window_mhl_pl_ncu.txt 2>&1
python profiles/src_lines.py /tmp/r2b_mhl_pl.ncu-rep 60 > gpurun_out/r2b_window_mhl_pl_lines.txt 2>&1
grep -E "time_duration|thread_inst_executed_per|inst_executed.sum|warps_active" gpurun_out/r2b_window_mhl_pl_ncu.txt
