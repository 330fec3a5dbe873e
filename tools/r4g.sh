#!/bin/bash
# lMHL on the call planes: parity under both kernels, timing at 50 M reads (planar vs byte rows), ncu summary of the planar kernel
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_parity.py -x -q -k "mhl" > gpurun_out/r4g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4g_pytest.log
tail -4 gpurun_out/r4g_pytest.log
python tools/mhl_run.py 50000000 > gpurun_out/r4g_mhl.txt 2>&1
EA_MHL_BYTES=1 python tools/mhl_run.py 50000000 >> gpurun_out/r4g_mhl.txt 2>&1
cat gpurun_out/r4g_mhl.txt
ncu --set full --clock-control none --import-source on -k regex:k_window_mhl_pl -s 1 -c 1 -o /tmp/r2b_mhl_pl -f python tools/mhl_run.py 50000000 > gpurun_out/r4g_ncu_run.log 2>&1
echo "ncu rc=$?"
python profiles/ncu_summary.py /tmp/r2b_mhl_pl.ncu-rep > gpurun_out/r2b_window_mhl_pl_ncu.txt 2>&1
python profiles/src_lines.py /tmp/r2b_mhl_pl.ncu-rep 60 > gpurun_out/r2b_window_mhl_pl_lines.txt 2>&1
head -50 gpurun_out/r2b_window_mhl_pl_ncu.txt
