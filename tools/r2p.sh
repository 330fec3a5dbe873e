#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
(python tools/bench_configs.py cfg4_long_nothr cfg4_long_thr 2>&1 | grep config
EA_NO_VSR=1 python tools/bench_configs.py cfg4_long_nothr 2>&1 | grep config ) > gpurun_out/r2p.txt
cut -c1-440 gpurun_out/r2p.txt
