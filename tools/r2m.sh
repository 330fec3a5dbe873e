#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
python tools/bench_configs.py cfg4_long_nothr cfg4_long_thr 2>&1 | grep config > gpurun_out/r2m.txt
cut -c1-420 gpurun_out/r2m.txt
