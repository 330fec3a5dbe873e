#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto and (synthetic or golden or calls_only)" > gpurun_out/r3d_pytest.log 2>&1; tail -3 gpurun_out/r3d_pytest.log
python tools/bench_configs.py cfg4_long_nothr cfg4_long_thr 2>&1 | grep config | cut -c1-30,150-460
