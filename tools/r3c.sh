#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto" > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
python tools/bench_configs.py cfg4_long_nothr cfg5_shard 2>&1 | grep config | cut -c1-30,150-460
python tools/e2e_breakdown.py 2>&1 | tail -2
