#!/bin/bash
# what the driver runs at round end: all GPU tests, smoke(), the reference arm, the default bench line
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/ -x -q -m gpu > gpurun_out/r4j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4j_pytest.log
tail -6 gpurun_out/r4j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
python bench.py --steps 10 --warmup 3 > gpurun_out/r4j_bench.json 2> gpurun_out/r4j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r4j_bench.err
python tools/multi_device_run.py > gpurun_out/r4j_multi_device_1gpu.json 2>/dev/null
