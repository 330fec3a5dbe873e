#!/bin/bash
# round-2 session-2 check: new boundary tests, then the default bench line (all configs)
cd "$GRAFT_REPO_ROOT"
python -m pytest tests/test_gpu_widen.py -x -q > gpurun_out/r4a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4a_pytest.log
tail -15 gpurun_out/r4a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r4a_bench.err
cut -c1-3000 gpurun_out/r4a_bench.json
free -g | head -2; nproc
