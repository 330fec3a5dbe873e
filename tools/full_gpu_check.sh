#!/bin/bash
python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_full_pytest.log
tail -6 gpurun_out/r2_full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
