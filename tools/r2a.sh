#!/bin/bash
# round-2 first GPU call: parity of the planar path, then A/B timings
set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_bench_pl.json 2> gpurun_out/r2a_bench_pl.err; tail -c 1500 gpurun_out/r2a_bench_pl.json
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --policy 5 > gpurun_out/r2a_bench_byte.json 2> gpurun_out/r2a_bench_byte.err
(
tools/ab_env.sh "X=1" pl_minb4 pl_minb5 pl_minb6
for S in 1024 1536 2048 2560 3072; do tools/ab_env.sh "EA_PL_S=$S" pl_minb4 pl_minb5; done
) > gpurun_out/r2a_ab.txt 2>&1
cat gpurun_out/r2a_ab.txt
