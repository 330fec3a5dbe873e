#!/bin/bash
# 2-GPU: the bench line on config 5 under torchrun, the one-process multi-device entry, the sharded GPU tests
cd "$GRAFT_REPO_ROOT"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r4c_bench_2gpu.json 2> gpurun_out/r4c_bench_2gpu.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r4c_bench_2gpu.json; tail -3 gpurun_out/r4c_bench_2gpu.err
python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r4c_ref_2gpu.json 2>/dev/null; cut -c1-300 gpurun_out/r4c_ref_2gpu.json
python tools/multi_device_run.py > gpurun_out/r4c_multi_device.json 2> gpurun_out/r4c_multi_device.err; echo "multi rc=$?"; cat gpurun_out/r4c_multi_device.json | head -60; tail -3 gpurun_out/r4c_multi_device.err
python -m pytest tests/test_gpu_widen.py -x -q -k "multi or shards" > gpurun_out/r4c_pytest.log 2>&1; tail -3 gpurun_out/r4c_pytest.log
free -g | head -2; nproc
