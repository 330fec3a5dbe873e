#!/bin/bash
# N-GPU bench line on BASELINE configs[4] under torchrun (N = number of visible GPUs)
cd "$GRAFT_REPO_ROOT"
N=$(nvidia-smi -L | wc -l)
free -g | head -2; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r4n_bench_${N}gpu.json 2> gpurun_out/r4n_bench_${N}gpu.err; echo "bench rc=$?"
python -c "
import json,sys; d=json.load(open('gpurun_out/r4n_bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['halo'], d['e2e'].get('host_threads'))"
tail -3 gpurun_out/r4n_bench_${N}gpu.err
