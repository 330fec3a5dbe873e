#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/exchange_compare.py > gpurun_out/r2_exchange.json 2> gpurun_out/r2_exchange.err; echo "exchange rc=$?"; tail -3 gpurun_out/r2_exchange.err
cat gpurun_out/r2_exchange.json | head -60
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print('N=2 value', d['value']/1e9, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9, d['e2e']['seconds_per_step'], d['e2e']['h2d_bytes_per_step'])
P
