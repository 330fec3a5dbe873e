#!/bin/bash
export EA_LIB_PATH=$PWD/epialleler_b200/lib/ab/v9.so
for cfg in "EA_HOST_PACK_THREADS=16" "EA_HOST_PACK_THREADS=14" "EA_HOST_PACK_THREADS=16 EA_PACK_CHUNK_MB=96" "EA_HOST_PACK_THREADS=8" "EA_HOST_PACK_THREADS=2"; do
echo "== $cfg"; env $cfg EA_TRACE=1 python tools/e2e_breakdown.py 2>&1 | grep -A3 "ea_pack_blob_calls" | tail -4; env $cfg python tools/e2e_breakdown.py 2>&1 | tail -2
done > gpurun_out/r2w.txt 2>&1
cat gpurun_out/r2w.txt
