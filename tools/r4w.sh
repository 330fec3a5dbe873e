#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/multi_device_run.py > gpurun_out/r4w_multi_device.json 2> gpurun_out/r4w_multi_device.err; echo "rc=$?"; tail -2 gpurun_out/r4w_multi_device.err
python -c "
import json; d=json.load(open('gpurun_out/r4w_multi_device.json'))
print([(r['devices'], round(r['total_s'],3)) for r in d['runs']], [(r['devices'], round(r['pack_s'],3), round(r['total_s'],3)) for r in d['from_reader_strings']], d['identical_across_device_counts'])"
