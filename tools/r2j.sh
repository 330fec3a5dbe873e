#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "auto or byte" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -3 gpurun_out/r2j_pytest.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('kernel_ms', round(d['roofline']['kernel_ms'],4), 'step_ms', round(d['ms_per_step'],4), 'value', round(d['value']/1e9,3), 'frac', round(d['roofline']['frac'],4), 'rows', d['config']['report_rows'], 'pass', d['config']['reads_passing'])" > gpurun_out/r2j.txt 2>&1
cat gpurun_out/r2j.txt
bash tools/ncu_cfg.sh cfg2b_pe_merged k_window_pl r2j_cfg2b
