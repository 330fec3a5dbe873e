#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for P in 0 6; do EA_POLICY=$P python tools/bench_configs.py cfg3_capture_cx cfg3_small 2>&1 | cut -c1-330; done
