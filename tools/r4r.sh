#!/bin/bash
cd "$GRAFT_REPO_ROOT"

python tools/shim_breakdown.py 2>&1 | grep -E "call|fresh" | head -14
