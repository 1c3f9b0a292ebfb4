#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
HB_K1_DENSITY=0 python tools/kernel_times.py C4 2>&1 | tail -2
python tools/kernel_times.py C4 2>&1 | tail -2
HB_FUSED=0 python tools/c2_probe.py 2>&1 | head -2
