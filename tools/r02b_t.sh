#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/kernel_times.py C4 2>&1 | tail -3
