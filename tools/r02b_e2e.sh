#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests -m gpu -x -q -k "python_api or sliced or fetch or golden or chain" 2>&1 | tail -3
HB_TIMING=1 python tools/e2e_api.py 2>&1 | grep -E "fetch_chain  |rep "
