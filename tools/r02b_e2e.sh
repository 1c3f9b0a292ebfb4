#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
HB_TIMING=1 python tools/e2e_api.py 2>&1 | tail -40
