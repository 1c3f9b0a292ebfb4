#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
HB_PERSIST_TRACE=1 timeout 300 python tools/c2_persist_probe.py > gpurun_out/r02_persist_trace.log 2>&1
