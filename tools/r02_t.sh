#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest4.log
python tools/delta_probe.py > gpurun_out/r02_delta_probe2.log 2>&1
