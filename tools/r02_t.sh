#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest7.log
python tools/kernel_times.py C3 > gpurun_out/r02_c3_plain2.log 2>&1
python tools/c5_probe.py > gpurun_out/r02_c5_plain2.log 2>&1
