#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest3.log
timeout 300 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
echo "bench rc=$?" >> gpurun_out/r02_bench2.err
