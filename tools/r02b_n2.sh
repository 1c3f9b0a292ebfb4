#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py > $O/r02b_n2check.log 2>&1; echo "rc=$?" >> $O/r02b_n2check.log
grep -E "MULTI_GPU_CHECK|rc=|: False" $O/r02b_n2check.log | head -20
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2
