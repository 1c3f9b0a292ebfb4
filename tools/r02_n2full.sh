#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=60
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_n${N}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02_n${N}_pytest.log
HB_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_n1_full.json 2> gpurun_out/r02_n1_full.err
echo "bench rc=$?" >> gpurun_out/r02_n1_full.err
HB_TIMING=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_n${N}_full.json 2> gpurun_out/r02_n${N}_full.err
echo "bench rc=$?" >> gpurun_out/r02_n${N}_full.err
