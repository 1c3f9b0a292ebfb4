#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=60
N=${1:-8}
HB_TIMING=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final_n${N}.json 2> gpurun_out/r02_final_n${N}.err
echo "bench rc=$?" >> gpurun_out/r02_final_n${N}.err
