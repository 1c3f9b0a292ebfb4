#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=30
N=${1:-8}; shift
for P in "$@"; do
HB_PIECES=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --quick > gpurun_out/r02_n${N}_p${P}_bench.json 2> gpurun_out/r02_n${N}_p${P}_bench.err
echo "bench rc=$?" >> gpurun_out/r02_n${N}_p${P}_bench.err
done
