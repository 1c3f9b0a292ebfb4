#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_PEER_TIMEOUT_S=30
N=${1:-8}; O=gpurun_out
for P in 3 4 2; do
HB_PIECES=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$P bench.py --gpus $N --steps 20 --warmup 5 --quick > $O/r02b_pieces${P}_n${N}.json 2> $O/r02b_pieces${P}_n${N}.err
echo "P=$P rc=$?"; python tools/show_bench.py $O/r02b_pieces${P}_n${N}.json | head -3
done
