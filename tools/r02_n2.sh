#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=30
N=${1:-2}
if [ "$2" = "tests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_n${N}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r02_n${N}_pytest.log
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r02_n${N}_check.log 2>&1
echo "check rc=$?" >> gpurun_out/r02_n${N}_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --quick > gpurun_out/r02_n${N}_bench.json 2> gpurun_out/r02_n${N}_bench.err
echo "bench rc=$?" >> gpurun_out/r02_n${N}_bench.err
