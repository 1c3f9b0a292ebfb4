#!/bin/bash
# usage: tools/r02b_multi.sh N [check]
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_PEER_TIMEOUT_S=30
N=${1:-2}; O=gpurun_out; mkdir -p $O
if [ "$2" = "check" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > $O/r02b_n${N}_check.log 2>&1
echo "check rc=$?" >> $O/r02b_n${N}_check.log; grep -E "MULTI_GPU_CHECK|rc=" $O/r02b_n${N}_check.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02b_bench_n${N}.json 2> $O/r02b_bench_n${N}.err
echo "bench rc=$?" >> $O/r02b_bench_n${N}.err; tail -3 $O/r02b_bench_n${N}.err
python tools/show_bench.py $O/r02b_bench_n${N}.json
