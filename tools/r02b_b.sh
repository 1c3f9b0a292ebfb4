#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; echo "bench rc=$?"; tail -2 $O/r02b_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02b_reference_n1.json 2> $O/r02b_reference_n1.err; echo "reference rc=$?"; cat $O/r02b_reference_n1.json | head -c 600
