#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1
HB_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
echo "bench rc=$?" >> gpurun_out/r02_bench_full.err
