#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1
timeout 900 python -m pytest tests -m gpu -x -q -k "fused or c2_size or wide_kernels_native or random_config or replay" > gpurun_out/r02_pytest6.log 2>&1
echo "rc=$?" >> gpurun_out/r02_pytest6.log
HB_PERSIST_TRACE=1 timeout 300 python -c "
import sys; sys.path.insert(0,'tests')
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A
cfg = A.default_config(nc=1024, t=2400, d=100, lik=2, seed=1, thin=10)
with Sampler(cfg, 0) as s:
    s.set_target('t_distribution'); s.set_initial(H.x0_tdist(1024, 100))
    s.run(400); s.run(50)
" > gpurun_out/r02_persist_trace.log 2>&1
timeout 300 python tools/c2_probe.py > gpurun_out/r02_c2_probe.log 2>&1
