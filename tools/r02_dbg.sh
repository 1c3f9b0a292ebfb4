#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export HB_BACKTRACE=1 HB_PEER_TIMEOUT_S=20
for t in 60 300 1500; do
timeout 120 python -c "
import sys; sys.path.insert(0,'tests')
import hydrobayes_b200 as hb
print('run t=$t', flush=True)
res = hb.dream_parallel('mixture', lik=1, par_info=dict(initial='uniform', min=-20, max=20, prior='flat'), nc=2048, t=$t, d=1, seed=4, verbose=False)
print('ok', res['chain'].shape, flush=True)
" >> gpurun_out/r02_dbg.log 2>&1
echo "rc=$?" >> gpurun_out/r02_dbg.log
done
timeout 200 python bench.py --steps 5 --warmup 3 --quick >> gpurun_out/r02_dbg.log 2>&1
echo "bench rc=$?" >> gpurun_out/r02_dbg.log
