#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_c2_pytest.log
tail -5 $O/r02b_c2_pytest.log
timeout 300 python tools/c2_probe.py 2>&1 | tail -6
timeout 300 python tools/c2_adapt_probe.py 2>&1 | head -2
HB_PERSIST_TRACE=1 timeout 100 python - <<'PY' 2>&1 | grep "persist trace" | sed -n 9,14p
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
from hydrobayes_b200 import Sampler, _abi as A
cfg = A.default_config(nc=1024, t=2400, d=100, lik=2, seed=1, thin=10)
with Sampler(cfg, 0) as s:
    s.set_target("t_distribution"); s.set_initial(H.x0_tdist(1024, 100)); s.run(260)
PY
