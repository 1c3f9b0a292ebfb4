#!/bin/bash
# last GPU call of round 2: the GPU tests and smoke() at HEAD (what the driver re-runs at round end)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02c_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02c_smoke.log
tail -4 $O/r02c_pytest.log; tail -2 $O/r02c_smoke.log
