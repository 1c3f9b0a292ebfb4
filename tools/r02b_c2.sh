#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_c2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_c2_pytest.log
tail -12 $O/r02b_c2_pytest.log
timeout 300 python tools/kernel_times.py C4 2>&1 | tail -3
HB_K1_DENSITY=0 timeout 300 python tools/kernel_times.py C4 2>&1 | tail -2
HB_K1_DENSITY=0 HB_TDIST_DENSE=1 timeout 300 python tools/kernel_times.py C4 2>&1 | tail -2
timeout 300 python tools/c2_probe.py 2>&1 | tail -6
timeout 300 python tools/c2_adapt_probe.py 2>&1 | head -2
