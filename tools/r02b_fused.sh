#!/bin/bash
# fused proposal + log-density kernel: parity tests, then per-kernel times on C4 with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_proposal or wide_kernels or group" > $O/r02b_fused_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_fused_pytest.log
tail -15 $O/r02b_fused_pytest.log
HB_K1K2=1 timeout 300 python tools/kernel_times.py C4 > $O/r02b_fused_ktimes.log 2>&1; cat $O/r02b_fused_ktimes.log
HB_K1K2=0 timeout 300 python tools/kernel_times.py C4 > $O/r02b_plain_ktimes.log 2>&1; cat $O/r02b_plain_ktimes.log
for r in 9 10 11; do echo "ring $r"; HB_K1F_RING=$r HB_K1K2=1 timeout 300 python tools/kernel_times.py C4 2>&1 | tail -2; done
