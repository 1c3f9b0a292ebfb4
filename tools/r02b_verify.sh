#!/bin/bash
# re-entry check of round 2: GPU tests, smoke, the driver's bench command
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02b_smoke.log
python tools/kernel_times.py C4 > $O/r02b_ktimes.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; echo "bench rc=$?" >> $O/r02b_bench_n1.err
tail -3 $O/r02b_pytest.log; tail -2 $O/r02b_smoke.log; cat $O/r02b_ktimes.log
