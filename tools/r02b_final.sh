#!/bin/bash
# round 2 (second session): GPU tests, smoke, the driver's bench command, launch list + full ncu captures of the kernels that changed
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O $O/prof
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log; tail -4 $O/r02b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02b_smoke.log; tail -2 $O/r02b_smoke.log
python tools/kernel_times.py C4 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; echo "bench rc=$?" >> $O/r02b_bench_n1.err; tail -2 $O/r02b_bench_n1.err
if [ "$1" = "ncu" ]; then
B="python bench.py --steps 20 --warmup 1 --no-extras"
$B > $O/r02b_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file $O/r02b_launches.csv $B > $O/r02b_ncu_l.log 2>&1
$B > $O/r02b_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"hb_k1_pair|hb_k3_wide" -s 240 -c 2 -o $O/r02b_k1k3 $B > $O/r02b_ncu_f.log 2>&1
python tools/c2_persist_probe.py > $O/r02b_c2p_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_persist_small -s 12 -c 1 -o $O/r02b_c2_persist python tools/c2_persist_probe.py > $O/r02b_c2p_ncu.log 2>&1
python tools/ncu_summary.py launches $O/r02b_launches.csv $O/prof/r02b_final_launches.txt > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02b_k1k3.ncu-rep $O/prof/r02b_final_k1k3_full.txt > /dev/null 2>&1
python tools/ncu_summary.py traffic $O/r02b_k1k3.ncu-rep $O/prof/ncu_traffic.json > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02b_c2_persist.ncu-rep $O/prof/r02b_c2_persist_full.txt > /dev/null 2>&1
rm -f $O/r02b_k1k3.ncu-rep $O/r02b_c2_persist.ncu-rep
ls -la $O/prof
fi
