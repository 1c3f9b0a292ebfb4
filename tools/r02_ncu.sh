#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --steps 20 --warmup 1 --no-extras"
$B > $O/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02_launches.csv $B > $O/r02_ncu_l.log 2>&1
$B > $O/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"hb_k1_pair|hb_tdist_dmma|hb_k3_wide" -s 360 -c 3 -o $O/r02_k1k2k3 $B > $O/r02_ncu_f.log 2>&1
python tools/c2_persist_probe.py > $O/r02_c2p_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_persist_small -c 2 -o $O/r02_c2_persist python tools/c2_persist_probe.py > $O/r02_c2p_ncu.log 2>&1
HB_FORCE_DELTA=1 HB_PIECES=2 python tools/delta_probe.py 131072 delta > $O/r02_deltap_plain.log 2>&1 &&
HB_FORCE_DELTA=1 HB_PIECES=2 ncu --set full --clock-control none --import-source on -k regex:"hb_delta_apply|hb_delta_push|hb_delta_send" -s 600 -c 3 -o $O/r02_delta python tools/delta_probe.py 131072 "delta P=2" > $O/r02_deltap_ncu.log 2>&1
python tools/kernel_times.py C3 > $O/r02_c3_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_hydro -s 40 -c 2 -o $O/r02_c3_hydro_v2 python tools/kernel_times.py C3 > $O/r02_c3_ncu2.log 2>&1
# summaries are produced here: the reports themselves exceed what gpurun brings back
mkdir -p $O/prof
python tools/ncu_summary.py launches $O/r02_launches.csv $O/prof/r02_final_launches.txt > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02_k1k2k3.ncu-rep $O/prof/r02_final_k1k2k3_full.txt > /dev/null 2>&1
python tools/ncu_summary.py traffic $O/r02_k1k2k3.ncu-rep $O/prof/ncu_traffic.json > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02_c2_persist.ncu-rep $O/prof/r02_c2_persist_full.txt > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02_delta.ncu-rep $O/prof/r02_delta_kernels_full.txt > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02_c3_hydro_v2.ncu-rep $O/prof/r02_c3_hydro_v2_full.txt > /dev/null 2>&1
ncu -i $O/r02_c2_persist.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/prof/r02_c2_persist_source.csv 2>/dev/null
rm -f $O/r02_k1k2k3.ncu-rep $O/r02_c2_persist.ncu-rep
du -sh $O
