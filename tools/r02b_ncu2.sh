#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O $O/prof
B="python bench.py --steps 4 --warmup 1 --no-extras"
$B > $O/r02b_ncu2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"hb_k3_wide" -s 25 -c 1 -o $O/r02b_k3adapt $B > $O/r02b_ncu2_a.log 2>&1
HB_K1_DENSITY=0 $B > $O/r02b_ncu2_plain2.log 2>&1 &&
HB_K1_DENSITY=0 ncu --set full --clock-control none --import-source on -k regex:"hb_tdist_struct" -s 25 -c 1 -o $O/r02b_struct $B > $O/r02b_ncu2_b.log 2>&1
python tools/ncu_summary.py full $O/r02b_k3adapt.ncu-rep $O/prof/r02b_k3_adaptation_full.txt > /dev/null 2>&1
python tools/ncu_summary.py full $O/r02b_struct.ncu-rep $O/prof/r02b_tdist_struct_full.txt > /dev/null 2>&1
ncu -i $O/r02b_k3adapt.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > $O/prof/r02b_k3_adaptation_stalls.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
h, u, r = rows[0], rows[1], rows[2]
for i, name in enumerate(h):
    if 'warp_issue_stalled' in name and 'pct' not in name and '_not_issued' not in name:
        try:
            v = float(r[i].replace(',', ''))
        except ValueError:
            continue
        if v > 0.2: print(f"{name:90s} {v:10.2f} {u[i]}")
for name in ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum'):
    if name in h: print(f"{name:90s} {r[h.index(name)]:>14s} {u[h.index(name)]}")
PY
rm -f $O/r02b_k3adapt.ncu-rep $O/r02b_struct.ncu-rep
cat $O/prof/r02b_k3_adaptation_full.txt $O/prof/r02b_k3_adaptation_stalls.txt $O/prof/r02b_tdist_struct_full.txt
