#!/bin/bash
# round 2, first GPU call: Rscript probe, baseline GPU tests, e2e split, ncu captures of the kernels that had none
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{ echo "nproc: $(nproc)"; which Rscript R 2>&1 || echo "Rscript: absent on GPU box"; Rscript --version 2>&1 | head -2; lscpu | grep -E 'Model name|Socket|Core|Thread|^CPU\(s\)'; free -g | head -2; nvidia-smi -L; } > gpurun_out/r02_probe_env.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest0.log
HB_TIMING=1 python tools/e2e_split.py > gpurun_out/r02_e2e_split.log 2>&1
python tools/kernel_times.py C3 > gpurun_out/r02_c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_hydro -s 40 -c 2 -o gpurun_out/r02_c3_hydro python tools/kernel_times.py C3 > gpurun_out/r02_c3_ncu.log 2>&1
python tools/c5_probe.py > gpurun_out/r02_c5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_hydro -s 200 -c 2 -o gpurun_out/r02_c5_hydro python tools/c5_probe.py > gpurun_out/r02_c5_ncu.log 2>&1
python tools/rstat_probe.py > gpurun_out/r02_rstat_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hb_rstat -s 2 -c 4 -o gpurun_out/r02_rstat python tools/rstat_probe.py > gpurun_out/r02_rstat_ncu.log 2>&1
HB_FUSED=1 python tools/c2_probe.py > gpurun_out/r02_c2_plain.log 2>&1 &&
HB_FUSED=1 ncu --set full --clock-control none --import-source on -k regex:hb_fused_small -s 300 -c 2 -o gpurun_out/r02_c2_fused python tools/c2_probe.py > gpurun_out/r02_c2_ncu.log 2>&1
ls -la gpurun_out | tail -20
