#!/bin/bash
# ncu --set full of the accumulate launches of one lockstep group of 8 (the roofline kernel of the bench line); CSV only.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/acc_prof_plain.log 2>&1
SKIP=$(grep LAUNCHES_BEFORE gpurun_out/acc_prof_plain.log | awk '{print $2}')
ncu --set full --clock-control none --import-source on --launch-skip-before-match $SKIP -k regex:"^msm_accumulate$" -c 4 -o gpurun_out/acc python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ncu -i gpurun_out/acc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_accumulate_final_raw.csv 2>/dev/null
ncu -i gpurun_out/acc.ncu-rep --page details --csv > gpurun_out/r2_ncu_accumulate_final_details.csv 2>/dev/null
rm -f gpurun_out/acc.ncu-rep
ls -la gpurun_out/r2_ncu_accumulate_final_*
