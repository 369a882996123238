#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"msm_accumulate$|ntt_reg_kernel" --launch-skip 6 -c 10 -o gpurun_out/r2_ncu_acc python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_ncu_acc.log 2>&1
tail -3 gpurun_out/r2_ncu_acc.log
ncu -i gpurun_out/r2_ncu_acc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_acc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_acc.ncu-rep --page details --csv > gpurun_out/r2_ncu_acc_details.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_acc.ncu-rep
ncu --set full --clock-control none -k regex:"msm_accumulate_flat|msm_red_tiles|msm_red_planes" --launch-skip 60 -c 3 -o gpurun_out/r2_ncu_lone python tests/gpu_scripts/r2b_msm.py 17:1 > gpurun_out/r2_ncu_lone.log 2>&1
ncu -i gpurun_out/r2_ncu_lone.ncu-rep --page raw --csv > gpurun_out/r2_ncu_lone_raw.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_lone.ncu-rep
ls -la gpurun_out/r2_ncu_*
