#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"msm_accumulate<|ntt_reg_kernel<9>|quotient_kernel|msm_red_strips" -c 40 -o gpurun_out/r2_ncu_group8 python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_ncu_group8.log 2>&1
tail -3 gpurun_out/r2_ncu_group8.log
ncu -i gpurun_out/r2_ncu_group8.ncu-rep --page raw --csv > gpurun_out/r2_ncu_group8_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_group8.ncu-rep --page details --csv > gpurun_out/r2_ncu_group8_details.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_group8.ncu-rep   # gpurun copies back at most 64 MiB: keep the CSV exports only
ls -la gpurun_out/r2_ncu_group8*
