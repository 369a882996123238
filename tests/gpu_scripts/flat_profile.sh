#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"msm_accumulate_flat" --launch-skip 20 -c 1 -o gpurun_out/r2_ncu_flat python tests/gpu_scripts/msm_latency.py 17:1 > /dev/null 2>&1
ncu -i gpurun_out/r2_ncu_flat.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_flat_source.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_flat.ncu-rep --page raw --csv > gpurun_out/r2_ncu_flat_raw.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_flat.ncu-rep
ls -la gpurun_out/r2_ncu_flat*
