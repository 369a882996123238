#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_scripts/sanitize.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tests/gpu_scripts/sanitize.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_racecheck.log
python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_prof_group_plain.log 2>&1
SKIP=$(grep LAUNCHES_BEFORE gpurun_out/r2_prof_group_plain.log | awk '{print $2}')
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 200 --csv --log-file gpurun_out/r2_launches_group8_final.csv python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"quotient_kernel" -c 2 -o gpurun_out/r2_ncu_quot python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ncu -i gpurun_out/r2_ncu_quot.ncu-rep --page raw --csv > gpurun_out/r2_ncu_quotient_group8_raw.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_quot.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_msm_lone.csv python tests/gpu_scripts/msm_latency.py 17:1 12:1 > /dev/null 2>&1
ls -la gpurun_out | tail -6
