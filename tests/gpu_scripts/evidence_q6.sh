#!/bin/bash
# Evidence for the 6n quotient domain build: parity tests, sanitizer, launch list of a lockstep group of 8, ncu --set full of the
# quotient-domain transforms and the quotient kernel.  CSV only (the .ncu-rep files are deleted: gpurun_out/ is capped at 64 MiB).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_scripts/sanitize.py > gpurun_out/r2_sanitizer_memcheck_q6.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck_q6.log
python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_prof_group_plain.log 2>&1
SKIP=$(grep LAUNCHES_BEFORE gpurun_out/r2_prof_group_plain.log | awk '{print $2}')
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 200 --csv --log-file gpurun_out/r2_launches_group8_q6.csv python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ncu --set full --clock-control none --launch-skip-before-match $SKIP -k regex:"ntt_reg_kernel|quotient_kernel|ntt3_recombine" -c 10 -o gpurun_out/q6 python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ncu -i gpurun_out/q6.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ntt3_quotient_group8_raw.csv 2>/dev/null
rm -f gpurun_out/q6.ncu-rep
ls -la gpurun_out | tail -5
