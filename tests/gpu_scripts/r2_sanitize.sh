#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_scripts/sanitize.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tests/gpu_scripts/sanitize.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.log
