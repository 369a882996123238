#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
./tests/gpu_scripts/mb/mb_latency > gpurun_out/r2b_mb_latency.txt 2>&1
cat gpurun_out/r2b_mb_latency.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
