#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_scripts/split_msm.py > gpurun_out/r2_split_${N}gpu.log 2>&1
grep split_msm_ms gpurun_out/r2_split_${N}gpu.log || tail -30 gpurun_out/r2_split_${N}gpu.log
