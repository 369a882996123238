#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -3
python tests/gpu_scripts/r2b_msm.py 12:1 13:1 14:1 17:1 2>&1 | grep -v "^{"
N=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/gpu_scripts/split_msm.py 2>&1 | grep split_msm_ms
