#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for fm in 303104 65536; do for sm in 8 12 16; do
echo "--- flat_min_entries $fm smin $sm"
CAPGPU_FLAT_MIN_ENTRIES=$fm CAPGPU_FLAT_SMIN=$sm python tests/gpu_scripts/r2b_msm.py 13:1 14:1 15:1 16:1 2>&1 | grep -v "^{"
[ $fm = 303104 ] && break
done; done
for sm in 8 16; do
echo "--- split 2 GPUs flat_min 65536 smin $sm"
CAPGPU_FLAT_MIN_ENTRIES=65536 CAPGPU_FLAT_SMIN=$sm python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_scripts/split_msm.py 2>&1 | grep split_msm_ms
done
