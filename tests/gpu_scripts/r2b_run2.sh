#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/r2b_msm.py 2>&1 | tee gpurun_out/r2b_msm_tree.txt | grep -v "^{" 
echo "--- old reduce"
CAPGPU_RED_TREE=0 python tests/gpu_scripts/r2b_msm.py 12:1 15:1 17:1 15:5 2>&1 | tee gpurun_out/r2b_msm_old.txt | grep -v "^{"
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_msm17.csv python tests/gpu_scripts/r2b_msm.py 17:1 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2b_launches_msm17.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; start = i; break
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
last = rows[start + 1:]
last = last[-12:]
for r in last: print(r[ki][:40], r[gi], r[bi], r[vi])
PY
