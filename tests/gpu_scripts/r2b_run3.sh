#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
./tests/gpu_scripts/mb/mb_latency 2>&1 | grep -E "fp_inv|quad" 
python tests/gpu_scripts/r2b_msm.py 2>&1 | tee gpurun_out/r2b_msm_tree2.txt | grep -v "^{" 
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_msm17.csv python tests/gpu_scripts/r2b_msm.py 17:1 12:1 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2b_launches_msm17.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; start = i; break
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
rr = rows[start + 1:]
names = [r[ki] for r in rr]
# last launch group of each size: find last two msm_recode
idx = [i for i, nm in enumerate(names) if "msm_recode" in nm]
for st in (idx[len(idx)//2 - 1], idx[-1]):
    for r in rr[st:st + 8]: print(r[ki][:40], r[gi], r[bi], r[vi])
    print()
PY
