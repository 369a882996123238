#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
./tests/gpu_scripts/mb/mb_latency 2>&1 | grep -E "fp_inv" 
python tests/gpu_scripts/r2b_msm.py 12:1 15:1 17:1 15:5 15:40 2>&1 | tee gpurun_out/r2b_msm_tree3.txt | grep -v "^{" 
timeout 900 python -m pytest tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_msm17.csv python tests/gpu_scripts/r2b_msm.py 17:1 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2b_launches_msm17.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; start = i; break
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
rr = rows[start + 1:]
for r in rr[-8:]: print(r[ki][:40], r[gi], r[bi], r[vi])
PY
ncu --set full --clock-control none --import-source on -k regex:"msm_red|msm_scan|msm_recode" --launch-skip 40 -c 5 -o gpurun_out/r2b_ncu_red python tests/gpu_scripts/r2b_msm.py 17:1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
