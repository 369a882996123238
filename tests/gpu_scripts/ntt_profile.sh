#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ntt_reg_kernel" --launch-skip 40 -c 24 -o gpurun_out/r2_ncu_ntt python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_ncu_ntt.log 2>&1
ncu -i gpurun_out/r2_ncu_ntt.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ntt_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_ncu_ntt_raw.csv")))
hdr=rows[0]
gi=hdr.index("Grid Size"); ti=hdr.index("gpu__time_duration.sum")
big=[i for i,r in enumerate(rows[2:]) if "(256, 56" in r[gi] or "(256, 8" in r[gi]]
print("big launches:", big[:6])
PY
# keep the source page of the largest launch only
ncu -i gpurun_out/r2_ncu_ntt.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_ntt_source_all.csv 2>/dev/null
head -c 20000000 gpurun_out/r2_ncu_ntt_source_all.csv > gpurun_out/r2_ncu_ntt_source.csv; rm -f gpurun_out/r2_ncu_ntt_source_all.csv gpurun_out/r2_ncu_ntt.ncu-rep
ls -la gpurun_out/r2_ncu_ntt*
