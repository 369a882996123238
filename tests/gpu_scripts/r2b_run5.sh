#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/r2b_msm.py 2>&1 | tee gpurun_out/r2b_msm_tree4.txt | grep -v "^{" 
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_msm17.csv python tests/gpu_scripts/r2b_msm.py 17:1 12:1 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2b_launches_msm17.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr = r; start = i; break
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
rr = rows[start + 1:]
names = [r[ki] for r in rr]
idx = [i for i, nm in enumerate(names) if "msm_recode" in nm]
for st in (idx[len(idx)//2 - 1], idx[-1]):
    for r in rr[st:st + 8]: print(r[ki][:40], r[gi], r[bi], r[vi])
    print()
PY
(time python bench.py --steps 5 --warmup 3) > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err
tail -3 gpurun_out/r2b_bench1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench1.json"))
print("value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
print(d["kernel_times_ms_per_proof"], d["msm_2p17"], d["single_proof_latency_ms"])
for k, v in d.get("configs", {}).get("note_shapes", {}).items(): print(k, round(v["proofs_per_s"], 1), round(v["roofline_frac"], 3))
for r in d.get("configs", {}).get("msm_sweep", []): print("msm", r["points"], round(r["gpu_ms"], 3), r["bit_exact_vs_cpu"], round(r["frac_of_imad_roofline_survey_formula"], 3))
PY
