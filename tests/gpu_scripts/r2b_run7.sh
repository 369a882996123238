#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/r2b_msm.py 12:1 14:1 17:1 15:5 15:40 2>&1 | tee gpurun_out/r2b_msm_tree6.txt | grep -v "^{" 
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "4 8" "6 8" "4 16"; do
set -- $cfg
python bench.py --steps 5 --warmup 3 --no-configs --ctxs $1 --group $2 > gpurun_out/r2b_bench3_c$1_g$2.json 2> gpurun_out/r2b_bench3.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench3_c$1_g$2.json"))
print("ctxs $1 group $2: value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["kernel_times_ms_per_proof"], d["single_proof_latency_ms"])
PY
done
