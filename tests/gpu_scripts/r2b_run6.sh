#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/r2b_msm.py 12:1 13:1 14:1 15:1 17:1 15:5 15:40 2>&1 | tee gpurun_out/r2b_msm_tree5.txt | grep -v "^{" 
echo "--- flat smin 16"
CAPGPU_FLAT_SMIN=16 python tests/gpu_scripts/r2b_msm.py 12:1 13:1 14:1 2>&1 | grep -v "^{"
echo "--- no flat, wmin 13"
CAPGPU_ACC_FLAT=0 python tests/gpu_scripts/r2b_msm.py 12:1 13:1 14:1 17:1 2>&1 | grep -v "^{"
echo "--- no flat, wmin 15"
CAPGPU_ACC_FLAT=0 CAPGPU_WINDOW_MIN=15 python tests/gpu_scripts/r2b_msm.py 12:1 13:1 14:1 2>&1 | grep -v "^{"
echo "--- flat, wmin 12"
CAPGPU_WINDOW_MIN=12 python tests/gpu_scripts/r2b_msm.py 12:1 13:1 2>&1 | grep -v "^{"
echo "--- tiles instead of strips"
CAPGPU_RED_STRIP_MIN=100000000 python tests/gpu_scripts/r2b_msm.py 15:40 2>&1 | grep -v "^{"
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -3
(time python bench.py --steps 5 --warmup 3 --no-configs) > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err || (time python bench.py --steps 5 --warmup 3) > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err
tail -3 gpurun_out/r2b_bench2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench2.json"))
print("value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
print(d["kernel_times_ms_per_proof"], d["msm_2p17"], d["single_proof_latency_ms"])
for k, v in d.get("configs", {}).get("note_shapes", {}).items(): print(k, round(v["proofs_per_s"], 1), round(v["roofline_frac"], 3))
PY
