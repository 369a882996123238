#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for mb in 2 3 4; do
CAPGPU_QUOT_MINB=$mb python bench.py --steps 4 --warmup 3 --no-configs --cpu-sample 0 > gpurun_out/r2b_bench5.json 2> gpurun_out/r2b_bench5.err || tail -3 gpurun_out/r2b_bench5.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench5.json"))
print("quot minb $mb: value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["kernel_times_ms_per_proof"], d["ntt"].get("lockstep_group"))
PY
done
timeout 900 python -m pytest tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -2
