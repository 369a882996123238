#!/bin/bash
# A/B of the quotient domain: 6n points as three 2n-point cosets (default) against the 8n-point coset (CAPGPU_QDOMAIN=8).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -5
for q in 6 8; do
  CAPGPU_QDOMAIN=$q python bench.py --no-configs --cpu-sample 0 > gpurun_out/qdomain_$q.json 2> gpurun_out/qdomain_$q.err
  python - <<PY
import json
d = json.load(open("gpurun_out/qdomain_$q.json"))
print("qdomain $q: value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["kernel_times_ms_per_proof"], d["single_proof_latency_ms"])
PY
done
