#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for mb in 4 5; do
CAPGPU_ACC_MINB=$mb python bench.py --steps 4 --warmup 3 --no-configs --cpu-sample 0 > gpurun_out/r2b_bench6.json 2> gpurun_out/r2b_bench6.err || tail -3 gpurun_out/r2b_bench6.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench6.json"))
print("acc minb $mb: value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["kernel_times_ms_per_proof"])
PY
done
