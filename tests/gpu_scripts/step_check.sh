#!/bin/bash
# One change, one call: prover / primitive parity tests, a short bench, and the launch list of a lockstep group of 8.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TAG=${1:-step}
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -5
python bench.py --no-configs --cpu-sample 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("${TAG}: value %.1f e2e %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["kernel_times_ms_per_proof"], d["single_proof_latency_ms"], d["msm_2p17"])
PY
python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/${TAG}_prof_group_plain.log 2>&1
SKIP=$(grep LAUNCHES_BEFORE gpurun_out/${TAG}_prof_group_plain.log | awk '{print $2}')
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 200 --csv --log-file gpurun_out/${TAG}_launches_group8.csv python tests/gpu_scripts/prof_group.py 8 1 > /dev/null 2>&1
ls -la gpurun_out | tail -4
