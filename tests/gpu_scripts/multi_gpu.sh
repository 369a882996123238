#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_primitives.py -m gpu -q -k "split_msm_across" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
    print("N=$N value %.1f e2e %.1f split %s" % (d["value"], d["e2e"]["value"], d.get("msm_2p17_split")))
    print(d["clocks"], d["config"])
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/r2_bench_${N}gpu.err
