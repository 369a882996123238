#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prover.py -m gpu -x -q) > gpurun_out/r2_tests4.log 2>&1
tail -3 gpurun_out/r2_tests4.log
run() { # name, env..., then bench args
  name=$1; shift
  env "$@" python bench.py --steps 5 --warmup 2 --no-extras > gpurun_out/r2_b4_$name.json 2> gpurun_out/r2_b4_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_b4_$name.json"))
    print("$name: value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("$name failed", e)
PY
}
run c15 CAPGPU_X=0
run c16 CAPGPU_WINDOW_BITS=16
run c14 CAPGPU_WINDOW_BITS=14
run c15_seg64 CAPGPU_RED_SEG=64
run c15_seg16 CAPGPU_RED_SEG=16
run c15_6ctx CAPGPU_X=0 BENCH_CTXS=6
