#!/bin/bash
# bench sweep over (contexts, lockstep group); args: list of "ctxs group" pairs
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  python bench.py --ctxs $1 --group $2 --steps 5 --warmup 2 --no-extras > gpurun_out/r2_bench_c$1_g$2.json 2> gpurun_out/r2_bench_c$1_g$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c$1_g$2.json"))
    print("ctxs $1 group $2: value %.1f e2e %.1f copy_ms/note %.3f avg_group %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["host_copy_ms_per_note"], d["e2e"]["avg_lockstep_group"]))
except Exception as e:
    print("ctxs $1 group $2 failed", e)
PY
done
