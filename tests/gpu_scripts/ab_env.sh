#!/bin/bash
# A/B runs of the short bench under environment / flag variants:  ab_env.sh "NAME=VAL ..." "--flags" ...
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i + 1))
  envs=""; flags=""
  for tok in $spec; do case "$tok" in --*|[0-9]*) flags="$flags $tok";; *) envs="$envs $tok";; esac; done
  env $envs python bench.py --no-configs --cpu-sample 0 $flags > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -3 gpurun_out/ab_$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$i.json"))
print("[$spec] value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["kernel_times_ms_per_proof"])
PY
done
