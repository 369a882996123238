#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "64 4 8" "128 4 8" "256 4 8" "256 4 16" "256 6 8"; do
set -- $cfg
python bench.py --steps 4 --warmup 3 --no-configs --no-extras --cpu-sample 0 --batch $1 --ctxs $2 --group $3 > gpurun_out/r2b_bench4.json 2> gpurun_out/r2b_bench4.err || tail -3 gpurun_out/r2b_bench4.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench4.json"))
print("batch $1 ctxs $2 group $3: value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), d["clocks"])
PY
done
ncu --set full --clock-control none -k regex:"^msm_accumulate$" --launch-skip 5 -c 4 -o gpurun_out/r2_ncu_acc2 python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_ncu_acc2.log 2>&1
ncu -i gpurun_out/r2_ncu_acc2.ncu-rep --page raw --csv > gpurun_out/r2_ncu_acc2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_ncu_acc2.ncu-rep --page details --csv > gpurun_out/r2_ncu_acc2_details.csv 2>/dev/null
rm -f gpurun_out/r2_ncu_acc2.ncu-rep
ls -la gpurun_out/r2_ncu_acc2*
