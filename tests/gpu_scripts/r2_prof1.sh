#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_prof_group_plain.log 2>&1
cat gpurun_out/r2_prof_group_plain.log | tail -3
# launch list of the second group only: skip the launches of setup + warm-up
SKIP=$(grep LAUNCHES_BEFORE gpurun_out/r2_prof_group_plain.log | awk '{print $2}')
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 200 --csv --log-file gpurun_out/r2_launches_group8.csv python tests/gpu_scripts/prof_group.py 8 1 > gpurun_out/r2_prof_group_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_group_ncu.log
python bench.py --ctxs 4 --group 8 --steps 5 --warmup 2 --cpu-sample 0 > gpurun_out/r2_bench_extras.json 2> gpurun_out/r2_bench_extras.err
tail -c 3000 gpurun_out/r2_bench_extras.json
