set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_final.err
for w in mint transfer_3x5 transfer_5x5; do python bench.py --workload $w --no-extras > gpurun_out/bench_$w.json 2>> gpurun_out/bench_final.err; done
python bench.py --witness sparse --no-extras > gpurun_out/bench_sparse.json 2>> gpurun_out/bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python tests/gpu_scripts/prof_one.py --proofs 1 --msm17 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"msm_accumulate|ntt_reg_kernel|quotient_kernel" -c 14 -o /tmp/full python tests/gpu_scripts/prof_one.py --proofs 1 > /dev/null 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > gpurun_out/ncu_full_final_raw.csv
ncu -i /tmp/full.ncu-rep --page details --csv > gpurun_out/ncu_full_final_details.csv
python tests/gpu_scripts/sweep.py > gpurun_out/sweep_final.json 2>> gpurun_out/bench_final.err
tail -3 gpurun_out/bench_final.err
cat gpurun_out/bench_final.json | head -c 600
