#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(time python bench.py) > gpurun_out/r2_bench_default_v2.json 2> gpurun_out/r2_bench_default_v2.err
tail -4 gpurun_out/r2_bench_default_v2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_default_v2.json"))
print("value %.1f e2e %.1f frac %.3f cpu %s launches %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"), d["gpu_launches"]))
print(d["kernel_times_ms_per_proof"], d["msm_2p17"], d["single_proof_latency_ms"], d["clocks"])
for k, v in d.get("configs", {}).get("note_shapes", {}).items(): print(k, round(v["proofs_per_s"], 1), round(v["roofline_frac"], 3))
for r in d.get("configs", {}).get("msm_sweep", []): print("msm", r["points"], round(r["gpu_ms"], 3), round(r["cpu_ms"], 1), r["bit_exact_vs_cpu"], round(r["frac_of_imad_roofline_survey_formula"], 3))
for r in d.get("configs", {}).get("ntt_sweep", []): print("ntt", r["size"], round(r["gpu_ms"], 4), round(r["cpu_ms"], 2), r["bit_exact_vs_cpu"])
print(d.get("configs", {}).get("batch_verification_g1_sum_1024_proofs"))
PY
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2_bench_ref_v2.json 2> gpurun_out/r2_bench_ref_v2.err
tail -c 600 gpurun_out/r2_bench_ref_v2.json
