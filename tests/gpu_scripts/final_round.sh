#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
(time python bench.py --csv gpurun_out/r2_cap_benchmark.csv) > gpurun_out/r2_bench_default_v4.json 2> gpurun_out/r2_bench_default_v4.err
tail -4 gpurun_out/r2_bench_default_v4.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_default_v4.json"))
print("value %.1f e2e %.1f frac %.3f cpu %s launches %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value"), d["gpu_launches"]))
print(d["kernel_times_ms_per_proof"], d["msm_2p17"], d["single_proof_latency_ms"], d["clocks"])
print(d["ntt"])
for k, v in d.get("configs", {}).get("note_shapes", {}).items(): print(k, round(v["proofs_per_s"], 1), round(v["roofline_frac"], 3))
for r in d.get("configs", {}).get("msm_sweep", []): print("msm", r["points"], round(r["gpu_ms"], 3), round(r["cpu_ms"], 1), r["bit_exact_vs_cpu"], round(r["frac_of_imad_roofline_survey_formula"], 3))
PY
cat gpurun_out/r2_cap_benchmark.csv
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2_bench_ref_v4.json 2> gpurun_out/r2_bench_ref_v4.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_ref_v4.json')); print('reference', d['value'], d['cpu_baseline'])"
