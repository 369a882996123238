#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_tests5.log 2>&1
tail -4 gpurun_out/r2_tests5.log
python bench.py --fixture tests/fixtures/oracle_transfer_n64.capfix --steps 2 --warmup 1 > gpurun_out/r2_bench_fixture.json 2> gpurun_out/r2_bench_fixture.err
tail -c 600 gpurun_out/r2_bench_fixture.json; tail -3 gpurun_out/r2_bench_fixture.err
