run() { timeout 400 python bench.py --no-extras $2 2>gpurun_out/b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['e2e']['value'],1))"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run a; run b; run c
timeout 300 python tests/gpu_scripts/msm_tune.py 2>&1 | tail -1
