run() { timeout 400 python bench.py --no-extras $2 2>gpurun_out/b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['e2e']['value'],1))"; }
run base
CAPGPU_RED_SEG=64 run L64
CAPGPU_RED_SEG=128 run L128
CAPGPU_RED_SEG=16 run L16
run base2
