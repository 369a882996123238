"""One lockstep group of G proofs on one context (for `ncu --metrics gpu__time_duration.sum` launch
lists and `ncu --set full` captures): python tests/gpu_scripts/prof_group.py [G] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from cap_b200 import device, plonk

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
circ, circs, wires, pubs, bl = bench.build_workload("transfer_2x2")
ctx = device.Context(0)
ctx.set_group(G)
srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, bench.TAU)
pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
dw = [torch.from_numpy(w.view(np.int64)).cuda() for w in wires]
ptrs = [dw[i % 4].data_ptr() for i in range(G)]
pp = [pubs[i % 4] for i in range(G)]
bb = [bl[i % 4] for i in range(G)]
mm = [b"x"] * G
plonk.prove_batch_raw([ctx], pk, ptrs, pp, bb, mm, on_device=True)  # warm-up (tables, workspaces)
ctx.sync()
print("LAUNCHES_BEFORE", ctx.launch_count, flush=True)
for _ in range(reps):
    plonk.prove_batch_raw([ctx], pk, ptrs, pp, bb, mm, on_device=True)
ctx.sync()
print("LAUNCHES_AFTER", ctx.launch_count, flush=True)
