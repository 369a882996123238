"""Profiling target (run under ncu via gpurun): sets up the transfer_2x2 workload, then proves
`--proofs` notes on ONE context inside a cudaProfilerStart/Stop window, so
`ncu --profile-from-start off` sees exactly the steady-state kernels of the hot path."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from cap_b200 import device, field, plonk, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--proofs", type=int, default=1)
ap.add_argument("--workload", default="transfer_2x2")
ap.add_argument("--msm17", action="store_true", help="also run one standalone 2^17-point MSM inside the window")
args = ap.parse_args()

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
log_n, nin = synth.NOTE_SHAPES[args.workload]
circ = synth.make_circuit(log_n, num_inputs=nin, seed=7)
ctx = device.Context(0)
ctx.set_latency_mode(True)  # the single-note schedule (flat accumulation, quad tiles)
srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
wires = plonk.wire_values(circ)
pub = field.fr_to_mont_array(plonk.public_input(circ))
bl = field.fr_raw_array(list(range(1, 18)))
plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"prof")  # warm-up (tables, workspaces)
srs17 = sc17 = None
if args.msm17:
    srs17 = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=1 << 17)
    sc17 = np.random.default_rng(1).integers(0, 1 << 60, size=(1 << 17, 4), dtype=np.uint64)
    srs17.msm(sc17, mont=False)
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.proofs):
    plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"prof")
if args.msm17:
    srs17.msm(sc17, mont=False)
ctx.sync()
torch.cuda.profiler.stop()
print("launches so far:", ctx.launch_count)
