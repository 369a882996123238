"""Dev script (run under gpurun): full-prover parity against the oracle on small domains,
oracle verification of a full-size proof, first proof timings."""
import json
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cap_b200 import device, field, plonk, synth  # noqa: E402
from oracle import bn254, plonk as oplonk  # noqa: E402

out = {}
ctx = device.Context(0)
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % bn254.R
rng = random.Random(99)


def blinders():
    return [rng.randrange(bn254.R) for _ in range(17)]


def compare(name, got, exp):
    if got != exp:
        nd = sum(1 for a, b in zip(got, exp) if a != b) + abs(len(got) - len(exp))
        print(f"  MISMATCH {name}: {nd} of {len(exp)} differ (len got {len(got)})", flush=True)
        return 1
    return 0


bad = 0
for log_n, nin in [(5, 3), (8, 5), (10, 27), (11, 27)]:
    t0 = time.time()
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=log_n)
    assert oplonk.check_gates(circ)
    n = circ.n
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    opk = oplonk.preprocess(circ, tau=TAU)
    sel, sig, sc, gc = pk.export()
    for s in range(13):
        bad += compare(f"selector poly {s}", field.fr_from_mont_array(sel[s]), opk["selectors"][s])
    for s in range(5):
        bad += compare(f"sigma poly {s}", field.fr_from_mont_array(sig[s]), opk["sigmas"][s])
    bad += compare("selector comms", pk.vk["selector_comms"], opk["vk"]["selector_comms"])
    bad += compare("sigma comms", pk.vk["sigma_comms"], opk["vk"]["sigma_comms"])
    bl = blinders()
    msg = b"ext-msg-%d" % log_n
    oproof = oplonk.prove(circ, opk, bl, tau=TAU, ext_msg=msg, keep=True)
    gproof = plonk.PlonkKzgSnark.prove(ctx, circ, pk, [bn254.to_mont(b, bn254.R) for b in bl], msg)
    dbg = oproof["_debug"]
    wp = plonk.debug_read(ctx, 0, 5 * (n + 2))
    bad += compare("wire polys", wp, [c for p in dbg["wire_polys"] for c in p])
    bad += compare("pi poly", plonk.debug_read(ctx, 8, n), dbg["pi_poly"])
    bad += compare("z evals", plonk.debug_read(ctx, 1, n), dbg["z_evals"])
    bad += compare("z poly", plonk.debug_read(ctx, 2, n + 3), dbg["z_poly"])
    tp = plonk.debug_read(ctx, 4, 8 * n)
    bad += compare("t poly", tp[:5 * n + 8], dbg["t_poly"])
    sp = plonk.debug_read(ctx, 9, 5 * (n + 3))
    exp_split = []
    for p in dbg["split"]:
        exp_split += list(p) + [0] * (n + 3 - len(p))
    bad += compare("split", sp, exp_split)
    bad += compare("lin poly", plonk.debug_read(ctx, 5, n + 3), dbg["lin"] + [0] * (n + 3 - len(dbg["lin"])))
    bad += compare("open poly", plonk.debug_read(ctx, 6, n + 3)[:n + 2], dbg["open_poly"] + [0] * (n + 2 - len(dbg["open_poly"])))
    bad += compare("shifted poly", plonk.debug_read(ctx, 7, n + 3)[:n + 2], dbg["shifted_poly"])
    for key in ["wires_poly_comms", "prod_perm_poly_comm", "split_quot_poly_comms", "opening_proof", "shifted_opening_proof",
                "wires_evals", "wire_sigma_evals", "perm_next_eval"]:
        if gproof[key] != oproof[key]:
            bad += 1
            print(f"  MISMATCH proof field {key}", flush=True)
    ok = oplonk.verify(opk["vk"], oplonk.public_input(circ), gproof, TAU, ext_msg=msg)
    print(f"log_n={log_n}: verify={ok} cumulative mismatches={bad} ({time.time() - t0:.1f}s)", flush=True)
    if not ok:
        bad += 1
    pk.close()
    srs.close()
out["parity_bad"] = bad

# ---- full-size proof: verified by the oracle verifier; timing
for name in ("transfer_2x2",):
    log_n, nin = synth.NOTE_SHAPES[name]
    t0 = time.time()
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=7)
    print(f"{name}: circuit built in {time.time() - t0:.1f}s", flush=True)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    t0 = time.time()
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    print(f"preprocess {time.time() - t0:.2f}s", flush=True)
    wires = plonk.wire_values(circ)
    pub = field.fr_to_mont_array(plonk.public_input(circ))
    blm = field.fr_raw_array([bn254.to_mont(b, bn254.R) for b in blinders()])
    proof = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, blm, b"bench")
    pd = plonk.proof_to_dict(proof)
    ok = oplonk.verify(pk.vk, plonk.public_input(circ), pd, TAU, ext_msg=b"bench")
    print(f"{name}: oracle verifier accepts: {ok}", flush=True)
    out[f"{name}_verify"] = ok
    l0 = ctx.launch_count
    ts = []
    for _ in range(8):
        t = time.perf_counter()
        plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, blm, b"bench")
        ts.append(time.perf_counter() - t)
    out[f"{name}_ms"] = min(ts) * 1e3
    out[f"{name}_launches_per_proof"] = (ctx.launch_count - l0) / 8
    print(f"{name}: {min(ts) * 1e3:.2f} ms/proof (min of 8), launches/proof {out[f'{name}_launches_per_proof']}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/check2.json", "w"), indent=1)
print(json.dumps(out))
