"""Generates tests/golden/*.json from the Python big-int oracle.

The reference holds NO golden vectors for this path and cannot be run here (no
Rust toolchain), so these fixtures freeze the ORACLE's outputs (they pin regressions of the
oracle and give the GPU tests oracle-independent constants to compare with); they are not
outputs of the reference.  External known answers (Keccak-256, ChaCha20 RFC 8439, the coset
constants of the CAP on-chain verifier) are checked separately in tests/test_oracle_hash.py.

Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cap_b200 import synth  # noqa: E402
from oracle import bn254, msm, ntt, plonk  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % bn254.R


def hx(x):
    return hex(x)


def pt(p):
    return None if p is None else [hex(p[0]), hex(p[1])]


def main():
    rng = random.Random(20221017)
    # NTT vectors
    vecs = []
    for log_n in (3, 6):
        x = [rng.randrange(bn254.R) for _ in range(1 << log_n)]
        vecs.append({
            "log_n": log_n, "input": [hx(v) for v in x],
            "fft": [hx(v) for v in ntt.fft(x, log_n)], "ifft": [hx(v) for v in ntt.ifft(x, log_n)],
            "coset_fft": [hx(v) for v in ntt.coset_fft(x, log_n)], "coset_ifft": [hx(v) for v in ntt.coset_ifft(x, log_n)],
        })
    json.dump(vecs, open(os.path.join(HERE, "ntt.json"), "w"), indent=0)
    # MSM vectors over the synthetic SRS tau^i * G
    n = 40
    srs = bn254.srs_powers(TAU, n)
    cases = []
    for name, sc in [
        ("random", [rng.randrange(bn254.R) for _ in range(n)]),
        ("zeros_ones", [rng.randrange(2) for _ in range(n)]),
        ("max", [bn254.R - 1] * n),
        ("single", [0] * 7 + [rng.randrange(bn254.R)] + [0] * (n - 8)),
    ]:
        cases.append({"name": name, "scalars": [hx(s) for s in sc], "result": pt(msm.msm_arkworks(srs, sc))})
    json.dump({"tau": hx(TAU), "srs": [pt(p) for p in srs], "cases": cases}, open(os.path.join(HERE, "msm.json"), "w"), indent=0)
    # one complete proof on a 32-row circuit
    circ = synth.make_circuit(5, num_inputs=3, seed=5)
    pk = plonk.preprocess(circ, tau=TAU)
    bl = [rng.randrange(bn254.R) for _ in range(17)]
    proof = plonk.prove(circ, pk, bl, tau=TAU, ext_msg=b"golden", keep=True)
    assert plonk.verify(pk["vk"], plonk.public_input(circ), proof, TAU, ext_msg=b"golden")
    dbg = proof.pop("_debug")
    out = {
        "tau": hx(TAU), "log_n": 5, "num_inputs": 3, "seed": 5, "ext_msg": "golden",
        "blinders": [hx(b) for b in bl],
        "challenges": {k: hx(v) for k, v in dbg["challenges"].items()},
        "vk": {"selector_comms": [pt(p) for p in pk["vk"]["selector_comms"]], "sigma_comms": [pt(p) for p in pk["vk"]["sigma_comms"]]},
        "proof": {
            "wires_poly_comms": [pt(p) for p in proof["wires_poly_comms"]],
            "prod_perm_poly_comm": pt(proof["prod_perm_poly_comm"]),
            "split_quot_poly_comms": [pt(p) for p in proof["split_quot_poly_comms"]],
            "opening_proof": pt(proof["opening_proof"]),
            "shifted_opening_proof": pt(proof["shifted_opening_proof"]),
            "wires_evals": [hx(v) for v in proof["wires_evals"]],
            "wire_sigma_evals": [hx(v) for v in proof["wire_sigma_evals"]],
            "perm_next_eval": hx(proof["perm_next_eval"]),
        },
    }
    json.dump(out, open(os.path.join(HERE, "proof_n32.json"), "w"), indent=0)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
