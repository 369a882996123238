"""Oracle prover / verifier: acceptance and rejection in the style of the reference's own proof
tests (/root/reference/src/proof/transfer.rs:600-760: prove, verify Ok; wrong public input,
wrong proof, wrong extra data -> Err) plus the frozen golden proof."""
import random

import pytest

from cap_b200 import synth
from oracle import bn254 as B
from oracle import ntt, plonk

from conftest import TAU


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


@pytest.fixture(scope="module")
def small():
    circ = synth.make_circuit(5, num_inputs=3, seed=5)
    pk = plonk.preprocess(circ, tau=TAU)
    return circ, pk


def test_synthetic_circuit_is_satisfied(small):
    circ, _ = small
    assert plonk.check_gates(circ)
    # tampering one witness value breaks a gate
    bad = synth.SynthCircuit(circ.log_n, circ.num_inputs, circ.selectors, circ.wire_variables, list(circ.witness), circ.k)
    bad.witness[bad.wire_variables[4][10]] = (bad.witness[bad.wire_variables[4][10]] + 1) % B.R
    assert not plonk.check_gates(bad)
    # re-solved witness still satisfies the same circuit
    assert plonk.check_gates(circ.with_witness(77))


def test_grand_product_closes(small):
    circ, _ = small
    z = plonk.grand_product(circ, 12345, 67890)
    assert z[0] == 1 and len(z) == circ.n
    # the full product over all n rows is 1 when the copy constraints hold
    n = circ.n
    w = plonk.wire_evals(circ)
    ext = plonk.extended_id_permutation(circ)
    sig = plonk.sigma_evals(circ)
    a = b = 1
    for i in range(5):
        a = a * (w[i][n - 1] + 12345 * ext[i * n + n - 1] + 67890) % B.R
        b = b * (w[i][n - 1] + 12345 * sig[i][n - 1] + 67890) % B.R
    assert z[n - 1] * a % B.R == b


def test_prove_verify_and_tamper(small):
    circ, pk = small
    rng = random.Random(4)
    bl = [rng.randrange(B.R) for _ in range(17)]
    pub = plonk.public_input(circ)
    proof = plonk.prove(circ, pk, bl, tau=TAU, ext_msg=b"memo-key")
    assert plonk.verify(pk["vk"], pub, proof, TAU, ext_msg=b"memo-key")
    assert not plonk.verify(pk["vk"], pub, proof, TAU, ext_msg=b"memo-kez")
    assert not plonk.verify(pk["vk"], pub, proof, TAU, ext_msg=None)
    bad_pub = list(pub)
    bad_pub[1] = (bad_pub[1] + 1) % B.R
    assert not plonk.verify(pk["vk"], bad_pub, proof, TAU, ext_msg=b"memo-key")
    for key in ("perm_next_eval",):
        p2 = dict(proof)
        p2[key] = (p2[key] + 1) % B.R
        assert not plonk.verify(pk["vk"], pub, p2, TAU, ext_msg=b"memo-key")
    p2 = dict(proof)
    p2["wires_poly_comms"] = list(reversed(proof["wires_poly_comms"]))
    assert not plonk.verify(pk["vk"], pub, p2, TAU, ext_msg=b"memo-key")
    p2 = dict(proof)
    p2["opening_proof"] = proof["shifted_opening_proof"]
    assert not plonk.verify(pk["vk"], pub, p2, TAU, ext_msg=b"memo-key")
    # a different circuit's key does not verify the proof
    other = plonk.preprocess(synth.make_circuit(5, num_inputs=3, seed=6), tau=TAU)
    assert not plonk.verify(other["vk"], pub, proof, TAU, ext_msg=b"memo-key")


def test_commit_msm_equals_tau_evaluation(small):
    circ, pk = small
    srs = B.srs_powers(TAU, circ.n + 3)
    for p in (pk["selectors"][0], pk["sigmas"][4]):
        assert plonk.commit(p, srs=srs) == plonk.commit(p, tau=TAU)


def test_quotient_identity_holds_at_random_point(small):
    """t(x) Z_H(x) equals the gate + permutation numerator at a point outside both domains."""
    circ, pk = small
    rng = random.Random(8)
    bl = [rng.randrange(B.R) for _ in range(17)]
    proof = plonk.prove(circ, pk, bl, tau=TAU, keep=True)
    d = proof["_debug"]
    ch = d["challenges"]
    n = circ.n
    x = rng.randrange(B.R)
    ev = lambda p: ntt.poly_eval(p, x)
    w = [ev(p) for p in d["wire_polys"]]
    s = [ev(p) for p in pk["selectors"]]
    sg = [ev(p) for p in pk["sigmas"]]
    z, zw, pi, t = ev(d["z_poly"]), ntt.poly_eval(d["z_poly"], x * B.fr_root_of_unity(circ.log_n) % B.R), ev(d["pi_poly"]), ev(d["t_poly"])
    gate = (s[11] + pi + sum(s[i] * w[i] for i in range(4)) + s[4] * w[0] * w[1] + s[5] * w[2] * w[3]
            + s[12] * w[0] * w[1] * w[2] * w[3] * w[4] + sum(s[6 + i] * pow(w[i], 5, B.R) for i in range(4)) - s[10] * w[4]) % B.R
    r1, r2 = z, zw
    for j in range(5):
        r1 = r1 * (w[j] + ch["beta"] * circ.k[j] * x + ch["gamma"]) % B.R
        r2 = r2 * (w[j] + ch["beta"] * sg[j] + ch["gamma"]) % B.R
    zh = (pow(x, n, B.R) - 1) % B.R
    l1 = zh * pow(n * (x - 1) % B.R, -1, B.R) % B.R
    lhs = t * zh % B.R
    rhs = (gate + ch["alpha"] * (r1 - r2) + ch["alpha"] ** 2 * l1 * (z - 1)) % B.R
    assert lhs == rhs


def test_golden_proof(golden):
    g = golden["proof_n32"]
    circ = synth.make_circuit(g["log_n"], num_inputs=g["num_inputs"], seed=g["seed"])
    pk = plonk.preprocess(circ, tau=TAU)
    assert pk["vk"]["selector_comms"] == [_pt(p) for p in g["vk"]["selector_comms"]]
    assert pk["vk"]["sigma_comms"] == [_pt(p) for p in g["vk"]["sigma_comms"]]
    proof = plonk.prove(circ, pk, [int(b, 16) for b in g["blinders"]], tau=TAU, ext_msg=g["ext_msg"].encode(), keep=True)
    assert {k: hex(v) for k, v in proof["_debug"]["challenges"].items()} == g["challenges"]
    gp = g["proof"]
    assert proof["wires_poly_comms"] == [_pt(p) for p in gp["wires_poly_comms"]]
    assert proof["prod_perm_poly_comm"] == _pt(gp["prod_perm_poly_comm"])
    assert proof["split_quot_poly_comms"] == [_pt(p) for p in gp["split_quot_poly_comms"]]
    assert proof["opening_proof"] == _pt(gp["opening_proof"])
    assert proof["shifted_opening_proof"] == _pt(gp["shifted_opening_proof"])
    assert [hex(v) for v in proof["wires_evals"]] == gp["wires_evals"]
    assert [hex(v) for v in proof["wire_sigma_evals"]] == gp["wire_sigma_evals"]
    assert hex(proof["perm_next_eval"]) == gp["perm_next_eval"]


def test_batch_verification_accepts_valid_batches_and_rejects_one_bad_proof(small):
    """Batched check in the style of /root/reference/src/lib.rs:732-820 (txn_batch_verify over
    several notes, including notes of a second circuit): two multi-scalar sums and one pairing
    equation for the whole batch."""
    circ, pk = small
    circ2 = synth.make_circuit(4, num_inputs=2, seed=9)
    pk2 = plonk.preprocess(circ2, tau=TAU)
    rng = random.Random(31)
    inst = []
    for k, (c, key) in enumerate([(circ, pk), (circ.with_witness(3), pk), (circ2, pk2), (circ.with_witness(4), pk)]):
        msg = b"note-%d" % k
        proof = plonk.prove(c, key, [rng.randrange(B.R) for _ in range(17)], tau=TAU, ext_msg=msg)
        inst.append((key["vk"], plonk.public_input(c), proof, msg))
    rs = [1] + [rng.randrange(1, B.R) for _ in inst[1:]]
    assert plonk.batch_verify(inst, rs, tau=TAU)
    A_terms, B_terms = plonk.batch_verify_terms(inst, rs)
    # bases shared by the three proofs of the first key are merged: 18 vk comms + generator once
    assert len(A_terms) == 2 * len(inst)
    assert len(B_terms) == (18 + 13) * 2 + 13 * 2 + 1
    bad = list(inst)
    vk, pub, proof, msg = bad[2]
    bad[2] = (vk, [(pub[0] + 1) % B.R] + pub[1:], proof, msg)
    assert not plonk.batch_verify(bad, rs, tau=TAU)
    # each proof alone still verifies / fails as expected
    assert all(plonk.verify(v, p, pr, TAU, ext_msg=m) for v, p, pr, m in inst)
    assert not plonk.verify(*bad[2][:3], TAU, ext_msg=bad[2][3])


def test_evaluation_form_commitment_identity():
    """What the GPU's Lagrange commit key relies on: for a masked wire polynomial
    w(X) + (b0 + b1 X)(X^n - 1), sum_j w_j L_j(tau) - b0 - b1 tau + b0 tau^n + b1 tau^(n+1) equals
    its evaluation at tau, so committing from evaluations gives KZG10::commit's group element; and a
    gadget-like witness (unused inputs wired to the zero variable, boolean inputs) still satisfies
    the synthetic circuit."""
    circ = synth.make_circuit(6, num_inputs=3, seed=13, zero_inputs=0.4, bool_inputs=0.5)
    assert plonk.check_gates(circ) and plonk.check_gates(circ.with_witness(2))
    n = circ.n
    cells = [circ.witness[v] for col in circ.wire_variables for v in col]
    assert sum(1 for c in cells if c < 2) > len(cells) // 3
    w = B.fr_root_of_unity(circ.log_n)
    zh = (pow(TAU, n, B.R) - 1) % B.R
    lag = [zh * pow(w, j, B.R) % B.R * B.inv(n * (TAU - pow(w, j, B.R)) % B.R, B.R) % B.R for j in range(n)]
    evals = [circ.witness[v] for v in circ.wire_variables[0]]
    coeffs = ntt.ifft(evals, circ.log_n)
    b0, b1 = 1234567, 7654321
    masked = list(coeffs) + [0, 0]
    masked[0] = (masked[0] - b0) % B.R
    masked[1] = (masked[1] - b1) % B.R
    masked[n] = b0
    masked[n + 1] = b1
    at_tau = sum(c * pow(TAU, i, B.R) for i, c in enumerate(masked)) % B.R
    from_evals = (sum(e * l for e, l in zip(evals, lag)) - b0 - b1 * TAU + b0 * pow(TAU, n, B.R) + b1 * pow(TAU, n + 1, B.R)) % B.R
    assert at_tau == from_evals
