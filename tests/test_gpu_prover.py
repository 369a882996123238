"""GPU parity tests for the whole prover through the C ABI (capgpu_preprocess, capgpu_prove, the
round API): bit-exact against the oracle on small domains (every intermediate polynomial, every
commitment and evaluation), the frozen golden proof, and acceptance by the oracle's verifier at
the benchmark size.  Mirrors /root/reference/src/proof/transfer.rs:600-760 (prove then verify;
wrong inputs are rejected) at the boundary this backend replaces."""
import ctypes
import random
from ctypes import byref, c_void_p

import numpy as np
import pytest

from cap_b200 import _lib, field, plonk, synth
from cap_b200.device import _ptr
from oracle import bn254 as B
from oracle import plonk as oplonk
from oracle.transcript import SolidityTranscript

from conftest import TAU

pytestmark = pytest.mark.gpu


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


def _mont(xs):
    return [B.to_mont(x, B.R) for x in xs]


@pytest.mark.parametrize("log_n,nin", [(5, 3), (7, 0), (9, 27), (11, 27)])
def test_prover_matches_oracle(ctx, log_n, nin):
    rng = random.Random(log_n)
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=log_n)
    n = circ.n
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    opk = oplonk.preprocess(circ, tau=TAU)
    sel, sig, _, _ = pk.export()
    assert [field.fr_from_mont_array(s) for s in sel] == opk["selectors"]
    assert [field.fr_from_mont_array(s) for s in sig] == opk["sigmas"]
    assert pk.vk["selector_comms"] == opk["vk"]["selector_comms"]
    assert pk.vk["sigma_comms"] == opk["vk"]["sigma_comms"]
    bl = [rng.randrange(B.R) for _ in range(17)]
    msg = b"memo-ver-key-%d" % log_n
    op = oplonk.prove(circ, opk, bl, tau=TAU, ext_msg=msg, keep=True)
    gp = plonk.PlonkKzgSnark.prove(ctx, circ, pk, _mont(bl), msg)
    d = op.pop("_debug")
    assert plonk.debug_read(ctx, 0, 5 * (n + 2)) == [c for p in d["wire_polys"] for c in p]
    assert plonk.debug_read(ctx, 8, n) == d["pi_poly"]
    assert plonk.debug_read(ctx, 1, n) == d["z_evals"]
    assert plonk.debug_read(ctx, 2, n + 3) == d["z_poly"]
    tp = plonk.debug_read(ctx, 4, 8 * n)
    assert tp[:5 * n + 8] == d["t_poly"] and not any(tp[5 * n + 8:])
    exp_split = []
    for p in d["split"]:
        exp_split += list(p) + [0] * (n + 3 - len(p))
    assert plonk.debug_read(ctx, 9, 5 * (n + 3)) == exp_split
    assert plonk.debug_read(ctx, 5, n + 3) == d["lin"] + [0] * (n + 3 - len(d["lin"]))
    assert plonk.debug_read(ctx, 6, n + 3)[:n + 2] == d["open_poly"] + [0] * (n + 2 - len(d["open_poly"]))
    assert plonk.debug_read(ctx, 7, n + 3)[:n + 2] == d["shifted_poly"]
    assert gp == op
    assert oplonk.verify(opk["vk"], oplonk.public_input(circ), gp, TAU, ext_msg=msg)
    assert not oplonk.verify(opk["vk"], oplonk.public_input(circ), gp, TAU, ext_msg=msg + b"x")
    pk.close()
    srs.close()


def test_golden_proof(ctx, golden):
    g = golden["proof_n32"]
    circ = synth.make_circuit(g["log_n"], num_inputs=g["num_inputs"], seed=g["seed"])
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, int(g["tau"], 16))
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    assert pk.vk["selector_comms"] == [_pt(p) for p in g["vk"]["selector_comms"]]
    assert pk.vk["sigma_comms"] == [_pt(p) for p in g["vk"]["sigma_comms"]]
    proof = plonk.PlonkKzgSnark.prove(ctx, circ, pk, _mont([int(b, 16) for b in g["blinders"]]), g["ext_msg"].encode())
    gp = g["proof"]
    assert proof["wires_poly_comms"] == [_pt(p) for p in gp["wires_poly_comms"]]
    assert proof["prod_perm_poly_comm"] == _pt(gp["prod_perm_poly_comm"])
    assert proof["split_quot_poly_comms"] == [_pt(p) for p in gp["split_quot_poly_comms"]]
    assert proof["opening_proof"] == _pt(gp["opening_proof"])
    assert proof["shifted_opening_proof"] == _pt(gp["shifted_opening_proof"])
    assert [hex(v) for v in proof["wires_evals"]] == gp["wires_evals"]
    assert [hex(v) for v in proof["wire_sigma_evals"]] == gp["wire_sigma_evals"]
    assert hex(proof["perm_next_eval"]) == gp["perm_next_eval"]
    pk.close()
    srs.close()


def test_round_api_and_uploaded_key_reproduce_the_fused_proof(ctx):
    """A host that keeps its own transcript (the Rust shim) drives the five rounds itself; a
    ProvingKey uploaded in jf-plonk's coefficient form behaves like the preprocessed one."""
    rng = random.Random(21)
    circ = synth.make_circuit(8, num_inputs=4, seed=3)
    n = circ.n
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk0 = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    sel, sig, sc, gc = pk0.export()
    pk = plonk.PlonkKzgSnark.upload_proving_key(ctx, srs, circ.log_n, circ.num_inputs, sel, sig, circ.k, sc, gc)
    bl = _mont([rng.randrange(B.R) for _ in range(17)])
    fused = plonk.PlonkKzgSnark.prove(ctx, circ, pk0, bl, b"ext")
    assert plonk.PlonkKzgSnark.prove(ctx, circ, pk, bl, b"ext") == fused

    lib = ctx.lib
    wires = plonk.wire_values(circ)
    pubv = plonk.public_input(circ)
    pub = field.fr_to_mont_array(pubv)
    blm = field.fr_raw_array(bl)
    tr = SolidityTranscript()
    tr.append_message(b"ext")
    tr.append_vk_and_pub_input(pk.vk, pubv)
    job = c_void_p()
    _lib.check(lib.capgpu_job_begin(ctx.h, pk.h, _ptr(wires), _ptr(pub), byref(job)), ctx.h)
    # out-of-order round is refused
    tmp = np.zeros((5, 8), dtype=np.uint64)
    assert lib.capgpu_job_round3(job, _ptr(blm), _ptr(blm), _ptr(tmp)) == -5
    c1 = np.zeros((5, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round1(job, _ptr(blm[:10].copy()), _ptr(c1)), ctx.h)
    tr.append_commitments(field.g1_from_mont_array(c1))
    beta, gamma = tr.get_and_append_challenge(), tr.get_and_append_challenge()
    c2 = np.zeros((1, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round2(job, _ptr(field.fr_to_mont_array([beta])), _ptr(field.fr_to_mont_array([gamma])), _ptr(blm[10:13].copy()), _ptr(c2)), ctx.h)
    tr.append_commitments(field.g1_from_mont_array(c2))
    alpha = tr.get_and_append_challenge()
    c3 = np.zeros((5, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round3(job, _ptr(field.fr_to_mont_array([alpha])), _ptr(blm[13:17].copy()), _ptr(c3)), ctx.h)
    tr.append_commitments(field.g1_from_mont_array(c3))
    zeta = tr.get_and_append_challenge()
    ev = np.zeros((10, 4), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round4(job, _ptr(field.fr_to_mont_array([zeta])), _ptr(ev)), ctx.h)
    evals = field.fr_from_mont_array(ev)
    tr.append_proof_evaluations(evals[:5], evals[5:9], evals[9])
    v = tr.get_and_append_challenge()
    c5 = np.zeros((2, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round5(job, _ptr(field.fr_to_mont_array([v])), _ptr(c5)), ctx.h)
    lib.capgpu_job_end(job)
    assert field.g1_from_mont_array(c1) == fused["wires_poly_comms"]
    assert field.g1_from_mont_array(c2)[0] == fused["prod_perm_poly_comm"]
    assert field.g1_from_mont_array(c3) == fused["split_quot_poly_comms"]
    assert evals[:5] == fused["wires_evals"] and evals[5:9] == fused["wire_sigma_evals"] and evals[9] == fused["perm_next_eval"]
    assert field.g1_from_mont_array(c5) == [fused["opening_proof"], fused["shifted_opening_proof"]]
    pk.close()
    pk0.close()
    srs.close()


def test_error_behaviour(ctx):
    circ = synth.make_circuit(6, num_inputs=2, seed=1)
    # SRS one point short of domain + 3 (src/utils/mod.rs:109-113: max_degree = domain + 2)
    short = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 1, TAU)
    with pytest.raises(plonk.PlonkError):
        plonk.PlonkKzgSnark.preprocess(ctx, short, circ)
    short.close()
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    bl = _mont(list(range(1, 18)))
    good = plonk.PlonkKzgSnark.prove(ctx, circ, pk, bl)
    assert oplonk.verify(pk.vk, plonk.public_input(circ), good, TAU)
    # an unsatisfied witness makes the quotient a non-polynomial: jf-plonk reports
    # WrongQuotientPolyDegree, the ABI returns CAPGPU_ERR_DEGREE
    bad = synth.SynthCircuit(circ.log_n, circ.num_inputs, circ.selectors, circ.wire_variables, list(circ.witness), circ.k)
    v = bad.wire_variables[4][circ.num_inputs + 3]
    bad.witness[v] = (bad.witness[v] + 1) % B.R
    with pytest.raises(plonk.PlonkError, match="degree"):
        plonk.PlonkKzgSnark.prove(ctx, bad, pk, bl)
    # the ctx stays usable afterwards
    assert plonk.PlonkKzgSnark.prove(ctx, circ, pk, bl) == good
    # argument checking of the newer entry points: null handles / buffers are CAPGPU_ERR_ARG, never a crash
    lib = ctx.lib
    assert lib.capgpu_pk_lagrange(None, 1) == -2
    assert lib.capgpu_pk_lagrange_export(ctx.h, pk.h, None, 4) == -2
    out = np.zeros(8, dtype=np.uint64)
    assert lib.capgpu_msm_g1_adhoc(ctx.h, None, None, 3, 1, _ptr(out)) == -2
    assert lib.capgpu_msm_g1_adhoc(ctx.h, None, None, 0, 1, _ptr(out)) == 0 and not out.any()
    assert lib.capgpu_prove_batch(None, 0, pk.h, 0, None, None, None, None, None, None, None) == -2
    pts = np.zeros((pk.n + 4, 8), dtype=np.uint64)
    assert lib.capgpu_pk_lagrange_export(ctx.h, pk.h, _ptr(pts), pk.n + 5) != 0  # more than n + 4 bases
    pk.close()
    srs.close()


def test_transfer_shape_proofs_are_accepted(ctx):
    """BASELINE config 1 shape (n = 2^15, 27 public inputs): proofs for several witnesses under one
    proving key are accepted by the oracle verifier, tampered ones rejected, proving is deterministic."""
    log_n, nin = synth.NOTE_SHAPES["transfer_2x2"]
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=7)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    rng = random.Random(15)
    for seed in (None, 1):
        c = circ if seed is None else circ.with_witness(seed)
        bl = _mont([rng.randrange(B.R) for _ in range(17)])
        proof = plonk.PlonkKzgSnark.prove(ctx, c, pk, bl, b"note")
        pub = plonk.public_input(c)
        assert oplonk.verify(pk.vk, pub, proof, TAU, ext_msg=b"note")
        if seed is None:
            # the reference verifier's own check: BN254 pairing equation, no trapdoor on the G1 side
            from oracle import pairing
            assert oplonk.verify(pk.vk, pub, proof, ext_msg=b"note", g2_tau=pairing.g2_mul(pairing.G2_GEN, TAU))
        assert plonk.PlonkKzgSnark.prove(ctx, c, pk, bl, b"note") == proof
        bad = dict(proof)
        bad["wires_evals"] = [proof["wires_evals"][1], proof["wires_evals"][0]] + proof["wires_evals"][2:]
        assert not oplonk.verify(pk.vk, pub, bad, TAU, ext_msg=b"note")
        assert not oplonk.verify(pk.vk, pub[:-1] + [(pub[-1] + 1) % B.R], proof, TAU, ext_msg=b"note")
        # commitments equal p(tau) G for the polynomials left on the device
        wp = plonk.debug_read(ctx, 0, 5 * (circ.n + 2))
        from oracle.msm import kzg_commit_tau
        assert kzg_commit_tau(wp[:circ.n + 2], TAU) == proof["wires_poly_comms"][0]
    pk.close()
    srs.close()


def test_device_resident_and_latency_mode_give_identical_proofs(ctx):
    """capgpu_prove_dev (witness columns already in HBM) and the low-latency MSM schedule are
    scheduling choices only: the proof bytes are identical to the default path."""
    import torch
    circ = synth.make_circuit(10, num_inputs=5, seed=4)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    wires = plonk.wire_values(circ)
    pub = field.fr_to_mont_array(plonk.public_input(circ))
    bl = field.fr_raw_array(_mont(list(range(3, 20))))
    base = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"x")
    lib = ctx.lib
    assert lib.capgpu_ctx_set_latency_mode(ctx.h, 1) == 0
    lat = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"x")
    assert lib.capgpu_ctx_set_latency_mode(ctx.h, 0) == 0
    assert bytes(lat) == bytes(base)
    dw = torch.from_numpy(wires.view(np.int64)).cuda()
    msg = (ctypes.c_uint8 * 1).from_buffer_copy(b"x")
    out = _lib.Proof()
    _lib.check(lib.capgpu_prove_dev(ctx.h, pk.h, c_void_p(dw.data_ptr()), _ptr(pub), _ptr(bl), msg, 1, byref(out)), ctx.h)
    assert bytes(out) == bytes(base)
    pk.close()
    srs.close()


def test_prove_batch_matches_single_proofs(ctx):
    """capgpu_prove_batch (worker thread per context inside the library) returns, note for note, the
    proof the single-note call returns; a bad witness fails only its own slot."""
    from cap_b200 import device
    circ = synth.make_circuit(9, num_inputs=4, seed=2)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    ctxs = [ctx, device.Context(0), device.Context(0)]
    circs = [circ] + [circ.with_witness(s) for s in range(1, 7)]
    wires = [plonk.wire_values(c) for c in circs]
    pubs = [field.fr_to_mont_array(plonk.public_input(c)) for c in circs]
    rng = random.Random(6)
    bls = [field.fr_raw_array(_mont([rng.randrange(B.R) for _ in range(17)])) for _ in circs]
    msgs = [b"note-%d" % i for i in range(len(circs))]
    proofs, status = plonk.prove_batch_raw(ctxs, pk, [w.ctypes.data for w in wires], pubs, bls, msgs)
    assert status == [0] * len(circs)
    for i in range(len(circs)):
        single = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires[i], pubs[i], bls[i], msgs[i])
        assert bytes(proofs[i]) == bytes(single)
        assert oplonk.verify(pk.vk, plonk.public_input(circs[i]), plonk.proof_to_dict(proofs[i]), TAU, ext_msg=msgs[i])
    bad = wires[3].copy()
    bad[4, circ.num_inputs + 1, 0] ^= 1
    with pytest.raises(plonk.PlonkError):
        plonk.prove_batch_raw(ctxs, pk, [w.ctypes.data for w in wires[:3]] + [bad.ctypes.data], pubs[:4], bls[:4], msgs[:4])
    for c in ctxs[1:]:
        c.close()
    pk.close()
    srs.close()


@pytest.mark.parametrize("group,n_ctxs,count", [(8, 1, 5), (3, 2, 7), (2, 2, 1), (64, 1, 9)])
def test_lockstep_groups_equal_single_proofs(ctx, group, n_ctxs, count):
    """A context proving `group` notes in lockstep (one launch per round for the whole group; SURVEY
    8f N2) returns, note for note, the bytes of the single-note call -- for full, partial and
    single-note groups, host and device-resident inputs; an unsatisfied witness fails its own slot
    (CAPGPU_ERR_DEGREE) and nothing else."""
    import torch
    from cap_b200 import device
    circ = synth.make_circuit(8, num_inputs=3, seed=12)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    ctxs = [device.Context(0) for _ in range(n_ctxs)]
    for c in ctxs:
        c.set_group(group)
    circs = [circ] + [circ.with_witness(s) for s in range(1, count)]
    wires = [plonk.wire_values(c) for c in circs]
    pubs = [field.fr_to_mont_array(plonk.public_input(c)) for c in circs]
    rng = random.Random(group * 100 + count)
    bls = [field.fr_raw_array(_mont([rng.randrange(B.R) for _ in range(17)])) for _ in circs]
    msgs = [b"g-%d" % i if i % 3 else b"" for i in range(count)]
    singles = [bytes(plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires[i], pubs[i], bls[i], msgs[i])) for i in range(count)]
    proofs, status = plonk.prove_batch_raw(ctxs, pk, [w.ctypes.data for w in wires], pubs, bls, msgs)
    assert status == [0] * count
    assert [bytes(p) for p in proofs] == singles
    dw = [torch.from_numpy(w.view(np.int64)).cuda() for w in wires]
    proofs, status = plonk.prove_batch_raw(ctxs, pk, [t.data_ptr() for t in dw], pubs, bls, msgs, on_device=True)
    assert status == [0] * count and [bytes(p) for p in proofs] == singles
    if count >= 3:
        bad = wires[1].copy()
        bad[4, circ.num_inputs + 1, 0] ^= 1
        ws = [wires[0], bad] + wires[2:]
        proofs, status = plonk.prove_batch_raw(ctxs, pk, [w.ctypes.data for w in ws], pubs, bls, msgs, raise_on_error=False)
        assert status == [0, -3] + [0] * (count - 2)
        assert [bytes(p) for i, p in enumerate(proofs) if i != 1] == [s for i, s in enumerate(singles) if i != 1]
    for c in ctxs:
        c.close()
    pk.close()
    srs.close()


def test_proving_queue_submit_poll_wait(ctx):
    """capgpu_submit copies its inputs (the caller's buffers are scribbled over right after the
    call), tickets complete in any wait order, poll turns true, a bad note reports its own status,
    and every proof equals the synchronous single-note proof."""
    import time
    from cap_b200 import device
    circ = synth.make_circuit(9, num_inputs=4, seed=5)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    count = 11
    circs = [circ] + [circ.with_witness(s) for s in range(1, count)]
    wires = [plonk.wire_values(c) for c in circs]
    pubs = [field.fr_to_mont_array(plonk.public_input(c)) for c in circs]
    rng = random.Random(77)
    bls = [field.fr_raw_array(_mont([rng.randrange(B.R) for _ in range(17)])) for _ in circs]
    msgs = [b"q-%d" % i for i in range(count)]
    singles = [bytes(plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires[i], pubs[i], bls[i], msgs[i])) for i in range(count)]
    ctxs = [device.Context(0), device.Context(0)]
    for c in ctxs:
        c.set_group(4)
    q = plonk.ProvingQueue(ctxs, pk, ring_slots=3)  # fewer slots than notes: submit exercises back-pressure
    tickets = []
    for i in range(count):
        w, p, b = wires[i].copy(), pubs[i].copy(), bls[i].copy()
        if i == 6:
            w[4, circ.num_inputs + 2, 0] ^= 1  # unsatisfied witness
        tickets.append(q.submit(w, p, b, msgs[i]))
        w[:] = 0xDEADBEEF
        p[:] = 1
        b[:] = 2
    deadline = time.time() + 60
    while not q.poll(tickets[-1]):
        assert time.time() < deadline
        time.sleep(0.001)
    for i in reversed(range(count)):
        if i == 6:
            with pytest.raises(plonk.PlonkError, match="degree"):
                q.wait(tickets[i])
        else:
            assert bytes(q.wait(tickets[i])) == singles[i], i
    st = q.stats()
    assert st["submitted"] == st["completed"] == count and 1 <= st["groups"] <= count
    assert ctx.lib.capgpu_poll(q.h, tickets[0], byref(ctypes.c_int())) == -2  # a ticket is consumed by wait
    q.close()
    for c in ctxs:
        c.close()
    pk.close()
    srs.close()


def test_failed_round_releases_the_context(ctx):
    """a round that fails (unsatisfied circuit -> CAPGPU_ERR_DEGREE in round 3) must not
    wedge the context: a fresh capgpu_job_begin succeeds without capgpu_job_end on the failed job."""
    circ = synth.make_circuit(6, num_inputs=2, seed=8)
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    lib = ctx.lib
    wires = plonk.wire_values(circ)
    bad = wires.copy()
    bad[4, circ.num_inputs + 1, 0] ^= 1
    pub = field.fr_to_mont_array(plonk.public_input(circ))
    bl = field.fr_raw_array(_mont(list(range(5, 22))))
    one = field.fr_to_mont_array([7])
    job = c_void_p()
    _lib.check(lib.capgpu_job_begin(ctx.h, pk.h, _ptr(bad), _ptr(pub), byref(job)), ctx.h)
    c5 = np.zeros((5, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_job_round1(job, _ptr(bl[:10].copy()), _ptr(c5)), ctx.h)
    _lib.check(lib.capgpu_job_round2(job, _ptr(one), _ptr(one), _ptr(bl[10:13].copy()), _ptr(c5)), ctx.h)
    assert lib.capgpu_job_round3(job, _ptr(one), _ptr(bl[13:].copy()), _ptr(c5)) == -3
    assert lib.capgpu_job_round4(job, _ptr(one), _ptr(c5)) == -5  # the failed job is closed
    job2 = c_void_p()
    _lib.check(lib.capgpu_job_begin(ctx.h, pk.h, _ptr(wires), _ptr(pub), byref(job2)), ctx.h)
    lib.capgpu_job_end(job2)
    good = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"")
    assert oplonk.verify(pk.vk, plonk.public_input(circ), plonk.proof_to_dict(good), TAU)
    pk.close()
    srs.close()


@pytest.mark.parametrize("log_n,nin", [(6, 2), (9, 27), (11, 5)])
def test_quotient_domains_agree(ctx, log_n, nin, monkeypatch):
    """The 6n-point quotient domain (three 2n-point cosets; the default from n = 64) and the 8n-point coset jf-plonk uses
    (CAPGPU_QDOMAIN=8) interpolate the same quotient: equal coefficients, equal proofs, and the same rejection of a bad witness."""
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=21)
    n = circ.n
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    rng = random.Random(log_n)
    bl = _mont([rng.randrange(B.R) for _ in range(17)])
    bad = synth.SynthCircuit(circ.log_n, circ.num_inputs, circ.selectors, circ.wire_variables, list(circ.witness), circ.k)
    v = bad.wire_variables[4][circ.num_inputs + 3]
    bad.witness[v] = (bad.witness[v] + 1) % B.R
    proofs, quotients, vk = [], [], None
    for q in ("6", "8"):
        monkeypatch.setenv("CAPGPU_QDOMAIN", q)
        pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
        vk = pk.vk
        proofs.append(plonk.PlonkKzgSnark.prove(ctx, circ, pk, bl, b"domains"))
        t = plonk.debug_read(ctx, 4, 8 * n)
        assert len(t) == (6 if q == "6" else 8) * n and not any(t[5 * n + 8:])
        quotients.append(t[:5 * n + 8])
        with pytest.raises(plonk.PlonkError, match="degree"):
            plonk.PlonkKzgSnark.prove(ctx, bad, pk, bl, b"domains")
        pk.close()
    assert quotients[0] == quotients[1] and proofs[0] == proofs[1]
    assert oplonk.verify(vk, plonk.public_input(circ), proofs[0], TAU, ext_msg=b"domains")
    srs.close()


@pytest.mark.parametrize("log_n", [3, 6, 9])
def test_lagrange_commit_key(ctx, log_n):
    """The derived evaluation-form commit key is [L_j(tau) G] followed by P_0, P_1, P_n, P_{n+1}."""
    circ = synth.make_circuit(log_n, num_inputs=2, seed=3)
    n = circ.n
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    got = pk.lagrange_bases()
    w = B.fr_root_of_unity(log_n)
    zh_over_n = (pow(TAU, n, B.R) - 1) * B.inv(n, B.R) % B.R
    lag = [zh_over_n * pow(w, j, B.R) % B.R * B.inv((TAU - pow(w, j, B.R)) % B.R, B.R) % B.R for j in range(n)]
    assert sum(lag) % B.R == 1
    idx = range(n) if n <= 64 else random.Random(1).sample(range(n), 24)
    for j in idx:
        assert got[j] == B.g1_mul(B.G1_GEN, lag[j]), j
    mono = [B.g1_mul(B.G1_GEN, pow(TAU, e, B.R)) for e in (0, 1, n, n + 1)]
    assert got[n:] == mono
    pk.close()
    srs.close()


def test_lagrange_and_coefficient_commitments_agree_on_sparse_witness(ctx):
    """Wire commitments from evaluations (zero / boolean cells skipped or single-window) equal the
    coefficient-form KZG commitments and the oracle's proof."""
    log_n = 10
    circ = synth.make_circuit(log_n, num_inputs=5, seed=21, zero_inputs=0.45, bool_inputs=0.5)
    n = circ.n
    cells = [circ.witness[v] for col in circ.wire_variables for v in col]
    assert sum(1 for c in cells if c < 2) > len(cells) // 3
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    opk = oplonk.preprocess(circ, tau=TAU)
    bl = [random.Random(5).randrange(B.R) for _ in range(17)]
    op = oplonk.prove(circ, opk, bl, tau=TAU, ext_msg=b"sparse")
    a = plonk.PlonkKzgSnark.prove(ctx, circ, pk, _mont(bl), b"sparse")
    pk.set_lagrange(False)
    b = plonk.PlonkKzgSnark.prove(ctx, circ, pk, _mont(bl), b"sparse")
    pk.set_lagrange(True)
    assert a == b == op
    # all-zero witness columns and zero blinders: commitments are the point at infinity
    zero = np.zeros((5, n, 4), dtype=np.uint64)
    bl0 = np.zeros((17, 4), dtype=np.uint64)
    pub = field.fr_to_mont_array([0] * 5)
    try:
        p = plonk.proof_to_dict(plonk.PlonkKzgSnark.prove_raw(ctx, pk, zero, pub, bl0, None))
        assert p["wires_poly_comms"] == [None] * 5
    except plonk.PlonkError:
        pass  # the all-zero assignment need not satisfy the random circuit (quotient degree check)
    pk.close()
    srs.close()


@pytest.mark.parametrize("shape", ["mint", "transfer_2x2", "transfer_3x5", "freeze_5", "transfer_5x5"])
def test_note_shapes_match_the_c_restatement_at_full_size(ctx, shape):
    """BASELINE configs 1-3 (n = 2^14 .. 2^17): key and proof from the GPU equal, byte for byte, what the C
    restatement of the arkworks / jf-plonk algorithms computes on the CPU from the same SRS, circuit,
    witness and blinders; the Python verifier accepts the proof."""
    import os
    from oracle import cpu
    log_n, nin = synth.NOTE_SHAPES[shape]
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=11, zero_inputs=0.2, bool_inputs=0.2)
    n = circ.n
    threads = os.cpu_count() or 1
    srs = plonk.PlonkKzgSnark.universal_setup(ctx, n + 2, TAU)
    pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
    srs_xy = cpu.srs(field.fr_to_mont_array([TAU])[0], n + 3, threads)
    assert np.array_equal(srs.export(), srs_xy)
    sel_e = np.stack([field.fr_to_mont_array(s) for s in circ.selectors])
    sig_e = np.stack([field.fr_to_mont_array(s) for s in plonk.sigma_evals(circ)])
    sel, sig, sc, gc = cpu.preprocess(log_n, sel_e, sig_e, srs_xy, nthreads=threads)
    gsel, gsig, gsc, ggc = pk.export()
    assert np.array_equal(gsel, sel) and np.array_equal(gsig, sig)
    assert np.array_equal(gsc, sc) and np.array_equal(ggc, gc)
    wires = plonk.wire_values(circ)
    pub = field.fr_to_mont_array(plonk.public_input(circ))
    bl = field.fr_raw_array(_mont([random.Random(log_n).randrange(B.R) for _ in range(17)]))
    rc, cp = cpu.prove(log_n, nin, sel, sig, sig_e, field.fr_to_mont_array(circ.k), srs_xy, sc, gc, wires, pub, bl, b"full-size", nthreads=threads)
    assert rc == 0
    gp = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, b"full-size")
    assert bytes(gp) == bytes(cp)
    assert oplonk.verify(pk.vk, plonk.public_input(circ), plonk.proof_to_dict(gp), TAU, ext_msg=b"full-size")
    pk.close()
    srs.close()


def test_batch_verification_sums_on_the_gpu(ctx):
    """The batched verifier's two aggregated commitment sums (txn_batch_verify,
    /root/reference/src/lib.rs:517) computed by capgpu_msm_g1_adhoc over GPU-made proofs of two
    note types equal the oracle's, and the batch passes / fails the final check accordingly."""
    from cap_b200 import device
    rng = random.Random(41)
    inst = []
    keys = []
    for log_n, nin, seed in ((9, 6, 1), (8, 3, 2)):
        circ = synth.make_circuit(log_n, num_inputs=nin, seed=seed)
        srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
        pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
        keys.append((srs, pk))
        for k in range(3):
            c = circ if k == 0 else circ.with_witness(k)
            msg = b"batch-%d-%d" % (log_n, k)
            proof = plonk.PlonkKzgSnark.prove(ctx, c, pk, _mont([rng.randrange(B.R) for _ in range(17)]), msg)
            inst.append((pk.vk, plonk.public_input(c), proof, msg))
    rs = [1] + [rng.randrange(1, B.R) for _ in inst[1:]]

    def gpu_sum(terms):
        pts = field.g1_to_mont_array([p for p, _ in terms])
        return field.g1_from_mont_array(device.msm_adhoc(ctx, pts, field.fr_to_mont_array([s for _, s in terms])))[0]

    A_terms, B_terms = oplonk.batch_verify_terms(inst, rs)
    A, Bsum = gpu_sum(A_terms), gpu_sum(B_terms)
    assert A == oplonk._msm_terms(A_terms) and Bsum == oplonk._msm_terms(B_terms)
    assert B.g1_mul(A, TAU) == Bsum
    assert oplonk.batch_verify(inst, rs, tau=TAU)
    bad = list(inst)
    vk, pub, proof, msg = bad[4]
    bad[4] = (vk, pub, dict(proof, perm_next_eval=(proof["perm_next_eval"] + 1) % B.R), msg)
    A_terms, B_terms = oplonk.batch_verify_terms(bad, rs)
    assert B.g1_mul(gpu_sum(A_terms), TAU) != gpu_sum(B_terms)
    # edge cases of the ad-hoc entry point: infinity bases, zero scalars, a single term
    P = B.g1_mul(B.G1_GEN, 5)
    pts = field.g1_to_mont_array([P, None, P])
    assert field.g1_from_mont_array(device.msm_adhoc(ctx, pts, field.fr_to_mont_array([3, 9, 0])))[0] == B.g1_mul(P, 3)
    assert field.g1_from_mont_array(device.msm_adhoc(ctx, pts[:1], field.fr_to_mont_array([B.R - 1])))[0] == B.g1_neg(P)
    for srs, pk in keys:
        pk.close()
        srs.close()
