"""TurboPlonk prover / verifier restatement (oracle; test infrastructure only).

Restates the 5-round prover behind ``PlonkKzgSnark::prove::<_, _, SolidityTranscript>``
(jf-plonk 0.1.2 ``proof_system/{snark,prover}.rs`` and jf-relation 0.1.2
``Arithmetization for PlonkCircuit`` @ jellyfish bcd92b2c; reference call sites
``src/proof/transfer.rs:181``, ``src/proof/mint.rs:113``, ``src/proof/freeze.rs:151``;
preprocess at ``src/proof/transfer.rs:133``) [UPSTREAM-RECALL: SURVEY.md App. A; parity
with upstream bytes is UNPINNED, the mathematics of every step is unique and is what the
CUDA path is checked against].

A circuit is any object with: ``log_n``, ``n``, ``num_inputs``, ``selectors`` (13 lists of
n evaluations, order q_lc[0..4], q_mul[0..2], q_hash[0..4], q_o, q_c, q_ecc),
``wire_variables`` (5 lists of n variable indices), ``witness`` (values per variable),
``k`` (5 coset representatives).  Values are canonical ints.
"""
from __future__ import annotations

from .bn254 import R, FR_GENERATOR, fr_root_of_unity, inv, g1_add, g1_mul, g1_neg, G1_GEN
from .ntt import fft, ifft, coset_fft, coset_ifft, poly_eval
from .msm import kzg_commit_tau, msm_arkworks
from .transcript import SolidityTranscript

NUM_WIRES = 5
NUM_SELECTORS = 13


# ----------------------------------------------------------------------------- helpers
def wire_permutation(circ):
    """jf-relation ``compute_wire_permutation``: cells sharing a variable form a cycle in
    (wire, row) visiting order; returns perm[i*n + j] = (i', j')."""
    n = circ.n
    cells = {}
    for i in range(NUM_WIRES):
        col = circ.wire_variables[i]
        for j in range(n):
            cells.setdefault(int(col[j]), []).append((i, j))
    perm = [None] * (NUM_WIRES * n)
    for lst in cells.values():
        for a, b in zip(lst, lst[1:] + lst[:1]):
            perm[a[0] * n + a[1]] = b
    return perm


def extended_id_permutation(circ):
    n = circ.n
    w = fr_root_of_unity(circ.log_n)
    pw = [1] * n
    for j in range(1, n):
        pw[j] = pw[j - 1] * w % R
    return [circ.k[i] * pw[j] % R for i in range(NUM_WIRES) for j in range(n)]


def sigma_evals(circ):
    """sigma_i(omega^j) = k_{i'} * omega^{j'} with (i', j') = perm(i, j)
    (jf-relation ``compute_extended_permutation_polynomials`` before the ifft)."""
    n = circ.n
    perm = wire_permutation(circ)
    ext = extended_id_permutation(circ)
    return [[ext[perm[i * n + j][0] * n + perm[i * n + j][1]] for j in range(n)] for i in range(NUM_WIRES)]


def wire_evals(circ):
    return [[int(circ.witness[int(v)]) for v in circ.wire_variables[i]] for i in range(NUM_WIRES)]


def public_input(circ):
    return [int(circ.witness[int(circ.wire_variables[4][j])]) for j in range(circ.num_inputs)]


def check_gates(circ) -> bool:
    """Gate identity of cap-specification.pdf section 4.2.1 eq. (1) on every row."""
    w = wire_evals(circ)
    pi = public_input(circ) + [0] * (circ.n - circ.num_inputs)
    s = circ.selectors
    for j in range(circ.n):
        w0, w1, w2, w3, w4 = (w[i][j] for i in range(5))
        v = (s[11][j] + pi[j] + s[0][j] * w0 + s[1][j] * w1 + s[2][j] * w2 + s[3][j] * w3
             + s[4][j] * w0 * w1 + s[5][j] * w2 * w3 + s[12][j] * w0 * w1 * w2 * w3 * w4
             + s[6][j] * pow(w0, 5, R) + s[7][j] * pow(w1, 5, R) + s[8][j] * pow(w2, 5, R)
             + s[9][j] * pow(w3, 5, R) - s[10][j] * w4) % R
        if v:
            return False
    return True


def commit(coeffs, srs=None, tau=None):
    """KZG10::commit (ark-poly-commit 0.3.0): one MSM over powers_of_g, hiding disabled.
    With a synthetic SRS of known tau the same group element is p(tau)*G."""
    if tau is not None:
        return kzg_commit_tau(coeffs, tau)
    return msm_arkworks(srs[:len(coeffs)], [c % R for c in coeffs])


def _poly_add(a, b):
    m = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R for i in range(m)]


def _poly_scale(a, s):
    return [x * s % R for x in a]


def mask_polynomial(coeffs, blinders, n):
    """Prover::mask_polynomial: p(X) + r(X) * (X^n - 1), r = sum blinders[i] X^i."""
    out = list(coeffs) + [0] * (n + len(blinders) - len(coeffs))
    for i, b in enumerate(blinders):
        out[i] = (out[i] - b) % R
        out[n + i] = (out[n + i] + b) % R
    return out


def divide_by_linear(coeffs, point):
    """&poly / (X - point): quotient of the synthetic division (remainder dropped)."""
    d = len(coeffs) - 1
    q = [0] * d
    carry = 0
    for i in range(d, 0, -1):
        carry = (coeffs[i] + carry * point) % R
        q[i - 1] = carry
    return q


# ----------------------------------------------------------------------------- preprocess
def preprocess(circ, srs=None, tau=None):
    """PlonkKzgSnark::preprocess: selector / sigma polynomials (ifft) and their commitments."""
    sel_polys = [ifft(list(s), circ.log_n) for s in circ.selectors]
    sig_ev = sigma_evals(circ)
    sig_polys = [ifft(s, circ.log_n) for s in sig_ev]
    vk = {
        "domain_size": circ.n,
        "num_inputs": circ.num_inputs,
        "k": list(circ.k),
        "selector_comms": [commit(p, srs, tau) for p in sel_polys],
        "sigma_comms": [commit(p, srs, tau) for p in sig_polys],
    }
    return {"selectors": sel_polys, "sigmas": sig_polys, "sigma_evals": sig_ev, "vk": vk}


# ----------------------------------------------------------------------------- prover
def grand_product(circ, beta, gamma, w=None, sig=None):
    """jf-relation ``compute_prod_permutation_polynomial`` evaluations z_0..z_{n-1}.  ``w`` / ``sig``:
    witness columns / sigma evaluations when they come from a recorded proof instead of ``circ``."""
    n = circ.n
    w = wire_evals(circ) if w is None else w
    ext = extended_id_permutation(circ)
    sig = sigma_evals(circ) if sig is None else sig
    z = [1]
    for j in range(n - 1):
        a = 1
        b = 1
        for i in range(NUM_WIRES):
            t = (w[i][j] + gamma) % R
            a = a * (t + beta * ext[i * n + j]) % R
            b = b * (t + beta * sig[i][j]) % R
        z.append(z[-1] * a % R * inv(b, R) % R)
    return z


def quotient_evals(circ, pk, wire_polys, z_poly, pi_poly, beta, gamma, alpha):
    """Prover::compute_quotient_polynomial's point-wise loop over the 8n coset."""
    n = circ.n
    log_m = circ.log_n + 3
    m = 1 << log_m
    ratio = m // n
    sel_c = [coset_fft(p, log_m) for p in pk["selectors"]]
    sig_c = [coset_fft(p, log_m) for p in pk["sigmas"]]
    w_c = [coset_fft(p, log_m) for p in wire_polys]
    z_c = coset_fft(z_poly, log_m)
    pi_c = coset_fft(pi_poly, log_m)
    wm = fr_root_of_unity(log_m)
    zh_inv = [inv((pow(FR_GENERATOR * pow(wm, i, R), n, R) - 1) % R, R) for i in range(ratio)]
    out = []
    x = FR_GENERATOR
    alpha2 = alpha * alpha % R
    for i in range(m):
        w = [w_c[j][i] for j in range(NUM_WIRES)]
        s = [sel_c[j][i] for j in range(NUM_SELECTORS)]
        t_circ = (s[11] + pi_c[i] + s[0] * w[0] + s[1] * w[1] + s[2] * w[2] + s[3] * w[3]
                  + s[4] * w[0] * w[1] + s[5] * w[2] * w[3] + s[12] * w[0] * w[1] * w[2] * w[3] * w[4]
                  + s[6] * pow(w[0], 5, R) + s[7] * pow(w[1], 5, R) + s[8] * pow(w[2], 5, R)
                  + s[9] * pow(w[3], 5, R) - s[10] * w[4]) % R
        r1 = z_c[i]
        r2 = z_c[(i + ratio) % m]
        for j in range(NUM_WIRES):
            r1 = r1 * ((w[j] + circ.k[j] * x % R * beta + gamma) % R) % R
            r2 = r2 * ((w[j] + sig_c[j][i] * beta + gamma) % R) % R
        t_perm1 = alpha * (r1 - r2) % R
        t_perm2 = alpha2 * (z_c[i] - 1) % R * inv(n * (x - 1) % R, R) % R
        out.append(((t_circ + t_perm1) * zh_inv[i % ratio] + t_perm2) % R)
        x = x * wm % R
    return out


def prove(circ, pk, blinders, srs=None, tau=None, ext_msg: bytes | None = None, keep=False):
    """Returns the proof dict (13 G1 + 10 Fr).  ``blinders``: 17 canonical Fr values in the
    order the prover draws them: wires 0..4 (2 each), z (3), split-quotient maskers (4)."""
    return prove_with_columns(circ, pk, wire_evals(circ), public_input(circ), blinders, srs=srs, tau=tau, ext_msg=ext_msg, keep=keep)


def prove_with_columns(circ, pk, w_ev, pub, blinders, srs=None, tau=None, ext_msg: bytes | None = None, keep=False):
    """The prover on explicit witness columns (5 x n values) and public inputs -- what a recorded
    proof (tests/test_replay.py) provides.  ``circ`` only supplies n, log_n and k; the permutation
    enters through ``pk["sigma_evals"]``."""
    assert len(blinders) == 17
    n = circ.n
    log_n = circ.log_n
    omega = fr_root_of_unity(log_n)
    vk = pk["vk"]
    tr = SolidityTranscript()
    if ext_msg is not None:
        tr.append_message(ext_msg)
    tr.append_vk_and_pub_input(vk, pub)

    # Round 1
    wire_polys = [mask_polynomial(ifft(w_ev[i], log_n), blinders[2 * i:2 * i + 2], n) for i in range(NUM_WIRES)]
    wire_comms = [commit(p, srs, tau) for p in wire_polys]
    pi_poly = ifft(pub + [0] * (n - len(pub)), log_n)
    tr.append_commitments(wire_comms)

    # Round 2
    beta = tr.get_and_append_challenge()
    gamma = tr.get_and_append_challenge()
    z_ev = grand_product(circ, beta, gamma, w_ev, pk["sigma_evals"])
    z_poly = mask_polynomial(ifft(z_ev, log_n), blinders[10:13], n)
    z_comm = commit(z_poly, srs, tau)
    tr.append_commitment(z_comm)

    # Round 3
    alpha = tr.get_and_append_challenge()
    t_ev = quotient_evals(circ, pk, wire_polys, z_poly, pi_poly, beta, gamma, alpha)
    t_poly = coset_ifft(t_ev, log_n + 3)
    deg = NUM_WIRES * (n + 1) + 2
    assert all(c == 0 for c in t_poly[deg + 1:]), "quotient degree too large"
    assert t_poly[deg] != 0, "WrongQuotientPolyDegree"
    t_poly = t_poly[:deg + 1]
    split = [t_poly[i * (n + 2):(i + 1) * (n + 2)] for i in range(NUM_WIRES - 1)] + [t_poly[(NUM_WIRES - 1) * (n + 2):]]
    last = 0
    for i in range(NUM_WIRES - 1):
        now = blinders[13 + i]
        split[i][0] = (split[i][0] - last) % R
        split[i].append(now)
        last = now
    split[NUM_WIRES - 1][0] = (split[NUM_WIRES - 1][0] - last) % R
    split_comms = [commit(p, srs, tau) for p in split]
    tr.append_commitments(split_comms)

    # Round 4
    zeta = tr.get_and_append_challenge()
    wires_evals = [poly_eval(p, zeta) for p in wire_polys]
    sigma_evs = [poly_eval(p, zeta) for p in pk["sigmas"][:NUM_WIRES - 1]]
    perm_next_eval = poly_eval(z_poly, zeta * omega % R)
    tr.append_proof_evaluations(wires_evals, sigma_evs, perm_next_eval)

    # linearisation polynomial
    lin = lin_poly(circ.k, n, pk["selectors"], pk["sigmas"][NUM_WIRES - 1], z_poly, split,
                   wires_evals, sigma_evs, perm_next_eval, alpha, beta, gamma, zeta)

    # Round 5
    v = tr.get_and_append_challenge()
    batch = []
    coeff = 1
    for p in [lin] + wire_polys + pk["sigmas"][:NUM_WIRES - 1]:
        batch = _poly_add(batch, _poly_scale(p, coeff))
        coeff = coeff * v % R
    open_poly = divide_by_linear(batch, zeta)
    shifted_poly = divide_by_linear(z_poly, zeta * omega % R)
    opening = commit(open_poly, srs, tau)
    shifted_opening = commit(shifted_poly, srs, tau)

    proof = {
        "wires_poly_comms": wire_comms,
        "prod_perm_poly_comm": z_comm,
        "split_quot_poly_comms": split_comms,
        "opening_proof": opening,
        "shifted_opening_proof": shifted_opening,
        "wires_evals": wires_evals,
        "wire_sigma_evals": sigma_evs,
        "perm_next_eval": perm_next_eval,
    }
    if keep:
        proof["_debug"] = {
            "wire_polys": wire_polys, "pi_poly": pi_poly, "z_evals": z_ev, "z_poly": z_poly,
            "t_evals": t_ev, "t_poly": t_poly, "split": split, "lin": lin,
            "open_poly": open_poly, "shifted_poly": shifted_poly,
            "challenges": {"beta": beta, "gamma": gamma, "alpha": alpha, "zeta": zeta, "v": v},
        }
    return proof


def lin_poly_scalars(k, n, wires_evals, sigma_evs, perm_next_eval, alpha, beta, gamma, zeta):
    """Scalars multiplying (13 selectors, z, sigma_4, 5 split-quotient polys) in the
    linearisation polynomial (Prover::compute_{non_,}quotient_component_for_lin_poly)."""
    w = wires_evals
    zh = (pow(zeta, n, R) - 1) % R
    l1 = zh * inv(n * (zeta - 1) % R, R) % R
    sel = [w[0], w[1], w[2], w[3], w[0] * w[1] % R, w[2] * w[3] % R,
           pow(w[0], 5, R), pow(w[1], 5, R), pow(w[2], 5, R), pow(w[3], 5, R),
           (-w[4]) % R, 1, w[0] * w[1] % R * w[2] % R * w[3] % R * w[4] % R]
    cz = alpha
    for j in range(NUM_WIRES):
        cz = cz * ((w[j] + k[j] * zeta % R * beta + gamma) % R) % R
    cz = (cz + alpha * alpha % R * l1) % R
    cs = alpha * beta % R * perm_next_eval % R
    for j in range(NUM_WIRES - 1):
        cs = cs * ((w[j] + beta * sigma_evs[j] + gamma) % R) % R
    cs = (-cs) % R
    zn2 = (zh + 1) * zeta % R * zeta % R
    ct = []
    c = 1
    for _ in range(NUM_WIRES):
        ct.append((-zh * c) % R)
        c = c * zn2 % R
    return sel, cz, cs, ct, zh, l1


def lin_poly(k, n, sel_polys, sigma_last, z_poly, split, wires_evals, sigma_evs, perm_next_eval,
             alpha, beta, gamma, zeta):
    sel, cz, cs, ct, _, _ = lin_poly_scalars(k, n, wires_evals, sigma_evs, perm_next_eval, alpha, beta, gamma, zeta)
    acc = []
    for p, s in zip(sel_polys, sel):
        acc = _poly_add(acc, _poly_scale(p, s))
    acc = _poly_add(acc, _poly_scale(z_poly, cz))
    acc = _poly_add(acc, _poly_scale(sigma_last, cs))
    for p, s in zip(split, ct):
        acc = _poly_add(acc, _poly_scale(p, s))
    return acc


# ----------------------------------------------------------------------------- verifier
def pcs_terms(vk, pub, proof, ext_msg: bytes | None = None):
    """The verifier's work up to the pairing check, as two multi-scalar sums over commitments:
    returns (A_terms, B_terms), lists of (G1 point, scalar), with the proof valid iff
    e(sum A, [tau]_2) = e(sum B, [1]_2).  (jf-plonk ``Verifier::prepare_pcs_info`` +
    ``aggregate_*``, reached from ``src/proof/transfer.rs:192-212`` and, batched, from
    ``src/lib.rs:517``; CPU-only in the reference.)"""
    n = vk["domain_size"]
    log_n = n.bit_length() - 1
    omega = fr_root_of_unity(log_n)
    tr = SolidityTranscript()
    if ext_msg is not None:
        tr.append_message(ext_msg)
    tr.append_vk_and_pub_input(vk, pub)
    tr.append_commitments(proof["wires_poly_comms"])
    beta = tr.get_and_append_challenge()
    gamma = tr.get_and_append_challenge()
    tr.append_commitment(proof["prod_perm_poly_comm"])
    alpha = tr.get_and_append_challenge()
    tr.append_commitments(proof["split_quot_poly_comms"])
    zeta = tr.get_and_append_challenge()
    w = proof["wires_evals"]
    se = proof["wire_sigma_evals"]
    zw = proof["perm_next_eval"]
    tr.append_proof_evaluations(w, se, zw)
    v = tr.get_and_append_challenge()
    tr.append_commitment(proof["opening_proof"])
    tr.append_commitment(proof["shifted_opening_proof"])
    u = tr.get_and_append_challenge()

    sel, cz, cs, ct, zh, l1 = lin_poly_scalars(vk["k"], n, w, se, zw, alpha, beta, gamma, zeta)
    # PI(zeta) = sum_i pub_i * L_i(zeta)
    pi_eval = 0
    wi = 1
    for x in pub:
        pi_eval = (pi_eval + x * wi % R * zh % R * inv(n * (zeta - wi) % R, R)) % R
        wi = wi * omega % R
    # expected value of the linearisation polynomial at zeta
    tmp = alpha * zw % R
    for j in range(NUM_WIRES - 1):
        tmp = tmp * ((w[j] + beta * se[j] + gamma) % R) % R
    tmp = tmp * ((w[4] + gamma) % R) % R
    lin_eval = (-pi_eval + tmp + alpha * alpha % R * l1) % R

    # [lin] from commitments (coefficient 1 in the batched opening), then wires and sigmas with powers of v
    B_terms = [(c, s) for c, s in zip(vk["selector_comms"], sel)]
    B_terms.append((proof["prod_perm_poly_comm"], (cz + u) % R))  # z also opens at zeta*omega (weight u)
    B_terms.append((vk["sigma_comms"][NUM_WIRES - 1], cs))
    B_terms += [(c, s) for c, s in zip(proof["split_quot_poly_comms"], ct)]
    E = lin_eval
    c = v
    for cm, ev in zip(proof["wires_poly_comms"] + vk["sigma_comms"][:NUM_WIRES - 1], w + se):
        B_terms.append((cm, c))
        E = (E + c * ev) % R
        c = c * v % R
    E = (E + u * zw) % R
    zeta_w = zeta * omega % R
    B_terms.append((proof["opening_proof"], zeta))
    B_terms.append((proof["shifted_opening_proof"], u * zeta_w % R))
    B_terms.append((G1_GEN, (-E) % R))
    A_terms = [(proof["opening_proof"], 1), (proof["shifted_opening_proof"], u)]
    return A_terms, B_terms


def _msm_terms(terms):
    acc = None
    for p, s in terms:
        acc = g1_add(acc, g1_mul(p, s % R))
    return acc


def _check_pairing(A, B, tau, g2_tau) -> bool:
    if g2_tau is not None:
        from .pairing import G2_GEN, pairing_product_is_one
        return pairing_product_is_one([(A, g2_tau), (g1_neg(B), G2_GEN)])
    return g1_mul(A, tau) == B


def verify(vk, pub, proof, tau=None, ext_msg: bytes | None = None, g2_tau=None) -> bool:
    """PLONK verifier restatement (jf-plonk ``PlonkKzgSnark::verify`` reached from
    ``src/proof/transfer.rs:192-212``; out of scope for the GPU and CPU-only in the reference).
    The final equation e(A, [tau]_2) = e(B, [1]_2) is checked either with the BN254 pairing
    (``g2_tau`` = [tau]_2, the way the reference does it; oracle/pairing.py) or, faster, in G1
    as tau*A == B using the synthetic SRS's known ``tau``."""
    A_terms, B_terms = pcs_terms(vk, pub, proof, ext_msg)
    return _check_pairing(_msm_terms(A_terms), _msm_terms(B_terms), tau, g2_tau)


def batch_verify_terms(instances, rs):
    """Random-linear-combination batching of several proofs (``PlonkKzgSnark::batch_verify`` as
    called by ``txn_batch_verify``, ``src/lib.rs:517``): instance i = (vk, pub, proof, ext_msg) is
    weighted by rs[i]; scalars of a base shared between instances (the selector / sigma
    commitments of a common verifying key, the generator) are merged, so the two returned
    multi-scalar sums have ~18 + 13 k and 2 k terms for k proofs of one note type.  The batch is
    valid iff e(sum A, [tau]_2) = e(sum B, [1]_2).  How upstream derives the weights (transcript
    or RNG) does not change the sums' structure; the caller supplies them."""
    A_map, B_map = {}, {}
    for (vk, pub, proof, msg), r in zip(instances, rs):
        A_terms, B_terms = pcs_terms(vk, pub, proof, msg)
        for m, terms in ((A_map, A_terms), (B_map, B_terms)):
            for p, s in terms:
                if p is None:
                    continue
                m[p] = (m.get(p, 0) + r * s) % R
    return list(A_map.items()), list(B_map.items())


def batch_verify(instances, rs, tau=None, g2_tau=None) -> bool:
    A_terms, B_terms = batch_verify_terms(instances, rs)
    return _check_pairing(_msm_terms(A_terms), _msm_terms(B_terms), tau, g2_tau)
