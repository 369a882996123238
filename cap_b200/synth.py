"""Synthetic TurboPlonk instances of CAP note shape (workload generator for tests / bench).

The prover's cost depends only on (domain size n, 5 wire columns, 13 selectors, number of
public inputs); the reference pins those per note type -- TransferNote 2-in/2-out: n = 2^15
(``src/utils/mod.rs:151-153``), 27 public inputs (``src/proof/transfer.rs:443-458``);
MintNote: 2^14 (``src/utils/mod.rs:163-165``); 3-in/5-out and Freeze 5-in: 2^16
(``src/utils/mod.rs:141-143,184-187``).  This module builds a random *satisfying* circuit of
that shape in the form jf-relation's ``PlonkCircuit`` has after
``finalize_for_arithmetization``: public-input gates first, ~94 % gate utilisation
(30 740 / 32 768 at ``src/proof/transfer.rs:602-603``), zero-padded tail, a variable ->
cells map that induces the copy-constraint permutation.  It does NOT re-implement the
Rescue / Jubjub gadgets of ``src/circuit/*`` (out of scope).
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

from .field import R, COSET_K

NUM_WIRES = 5
NUM_SELECTORS = 13

# note shape -> (log2 domain size, number of public inputs)
NOTE_SHAPES = {
    "mint": (14, 22),
    "transfer_2x2": (15, 27),
    "transfer_3x5": (16, 44),
    "freeze_5": (16, 20),
    "transfer_5x5": (17, 52),
}


@dataclass
class SynthCircuit:
    log_n: int
    num_inputs: int
    selectors: list  # 13 x n canonical ints
    wire_variables: list  # 5 x n variable indices
    witness: list  # canonical ints per variable
    k: tuple = COSET_K
    n: int = field(init=False)

    def __post_init__(self):
        self.n = 1 << self.log_n

    def with_witness(self, seed: int) -> "SynthCircuit":
        """Same circuit shape (selectors, permutation, hence same proving key) with a fresh
        satisfying witness: w0..w3 re-drawn, w4 solved from the gate equation."""
        return _resolve_witness(self, seed)


def _gate_rest(s, j, w0, w1, w2, w3):
    return (s[11][j] + s[0][j] * w0 + s[1][j] * w1 + s[2][j] * w2 + s[3][j] * w3
            + s[4][j] * w0 * w1 + s[5][j] * w2 * w3
            + s[6][j] * pow(w0, 5, R) + s[7][j] * pow(w1, 5, R)
            + s[8][j] * pow(w2, 5, R) + s[9][j] * pow(w3, 5, R)) % R


def make_circuit(log_n: int, num_inputs: int = 27, seed: int = 1, utilization: float = 0.94,
                 reuse: float = 0.35, zero_inputs: float = 0.0, bool_inputs: float = 0.0) -> SynthCircuit:
    """Random satisfying circuit.  Input variables of a gate are drawn from earlier
    variables with probability ``reuse`` (creating copy-constraint cycles), outputs are
    fresh variables whose value solves the gate, so every row satisfies
        q_c + PI + sum q_lc_i w_i + q_mul0 w0 w1 + q_mul1 w2 w3 + q_ecc w0 w1 w2 w3 w4
            + sum q_hash_i w_i^5 - q_o w4 = 0      (cap-specification.pdf 4.2.1 eq. (1)).
    ``zero_inputs``: probability that an input slot is unused (wired to the constant-zero variable,
    as in jf-relation's addition / multiplication / boolean gates, which use 1-2 of the 4 input
    wires); ``bool_inputs``: probability that a fresh input variable is a bit (range-check and
    scalar-decomposition witnesses).  Both default to 0 = dense uniform witness."""
    rng = random.Random(seed)
    n = 1 << log_n
    assert num_inputs < n
    n_gates = max(num_inputs + 1, min(n - 1, int(n * utilization)))
    sel = [[0] * n for _ in range(NUM_SELECTORS)]
    wv = [[0] * n for _ in range(NUM_WIRES)]
    witness = [0, 1]  # jf-relation: variable 0 is the constant zero, variable 1 the constant one
    # public-input gates: q_o = 1, output wire carries the input, PI(omega^j) = value
    for j in range(num_inputs):
        witness.append(rng.randrange(R))
        wv[4][j] = len(witness) - 1
        sel[10][j] = 1
    for j in range(num_inputs, n_gates):
        ins = []
        for i in range(4):
            if zero_inputs and rng.random() < zero_inputs:
                ins.append(0)
            elif rng.random() < reuse:
                ins.append(rng.randrange(len(witness)))
            elif bool_inputs and rng.random() < bool_inputs:
                witness.append(rng.randrange(2))
                ins.append(len(witness) - 1)
            else:
                witness.append(rng.randrange(R))
                ins.append(len(witness) - 1)
            wv[i][j] = ins[-1]
        kind = rng.random()
        if kind < 0.5:  # arithmetic gate: linear combination + multiplications
            for t in (0, 1, 2, 3, 4, 5, 11):
                sel[t][j] = rng.randrange(R)
        elif kind < 0.9:  # rescue-style power-5 gate
            for t in (6, 7, 8, 9, 11):
                sel[t][j] = rng.randrange(R)
        else:  # everything on, including the degree-5 ecc selector
            for t in range(NUM_SELECTORS):
                sel[t][j] = rng.randrange(R)
        sel[10][j] = rng.randrange(1, R)
        w0, w1, w2, w3 = (witness[v] for v in ins)
        rest = _gate_rest(sel, j, w0, w1, w2, w3)
        # rest + q_ecc*w0w1w2w3*w4 - q_o*w4 = 0  ->  w4 = rest / (q_o - q_ecc*w0w1w2w3)
        den = (sel[10][j] - sel[12][j] * w0 * w1 * w2 * w3) % R
        if den == 0:
            sel[12][j] = 0
            den = sel[10][j]
        witness.append(rest * pow(den, -1, R) % R)
        wv[4][j] = len(witness) - 1
    return SynthCircuit(log_n, num_inputs, sel, wv, witness)


def _resolve_witness(c: SynthCircuit, seed: int) -> SynthCircuit:
    rng = random.Random(seed ^ 0x5EED)
    n = c.n
    witness = list(c.witness)
    is_output = [False] * len(witness)
    for j in range(n):
        v = c.wire_variables[4][j]
        if v > 1:
            is_output[v] = True
    for v in range(2, len(witness)):
        if not is_output[v]:
            witness[v] = rng.randrange(2) if witness[v] < 2 else rng.randrange(R)  # bits stay bits
    for j in range(c.num_inputs):
        witness[c.wire_variables[4][j]] = rng.randrange(R)
    s = c.selectors
    for j in range(c.num_inputs, n):
        if s[10][j] == 0:
            continue
        w0, w1, w2, w3 = (witness[c.wire_variables[i][j]] for i in range(4))
        rest = _gate_rest(s, j, w0, w1, w2, w3)
        den = (s[10][j] - s[12][j] * w0 * w1 * w2 * w3) % R
        if den == 0:
            raise ValueError("degenerate gate while re-solving witness; pick another seed")
        witness[c.wire_variables[4][j]] = rest * pow(den, -1, R) % R
    return SynthCircuit(c.log_n, c.num_inputs, c.selectors, c.wire_variables, witness, c.k)
