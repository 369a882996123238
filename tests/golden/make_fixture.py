"""Writes tests/fixtures/oracle_*.capfix: replay fixtures in the CAPFIX01 container that
rust/parity-dump writes from inside the reference (layout: oracle/serialize.py).

THESE ARE ORACLE-MADE: key, witness and proof come from oracle/plonk.py, so replaying them pins the
container, the CanonicalSerialize readers and the CUDA prover against the oracle -- NOT against
upstream bytes.  A fixture written by the reference itself goes beside them as
tests/fixtures/upstream_*.capfix and is picked up by tests/test_replay.py with no code change.

    python tests/golden/make_fixture.py
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cap_b200 import synth  # noqa: E402  (workload generator only; no GPU code)
from oracle import bn254 as B  # noqa: E402
from oracle import pairing, plonk as oplonk, serialize as S  # noqa: E402
from oracle.chacha import ChaChaRng  # noqa: E402

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % B.R


def make(log_n: int, nin: int, seed: int, meta, ext_msg: bytes, name: str):
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=seed)
    n = circ.n
    powers = B.srs_powers(TAU, n + 3)
    pk = oplonk.preprocess(circ, tau=TAU)
    h = pairing.G2_GEN
    beta_h = pairing.g2_mul(h, TAU)
    pk_bytes = S.write_proving_key(pk, powers, h, beta_h)
    # the prover's RNG: ark_std::test_rng() (ChaCha with the fixed test seed); every next_u64 it returns
    # while the 17 blinders are drawn is recorded, rejected attempts included
    rng = ChaChaRng()
    words = []
    real_next = rng.next_u64

    def recording_next():
        w = real_next()
        words.append(w)
        return w

    rng.next_u64 = recording_next
    blinders_mont = [rng.fr_rand_mont() for _ in range(17)]
    blinders = [B.from_mont(b, B.R) for b in blinders_mont]
    proof = oplonk.prove(circ, pk, blinders, tau=TAU, ext_msg=ext_msg, keep=True)
    ch = proof.pop("_debug")["challenges"]
    data = S.write_fixture(meta, pk_bytes, oplonk.wire_evals(circ), oplonk.public_input(circ), ext_msg, words, S.write_proof(proof),
                           challenges=[ch[k] for k in ("beta", "gamma", "alpha", "zeta", "v")])
    path = os.path.join(ROOT, "tests", "fixtures", name)
    with open(path, "wb") as f:
        f.write(data)
    print(path, len(data), "bytes;", len(words), "RNG words for 17 draws")


if __name__ == "__main__":
    make(6, 5, 31, (0, 2, 2, 4), b"oracle-fixture-transfer", "oracle_transfer_n64.capfix")
    make(5, 3, 32, (1, 1, 2, 4), b"", "oracle_mint_n32.capfix")
