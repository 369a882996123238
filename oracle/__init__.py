"""CPU oracle for the CAP / jf-plonk proving hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, with exact Python integers (and a C restatement under
``oracle/c``), the algorithms that the reference crate ``jf-cap`` reaches
through ``PlonkKzgSnark::prove`` (reference call sites:
``src/proof/transfer.rs:181``, ``src/proof/mint.rs:113``,
``src/proof/freeze.rs:151``).  The arithmetic itself lives in third-party
crates that are NOT vendored under /root/reference (jf-plonk / jf-relation
0.1.2 @ jellyfish bcd92b2c, ark-poly-commit 0.3.0 @ cafc05e3, ark-ec / ark-ff /
ark-poly / ark-bn254 0.3.0 -- Cargo.toml:14-47), so every function cites the
published algorithm it follows and the reference call site it serves.

PARITY UNPINNED: the reference's tests hold no golden vectors for commitments,
evaluations, challenges or proof bytes, and no Rust toolchain is
available to run the reference here.  The oracle is pinned only by
(i) mathematical uniqueness of MSM / NTT / grand-product / quotient results,
(ii) cross-checks between independent algorithms (Pippenger vs double-and-add
vs p(tau)*G; radix-2 NTT vs O(n^2) DFT; Keccak vs hashlib.sha3 permutation,
published Keccak-256 / ChaCha20 known answers), and (iii) the oracle's own
verifier restatement accepting the proofs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import or execute anything in this package, and
only as the checker.  The product path (``cap_b200``) never imports it.
"""
