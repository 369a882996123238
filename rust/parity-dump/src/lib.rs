//! Replay-fixture dumper: proves a circuit with the REFERENCE prover and records everything a
//! replay needs, so that libcapgpu's parity tests (`tests/test_replay.py`, `bench.py --fixture`)
//! can compare bytes with upstream.  NOT COMPILED in the capgpu repository's container (no Rust
//! toolchain, crates not vendored); it uses only public items of the pinned crates
//! [UPSTREAM-RECALL: names per jf-plonk / jf-relation 0.1.2 @ bcd92b2c, ark-* 0.3.0].
//!
//! Container (`CAPFIX01`, little-endian; grammar restated in `oracle/serialize.py`):
//!
//! ```text
//! "CAPFIX01" | u64 n_sections | { 8-byte zero-padded ASCII tag | u64 len | payload }*
//!   META    4 x u64   note type (0 transfer, 1 mint, 2 freeze), n_inputs, n_outputs, tree_depth
//!   PK      ProvingKey::<Bn254>::serialize          (CanonicalSerialize, compressed)
//!   WIRES   Vec<Vec<Fr>>::serialize                 5 columns of n values: witness[wire_variables[i][j]]
//!   PUBIN   Vec<Fr>::serialize                      public inputs
//!   EXTMSG  Vec<u8>::serialize                      extra_transcript_init_msg (empty = None)
//!   RNGU64  Vec<u64>::serialize                     every next_u64 the prover drew, in order
//!   PROOF   Proof::<Bn254>::serialize
//! ```
//!
//! The RNG is recorded as raw words by a wrapping RNG, so the dumper needs no knowledge of which
//! blinders the prover draws or in which order: the replay side re-applies `Fr::rand`'s rejection
//! sampling to the words (`capgpu_fr_rand_from_words`) and learns from the word count whether this
//! revision masks the split quotient (17 draws) or not (13).
use ark_bn254::{Bn254, Fr};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::CanonicalSerialize;
use ark_std::rand::{CryptoRng, Error, RngCore};
use jf_plonk::{
    errors::PlonkError,
    proof_system::{structs::{Proof, ProvingKey}, PlonkKzgSnark, UniversalSNARK},
    transcript::SolidityTranscript,
};
use jf_relation::Arithmetization;
use std::{fs::File, io::Write, path::Path};

/// Passes every request through to the inner RNG and records the 64-bit words it returned.
/// `Fr::rand` (ark-ff 0.3, `impl Distribution<Fp256<P>> for Standard`) samples a `BigInteger256` with
/// four `next_u64` calls, so recording `next_u64` captures the prover's blinders exactly; `next_u32`
/// and `fill_bytes` are recorded too (zero-extended / chunked) so that an unexpected draw pattern
/// shows up on the replay side as a word-count mismatch instead of passing silently.
pub struct RecordingRng<R> {
    pub inner: R,
    pub words: Vec<u64>,
}
impl<R: RngCore> RecordingRng<R> {
    pub fn new(inner: R) -> Self { Self { inner, words: Vec::new() } }
}
impl<R: RngCore> RngCore for RecordingRng<R> {
    fn next_u32(&mut self) -> u32 { let w = self.inner.next_u32(); self.words.push(w as u64); w }
    fn next_u64(&mut self) -> u64 { let w = self.inner.next_u64(); self.words.push(w); w }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        self.inner.fill_bytes(dest);
        for c in dest.chunks(8) { let mut b = [0u8; 8]; b[..c.len()].copy_from_slice(c); self.words.push(u64::from_le_bytes(b)); }
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> { self.fill_bytes(dest); Ok(()) }
}
impl<R: CryptoRng> CryptoRng for RecordingRng<R> {}

fn section(out: &mut Vec<u8>, tag: &str, payload: &[u8]) {
    let mut t = [0u8; 8];
    t[..tag.len()].copy_from_slice(tag.as_bytes());
    out.extend_from_slice(&t);
    out.extend_from_slice(&(payload.len() as u64).to_le_bytes());
    out.extend_from_slice(payload);
}

fn ser<T: CanonicalSerialize>(x: &T) -> Vec<u8> {
    let mut v = Vec::new();
    x.serialize(&mut v).expect("serialize");
    v
}

/// The five witness columns `witness[wire_variables[i][j]]`, recovered through public API only:
/// `compute_wire_polynomials` interpolates them (jf-relation), an FFT over the same domain
/// evaluates them back.
pub fn wire_columns<C: Arithmetization<Fr>>(circuit: &C) -> Result<Vec<Vec<Fr>>, PlonkError> {
    let n = circuit.eval_domain_size()?;
    let domain = Radix2EvaluationDomain::<Fr>::new(n).expect("radix-2 domain");
    Ok(circuit.compute_wire_polynomials()?.iter().map(|p| domain.fft(&p.coeffs)).collect())
}

/// Proves `circuit` under `pk` with upstream's prover and `rng` wrapped in a [`RecordingRng`], then
/// writes the fixture.  `meta` = (note type, n_inputs, n_outputs, tree_depth).  Returns the proof so
/// the calling test can go on to verify it as before.
pub fn dump_fixture<C, R>(path: &Path, meta: [u64; 4], rng: R, circuit: &C, pk: &ProvingKey<Bn254>,
                          ext_msg: Option<Vec<u8>>) -> Result<Proof<Bn254>, PlonkError>
where C: Arithmetization<Fr>, R: RngCore + CryptoRng {
    let mut rec = RecordingRng::new(rng);
    let proof = PlonkKzgSnark::<Bn254>::prove::<_, _, SolidityTranscript>(&mut rec, circuit, pk, ext_msg.clone())?;
    let mut out = Vec::new();
    out.extend_from_slice(b"CAPFIX01");
    out.extend_from_slice(&7u64.to_le_bytes());
    let mut m = Vec::new();
    for x in meta.iter() { m.extend_from_slice(&x.to_le_bytes()); }
    section(&mut out, "META", &m);
    section(&mut out, "PK", &ser(pk));
    section(&mut out, "WIRES", &ser(&wire_columns(circuit)?));
    section(&mut out, "PUBIN", &ser(&circuit.public_input()?));
    section(&mut out, "EXTMSG", &ser(&ext_msg.unwrap_or_default()));
    section(&mut out, "RNGU64", &ser(&rec.words));
    section(&mut out, "PROOF", &ser(&proof));
    File::create(path).and_then(|mut f| f.write_all(&out)).expect("write fixture");
    Ok(proof)
}
