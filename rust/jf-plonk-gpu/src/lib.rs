//! `GpuPlonkKzgSnark`: drop-in for `jf_plonk::proof_system::PlonkKzgSnark::<Bn254>::prove` at the
//! three CAP call sites (src/proof/transfer.rs:181, src/proof/mint.rs:113, src/proof/freeze.rs:151).
//! Keys, proofs, verification and serialization stay jf-plonk's own types; only the prover's
//! arithmetic moves to the GPU.  The transcript is upstream's `SolidityTranscript`, driven here on
//! the host through the round-level ABI, so Fiat-Shamir bytes are upstream's by construction.
//!
//! NOT COMPILED in this repository's container (no Rust toolchain, crates not vendored).
//! Symbols used, all public in the pinned crates [UPSTREAM-RECALL: jf-plonk / jf-relation 0.1.2 @
//! bcd92b2c]: `ProvingKey::{domain_size, serialize}` and its `pub vk`; `VerifyingKey` via
//! `CanonicalSerialize`; `Arithmetization::{eval_domain_size, compute_wire_polynomials, public_input}`;
//! `PlonkTranscript::{new, append_message, append_vk_and_pub_input, append_commitments,
//! append_commitment, append_proof_evaluations, get_and_append_challenge}`; `Proof`'s and
//! `ProofEvaluations`' public fields.  NOTHING crate-private is touched:
//!   * the key goes to the device as its own `CanonicalSerialize` bytes (`capgpu_pk_load_serialized`),
//!     so `pk.sigmas` / `pk.selectors` / `pk.commit_key` (all `pub(crate)`) are never read;
//!   * the witness columns come from `compute_wire_polynomials` + one forward NTT on the device.
//! One OPTIONAL fork point, behind the cargo feature `wire-evaluations`: a six-line
//! `PlonkCircuit::compute_wire_evaluations()` in jf-relation (the first half of
//! `compute_wire_polynomials`, before its `ifft`) saves the CPU the five inverse FFTs per note.
use ark_bn254::{Bn254, Fq, Fr, G1Affine};
use ark_ff::{UniformRand, Zero};
use ark_serialize::CanonicalSerialize;
use ark_std::rand::{CryptoRng, RngCore};
use capgpu_sys as sys;
use jf_plonk::{
    errors::PlonkError,
    proof_system::structs::{Proof, ProofEvaluations, ProvingKey},
    transcript::{PlonkTranscript, SolidityTranscript},
};
use jf_relation::Arithmetization;

/// `Fp256<P>(BigInteger256([u64; 4]), PhantomData)`: a `&[Fr]` is `n x 4` u64 in Montgomery form.
fn fr_ptr(v: &[Fr]) -> *const u64 { v.as_ptr() as *const u64 }

fn xy_to_g1(xy: &[u64; 8]) -> G1Affine {
    if xy.iter().all(|w| *w == 0) { return G1Affine::zero(); }
    let (mut x, mut y) = (Fq::zero(), Fq::zero());
    (x.0).0.copy_from_slice(&xy[..4]);
    (y.0).0.copy_from_slice(&xy[4..]);
    G1Affine::new(x, y, false)
}

fn check(rc: i32) -> Result<(), PlonkError> {
    if rc == sys::CAPGPU_OK { return Ok(()); }
    if rc == sys::CAPGPU_ERR_DEGREE { return Err(PlonkError::WrongQuotientPolyDegree(0, 0)); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::capgpu_strerror(rc)) }.to_string_lossy().into_owned();
    Err(PlonkError::InvalidParameters(format!("capgpu: {msg}")))
}

/// Device-resident proving key and its context: upload once per (note type, n_inputs, n_outputs,
/// tree depth) -- the analogue of the reference's on-disk key cache (src/parameters.rs:485-503).
pub struct GpuProvingKey { ctx: *mut sys::capgpu_ctx, pk: *mut sys::capgpu_pk, log_n: u32, n: usize }

// the handles are only used behind &mut self / one job at a time
unsafe impl Send for GpuProvingKey {}

impl GpuProvingKey {
    pub fn upload(device: i32, pk: &ProvingKey<Bn254>) -> Result<Self, PlonkError> {
        let mut bytes = Vec::new();
        pk.serialize(&mut bytes).map_err(|e| PlonkError::InvalidParameters(format!("{e:?}")))?;
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::capgpu_ctx_create(device, &mut ctx) })?;
        let mut this = Self { ctx, pk: std::ptr::null_mut(), log_n: 0, n: 0 }; // Drop releases ctx on every error path below
        check(unsafe { sys::capgpu_pk_load_serialized(ctx, bytes.as_ptr(), bytes.len(), std::ptr::null_mut(), &mut this.pk) })?;
        let mut num_inputs = 0usize;
        check(unsafe { sys::capgpu_pk_info(this.pk, &mut this.log_n, &mut num_inputs, std::ptr::null_mut()) })?;
        this.n = 1usize << this.log_n;
        Ok(this)
    }
}

impl Drop for GpuProvingKey {
    fn drop(&mut self) {
        unsafe {
            if !self.pk.is_null() { sys::capgpu_pk_destroy(self.pk); }   // also frees the embedded commit key
            if !self.ctx.is_null() { sys::capgpu_ctx_destroy(self.ctx); }
        }
    }
}

/// RAII guard: `capgpu_job_end` runs on every exit path, including `?` after a failed round.
struct Job(*mut sys::capgpu_job);
impl Drop for Job {
    fn drop(&mut self) { unsafe { sys::capgpu_job_end(self.0) } }
}

pub struct GpuPlonkKzgSnark;

impl GpuPlonkKzgSnark {
    /// Same signature shape as `UniversalSNARK::prove`; `gpk` is the uploaded form of `pk`.
    pub fn prove<C, R>(rng: &mut R, circuit: &C, pk: &ProvingKey<Bn254>, gpk: &mut GpuProvingKey,
                       extra_transcript_init_msg: Option<Vec<u8>>) -> Result<Proof<Bn254>, PlonkError>
    where C: Arithmetization<Fr>, R: CryptoRng + RngCore {
        let n = gpk.n;
        if circuit.eval_domain_size()? != n {
            return Err(PlonkError::InvalidParameters("circuit / proving key domain size mismatch".into()));
        }
        // witness columns w_i[j] = witness[wire_variables[i][j]]
        #[cfg(feature = "wire-evaluations")]
        let wires: Vec<Fr> = circuit.compute_wire_evaluations()?.into_iter().flatten().collect();
        #[cfg(not(feature = "wire-evaluations"))]
        let wires: Vec<Fr> = {
            // public API only: the coefficient polynomials, evaluated back on the device
            let mut coeffs: Vec<Fr> = Vec::with_capacity(5 * n);
            for p in circuit.compute_wire_polynomials()? {
                coeffs.extend_from_slice(&p.coeffs);
                coeffs.resize(coeffs.len() + n - p.coeffs.len(), Fr::zero());
            }
            let mut evals = vec![Fr::zero(); 5 * n];
            check(unsafe { sys::capgpu_ntt(gpk.ctx, fr_ptr(&coeffs), n, evals.as_mut_ptr() as *mut u64, gpk.log_n, 5, 0, 0) })?;
            evals
        };
        let pub_input = circuit.public_input()?;
        let mut tr = <SolidityTranscript as PlonkTranscript<Fq>>::new(b"PlonkProof");
        if let Some(msg) = extra_transcript_init_msg { tr.append_message(b"extra info", &msg)?; }
        tr.append_vk_and_pub_input(&pk.vk, &pub_input)?;
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::capgpu_job_begin(gpk.ctx, gpk.pk, fr_ptr(&wires), fr_ptr(&pub_input), &mut raw) })?;
        let job = Job(raw);
        // blinders in upstream's draw order: DensePolynomial::rand(1) per wire, rand(2) for z, 4 split maskers
        let b1: Vec<Fr> = (0..10).map(|_| Fr::rand(rng)).collect();
        let mut c1 = [[0u64; 8]; 5];
        check(unsafe { sys::capgpu_job_round1(job.0, fr_ptr(&b1), c1.as_mut_ptr() as *mut u64) })?;
        let wires_poly_comms: Vec<_> = c1.iter().map(|c| xy_to_g1(c).into()).collect();
        tr.append_commitments(b"witness_poly_comms", &wires_poly_comms)?;
        let beta = tr.get_and_append_challenge::<Bn254>(b"beta")?;
        let gamma = tr.get_and_append_challenge::<Bn254>(b"gamma")?;
        let b2: Vec<Fr> = (0..3).map(|_| Fr::rand(rng)).collect();
        let mut c2 = [0u64; 8];
        check(unsafe { sys::capgpu_job_round2(job.0, fr_ptr(&[beta]), fr_ptr(&[gamma]), fr_ptr(&b2), c2.as_mut_ptr()) })?;
        let prod_perm_poly_comm = xy_to_g1(&c2).into();
        tr.append_commitment(b"perm_poly_comms", &prod_perm_poly_comm)?;
        let alpha = tr.get_and_append_challenge::<Bn254>(b"alpha")?;
        let b3: Vec<Fr> = (0..4).map(|_| Fr::rand(rng)).collect();
        let mut c3 = [[0u64; 8]; 5];
        check(unsafe { sys::capgpu_job_round3(job.0, fr_ptr(&[alpha]), fr_ptr(&b3), c3.as_mut_ptr() as *mut u64) })?;
        let split_quot_poly_comms: Vec<_> = c3.iter().map(|c| xy_to_g1(c).into()).collect();
        tr.append_commitments(b"quot_poly_comms", &split_quot_poly_comms)?;
        let zeta = tr.get_and_append_challenge::<Bn254>(b"zeta")?;
        let mut ev = [Fr::zero(); 10];
        check(unsafe { sys::capgpu_job_round4(job.0, fr_ptr(&[zeta]), ev.as_mut_ptr() as *mut u64) })?;
        let poly_evals = ProofEvaluations { wires_evals: ev[..5].to_vec(), wire_sigma_evals: ev[5..9].to_vec(), perm_next_eval: ev[9] };
        tr.append_proof_evaluations::<Bn254>(&poly_evals)?;
        let v = tr.get_and_append_challenge::<Bn254>(b"v")?;
        let mut c5 = [[0u64; 8]; 2];
        check(unsafe { sys::capgpu_job_round5(job.0, fr_ptr(&[v]), c5.as_mut_ptr() as *mut u64) })?;
        drop(job);
        Ok(Proof { wires_poly_comms, prod_perm_poly_comm, split_quot_poly_comms, opening_proof: xy_to_g1(&c5[0]).into(),
                   shifted_opening_proof: xy_to_g1(&c5[1]).into(), poly_evals, plookup_proof: None })
    }
}
