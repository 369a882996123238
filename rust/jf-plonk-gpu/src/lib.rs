//! `GpuPlonkKzgSnark`: drop-in for `jf_plonk::proof_system::PlonkKzgSnark::<Bn254>::prove` at the
//! three CAP call sites (src/proof/transfer.rs:181, src/proof/mint.rs:113, src/proof/freeze.rs:151).
//! Keys, proofs, verification and serialization stay jf-plonk's own types; only the prover's
//! arithmetic moves to the GPU.  The transcript is upstream's `SolidityTranscript`, driven here on
//! the host through the round-level ABI, so Fiat-Shamir bytes are upstream's by construction.
//!
//! NOT COMPILED in this repository's container (no Rust toolchain, crates not vendored): this is
//! the binding a maintainer adds; field names follow jf-plonk 0.1.2 [UPSTREAM-RECALL].
use ark_bn254::{Bn254, Fr, G1Affine};
use ark_ff::{Field, UniformRand, Zero};
use ark_std::rand::{CryptoRng, RngCore};
use capgpu_sys as sys;
use jf_plonk::{
    errors::PlonkError,
    proof_system::structs::{Proof, ProofEvaluations, ProvingKey},
    transcript::{PlonkTranscript, SolidityTranscript},
};
use jf_relation::Arithmetization;

/// `Fp256` is `#[repr(transparent)]`-like over `BigInteger256([u64; 4])`: a `&[Fr]` is `n x 4` u64.
fn fr_ptr(v: &[Fr]) -> *const u64 { v.as_ptr() as *const u64 }

fn g1_to_xy(p: &G1Affine) -> [u64; 8] {
    let mut o = [0u64; 8];
    if !p.infinity { o[..4].copy_from_slice(&(p.x.0).0); o[4..].copy_from_slice(&(p.y.0).0); }
    o
}
fn xy_to_g1(xy: &[u64; 8]) -> G1Affine {
    if xy.iter().all(|w| *w == 0) { return G1Affine::zero(); }
    let mut x = ark_bn254::Fq::zero(); let mut y = ark_bn254::Fq::zero();
    (x.0).0.copy_from_slice(&xy[..4]); (y.0).0.copy_from_slice(&xy[4..]);
    G1Affine::new(x, y, false)
}
fn check(rc: i32) -> Result<(), PlonkError> {
    if rc == sys::CAPGPU_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::capgpu_strerror(rc)) }.to_string_lossy().into_owned();
    Err(PlonkError::InvalidParameters(format!("capgpu: {msg}")))
}

/// Device-resident proving key: upload once per (note type, n_inputs, n_outputs, tree depth).
pub struct GpuProvingKey { ctx: *mut sys::capgpu_ctx, srs: *mut sys::capgpu_srs, pk: *mut sys::capgpu_pk, n: usize }

impl GpuProvingKey {
    pub fn upload(device: i32, pk: &ProvingKey<Bn254>) -> Result<Self, PlonkError> {
        let n = pk.domain_size();
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::capgpu_ctx_create(device, &mut ctx) })?;
        let bases: Vec<[u64; 8]> = pk.commit_key.powers_of_g.iter().map(g1_to_xy).collect();
        let mut srs = std::ptr::null_mut();
        check(unsafe { sys::capgpu_srs_upload(ctx, bases.as_ptr() as *const u64, bases.len(), 0, &mut srs) })?;
        let pad = |p: &ark_poly::univariate::DensePolynomial<Fr>| { let mut c = p.coeffs.clone(); c.resize(n, Fr::zero()); c };
        let sel: Vec<Fr> = pk.selectors.iter().flat_map(pad).collect();
        let sig: Vec<Fr> = pk.sigmas.iter().flat_map(pad).collect();
        let sc: Vec<[u64; 8]> = pk.vk.selector_comms.iter().map(|c| g1_to_xy(&c.0)).collect();
        let gc: Vec<[u64; 8]> = pk.vk.sigma_comms.iter().map(|c| g1_to_xy(&c.0)).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::capgpu_pk_upload(ctx, srs, n.trailing_zeros(), pk.vk.num_inputs, fr_ptr(&sel), fr_ptr(&sig), fr_ptr(&pk.vk.k),
                                             sc.as_ptr() as *const u64, gc.as_ptr() as *const u64, &mut h) })?;
        Ok(Self { ctx, srs, pk: h, n })
    }
}

pub struct GpuPlonkKzgSnark;

impl GpuPlonkKzgSnark {
    /// Same signature shape as `UniversalSNARK::prove`; `gpk` is the uploaded form of `pk`.
    pub fn prove<C, R>(rng: &mut R, circuit: &C, pk: &ProvingKey<Bn254>, gpk: &GpuProvingKey,
                       extra_transcript_init_msg: Option<Vec<u8>>) -> Result<Proof<Bn254>, PlonkError>
    where C: Arithmetization<Fr>, R: CryptoRng + RngCore {
        let n = gpk.n;
        // witness columns w_i[j] = witness[wire_variables[i][j]] (what compute_wire_polynomials interpolates)
        let wires: Vec<Fr> = circuit.compute_wire_evaluations()?; // 5 * n values, helper added next to compute_wire_polynomials
        let pub_input = circuit.public_input()?;
        let mut tr = <SolidityTranscript as PlonkTranscript<ark_bn254::Fq>>::new(b"PlonkProof");
        if let Some(msg) = extra_transcript_init_msg { tr.append_message(b"extra info", &msg)?; }
        tr.append_vk_and_pub_input(&pk.vk, &pub_input)?;
        let mut job = std::ptr::null_mut();
        check(unsafe { sys::capgpu_job_begin(gpk.ctx, gpk.pk, fr_ptr(&wires), fr_ptr(&pub_input), &mut job) })?;
        // blinders in upstream's draw order: DensePolynomial::rand(1) per wire, rand(2) for z, 4 split maskers
        let b1: Vec<Fr> = (0..10).map(|_| Fr::rand(rng)).collect();
        let mut c1 = [[0u64; 8]; 5];
        check(unsafe { sys::capgpu_job_round1(job, fr_ptr(&b1), c1.as_mut_ptr() as *mut u64) })?;
        let wires_poly_comms: Vec<_> = c1.iter().map(|c| xy_to_g1(c).into()).collect();
        tr.append_commitments(b"witness_poly_comms", &wires_poly_comms)?;
        let beta = tr.get_and_append_challenge::<Bn254>(b"beta")?;
        let gamma = tr.get_and_append_challenge::<Bn254>(b"gamma")?;
        let b2: Vec<Fr> = (0..3).map(|_| Fr::rand(rng)).collect();
        let mut c2 = [0u64; 8];
        check(unsafe { sys::capgpu_job_round2(job, fr_ptr(&[beta]), fr_ptr(&[gamma]), fr_ptr(&b2), c2.as_mut_ptr()) })?;
        let prod_perm_poly_comm = xy_to_g1(&c2).into();
        tr.append_commitment(b"perm_poly_comms", &prod_perm_poly_comm)?;
        let alpha = tr.get_and_append_challenge::<Bn254>(b"alpha")?;
        let b3: Vec<Fr> = (0..4).map(|_| Fr::rand(rng)).collect();
        let mut c3 = [[0u64; 8]; 5];
        check(unsafe { sys::capgpu_job_round3(job, fr_ptr(&[alpha]), fr_ptr(&b3), c3.as_mut_ptr() as *mut u64) })?;
        let split_quot_poly_comms: Vec<_> = c3.iter().map(|c| xy_to_g1(c).into()).collect();
        tr.append_commitments(b"quot_poly_comms", &split_quot_poly_comms)?;
        let zeta = tr.get_and_append_challenge::<Bn254>(b"zeta")?;
        let mut ev = [Fr::zero(); 10];
        check(unsafe { sys::capgpu_job_round4(job, fr_ptr(&[zeta]), ev.as_mut_ptr() as *mut u64) })?;
        let poly_evals = ProofEvaluations { wires_evals: ev[..5].to_vec(), wire_sigma_evals: ev[5..9].to_vec(), perm_next_eval: ev[9] };
        tr.append_proof_evaluations::<Bn254>(&poly_evals)?;
        let v = tr.get_and_append_challenge::<Bn254>(b"v")?;
        let mut c5 = [[0u64; 8]; 2];
        check(unsafe { sys::capgpu_job_round5(job, fr_ptr(&[v]), c5.as_mut_ptr() as *mut u64) })?;
        unsafe { sys::capgpu_job_end(job) };
        let _ = n;
        Ok(Proof { wires_poly_comms, prod_perm_poly_comm, split_quot_poly_comms, opening_proof: xy_to_g1(&c5[0]).into(),
                   shifted_opening_proof: xy_to_g1(&c5[1]).into(), poly_evals, plookup_proof: None })
    }
}
