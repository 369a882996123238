//! Raw bindings to `include/capgpu.h`.  Layouts: `Fr`/`Fq` = `[u64; 4]` Montgomery limbs (the
//! in-memory form of ark-ff 0.3 `Fp256`), G1 affine = `[u64; 8]` = x || y, all-zero = infinity.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint, c_void};

#[repr(C)] pub struct capgpu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct capgpu_srs { _p: [u8; 0] }
#[repr(C)] pub struct capgpu_pk { _p: [u8; 0] }
#[repr(C)] pub struct capgpu_job { _p: [u8; 0] }
#[repr(C)] pub struct capgpu_queue { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct capgpu_proof {
    pub wires_poly_comms: [[u64; 8]; 5],
    pub prod_perm_poly_comm: [u64; 8],
    pub split_quot_poly_comms: [[u64; 8]; 5],
    pub opening_proof: [u64; 8],
    pub shifted_opening_proof: [u64; 8],
    pub wires_evals: [[u64; 4]; 5],
    pub wire_sigma_evals: [[u64; 4]; 4],
    pub perm_next_eval: [u64; 4],
}

pub const CAPGPU_OK: c_int = 0;
pub const CAPGPU_ERR_CUDA: c_int = -1;
pub const CAPGPU_ERR_ARG: c_int = -2;
pub const CAPGPU_ERR_DEGREE: c_int = -3;
pub const CAPGPU_ERR_SRS_TOO_SMALL: c_int = -4;
pub const CAPGPU_ERR_STATE: c_int = -5;

extern "C" {
    pub fn capgpu_strerror(code: c_int) -> *const c_char;
    pub fn capgpu_last_error(ctx: *const capgpu_ctx) -> *const c_char;
    pub fn capgpu_ctx_create(device: c_int, out: *mut *mut capgpu_ctx) -> c_int;
    pub fn capgpu_ctx_destroy(ctx: *mut capgpu_ctx);
    pub fn capgpu_ctx_sync(ctx: *mut capgpu_ctx) -> c_int;
    pub fn capgpu_srs_upload(ctx: *mut capgpu_ctx, points_xy: *const u64, n_points: usize, window_bits: c_int, out: *mut *mut capgpu_srs) -> c_int;
    pub fn capgpu_srs_destroy(srs: *mut capgpu_srs);
    pub fn capgpu_msm_g1(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, base_off: usize, scalars: *const u64, n: usize, batch: usize, scalars_mont: c_int, out_xy: *mut u64) -> c_int;
    pub fn capgpu_ntt(ctx: *mut capgpu_ctx, input: *const u64, in_len: usize, out: *mut u64, log_n: c_uint, batch: usize, inverse: c_int, coset: c_int) -> c_int;
    pub fn capgpu_pk_upload(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, log_n: c_uint, num_inputs: usize, selectors: *const u64, sigmas: *const u64, k: *const u64, selector_comms_xy: *const u64, sigma_comms_xy: *const u64, out: *mut *mut capgpu_pk) -> c_int;
    pub fn capgpu_pk_destroy(pk: *mut capgpu_pk);
    pub fn capgpu_prove(ctx: *mut capgpu_ctx, pk: *const capgpu_pk, wires: *const u64, pub_inputs: *const u64, blinders: *const u64, ext_msg: *const u8, ext_msg_len: usize, out: *mut capgpu_proof) -> c_int;
    pub fn capgpu_job_begin(ctx: *mut capgpu_ctx, pk: *const capgpu_pk, wires: *const u64, pub_inputs: *const u64, out: *mut *mut capgpu_job) -> c_int;
    pub fn capgpu_job_round1(job: *mut capgpu_job, blinders10: *const u64, wire_comms_xy: *mut u64) -> c_int;
    pub fn capgpu_job_round2(job: *mut capgpu_job, beta: *const u64, gamma: *const u64, blinders3: *const u64, z_comm_xy: *mut u64) -> c_int;
    pub fn capgpu_job_round3(job: *mut capgpu_job, alpha: *const u64, blinders4: *const u64, split_comms_xy: *mut u64) -> c_int;
    pub fn capgpu_job_round4(job: *mut capgpu_job, zeta: *const u64, evals: *mut u64) -> c_int;
    pub fn capgpu_job_round5(job: *mut capgpu_job, v: *const u64, opening_comms_xy: *mut u64) -> c_int;
    pub fn capgpu_job_end(job: *mut capgpu_job);

    // scheduling, SRS variants, device-resident and batched entry points
    pub fn capgpu_ctx_set_latency_mode(ctx: *mut capgpu_ctx, on: c_int) -> c_int;
    pub fn capgpu_ctx_stream(ctx: *mut capgpu_ctx) -> *mut c_void;
    pub fn capgpu_srs_upload_compressed(ctx: *mut capgpu_ctx, bytes: *const u8, n_points: usize, window_bits: c_int, out: *mut *mut capgpu_srs) -> c_int;
    pub fn capgpu_srs_setup(ctx: *mut capgpu_ctx, tau: *const u64, n_points: usize, window_bits: c_int, out: *mut *mut capgpu_srs) -> c_int;
    pub fn capgpu_srs_export(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, points_xy: *mut u64, n_points: usize) -> c_int;
    pub fn capgpu_srs_size(srs: *const capgpu_srs) -> usize;
    pub fn capgpu_msm_g1_dev(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, base_off: usize, d_scalars: *const c_void, n: usize, batch: usize, scalars_mont: c_int, d_out_xy: *mut c_void) -> c_int;
    pub fn capgpu_msm_g1_dev_part(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, base_off: usize, d_scalars: *const c_void, n: usize, scalars_mont: c_int, part: usize, parts: usize, d_out_xy: *mut c_void) -> c_int;
    pub fn capgpu_msm_g1_adhoc(ctx: *mut capgpu_ctx, points_xy: *const u64, scalars: *const u64, n: usize, scalars_mont: c_int, out_xy: *mut u64) -> c_int;
    pub fn capgpu_g1_sum_dev(ctx: *mut capgpu_ctx, d_points_xy: *const c_void, count: usize, d_out_xy: *mut c_void) -> c_int;
    pub fn capgpu_ntt_dev(ctx: *mut capgpu_ctx, d_in: *const c_void, in_len: usize, d_out: *mut c_void, log_n: c_uint, batch: usize, inverse: c_int, coset: c_int) -> c_int;
    pub fn capgpu_ntt3_dev(ctx: *mut capgpu_ctx, d_in: *const c_void, in_len: usize, d_out: *mut c_void, log_n: c_uint, batch: usize, inverse: c_int) -> c_int;
    pub fn capgpu_preprocess(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, log_n: c_uint, num_inputs: usize, selector_evals: *const u64, sigma_evals: *const u64, k: *const u64, out: *mut *mut capgpu_pk) -> c_int;
    pub fn capgpu_pk_export(ctx: *mut capgpu_ctx, pk: *const capgpu_pk, selectors: *mut u64, sigmas: *mut u64, selector_comms_xy: *mut u64, sigma_comms_xy: *mut u64) -> c_int;
    pub fn capgpu_pk_lagrange(pk: *mut capgpu_pk, enable: c_int) -> c_int;
    pub fn capgpu_pk_lagrange_export(ctx: *mut capgpu_ctx, pk: *const capgpu_pk, points_xy: *mut u64, count: usize) -> c_int;
    pub fn capgpu_prove_dev(ctx: *mut capgpu_ctx, pk: *const capgpu_pk, d_wires: *const c_void, pub_inputs: *const u64, blinders: *const u64, ext_msg: *const u8, ext_msg_len: usize, out: *mut capgpu_proof) -> c_int;
    pub fn capgpu_prove_batch(ctxs: *const *mut capgpu_ctx, n_ctxs: usize, pk: *const capgpu_pk, count: usize, wires: *const *const u64, pub_inputs: *const *const u64, blinders: *const *const u64, ext_msgs: *const *const u8, ext_msg_lens: *const usize, out: *mut capgpu_proof, status: *mut c_int) -> c_int;

    // lockstep groups, device-resident batches, the asynchronous proving queue
    pub fn capgpu_ctx_set_group(ctx: *mut capgpu_ctx, group: c_int) -> c_int;
    pub fn capgpu_pk_info(pk: *const capgpu_pk, log_n: *mut c_uint, num_inputs: *mut usize, k: *mut u64) -> c_int;
    pub fn capgpu_prove_batch_dev(ctxs: *const *mut capgpu_ctx, n_ctxs: usize, pk: *const capgpu_pk, count: usize, d_wires: *const *const c_void, pub_inputs: *const *const u64, blinders: *const *const u64, ext_msgs: *const *const u8, ext_msg_lens: *const usize, out: *mut capgpu_proof, status: *mut c_int) -> c_int;
    pub fn capgpu_queue_create(ctxs: *const *mut capgpu_ctx, n_ctxs: usize, pk: *const capgpu_pk, ring_slots: usize, out: *mut *mut capgpu_queue) -> c_int;
    pub fn capgpu_queue_destroy(q: *mut capgpu_queue);
    pub fn capgpu_submit(q: *mut capgpu_queue, wires: *const u64, pub_inputs: *const u64, blinders: *const u64, ext_msg: *const u8, ext_msg_len: usize, ticket: *mut u64) -> c_int;
    pub fn capgpu_poll(q: *mut capgpu_queue, ticket: u64, done: *mut c_int) -> c_int;
    pub fn capgpu_wait(q: *mut capgpu_queue, ticket: u64, out: *mut capgpu_proof) -> c_int;
    pub fn capgpu_queue_stats(q: *mut capgpu_queue, submitted: *mut u64, completed: *mut u64, groups: *mut u64, copy_ms: *mut f64, wait_slot_ms: *mut f64) -> c_int;

    // the reference's on-disk formats, proof bytes, RNG replay
    pub fn capgpu_sha256(data: *const u8, len: usize, out: *mut u8) -> c_int;
    pub fn capgpu_srs_load_serialized(ctx: *mut capgpu_ctx, bytes: *const u8, len: usize, expect_sha256: *const u8, max_points: usize, window_bits: c_int, out: *mut *mut capgpu_srs) -> c_int;
    pub fn capgpu_pk_load_serialized(ctx: *mut capgpu_ctx, bytes: *const u8, len: usize, consumed: *mut usize, out: *mut *mut capgpu_pk) -> c_int;
    pub fn capgpu_proof_serialize(proof: *const capgpu_proof, out: *mut u8, cap: usize, len: *mut usize) -> c_int;
    pub fn capgpu_fr_rand_from_words(words: *const u64, n_words: usize, out: *mut u64, n_out: usize, used: *mut usize) -> c_int;

    // diagnostics
    pub fn capgpu_debug_read(ctx: *mut capgpu_ctx, what: c_int, out: *mut u64, max_elems: usize, n_elems: *mut usize) -> c_int;
    pub fn capgpu_launch_count(ctx: *const capgpu_ctx) -> u64;
    pub fn capgpu_profile_enable(ctx: *mut capgpu_ctx, on: c_int) -> c_int;
    pub fn capgpu_profile_read(ctx: *const capgpu_ctx, id: c_int, total_ms: *mut f64, launches: *mut u64, units: *mut f64) -> c_int;
    pub fn capgpu_calibrate(ctx: *mut capgpu_ctx, gimad_per_s: *mut f64, gimad_wide_per_s: *mut f64, gfmul_per_s: *mut f64) -> c_int;
    pub fn capgpu_msm_g1_dev_part_xyzz(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, base_off: usize, d_scalars: *const c_void, n: usize, scalars_mont: c_int, part: usize, parts: usize, d_out_xyzz: *mut c_void) -> c_int;
    pub fn capgpu_g1_sum_xyzz_dev(ctx: *mut capgpu_ctx, d_points_xyzz: *const c_void, count: usize, d_out_xy: *mut c_void) -> c_int;
    pub fn capgpu_msm_g1_dev_part_peer(ctx: *mut capgpu_ctx, srs: *const capgpu_srs, base_off: usize, d_scalars: *const c_void, n: usize, scalars_mont: c_int, part: usize, parts: usize, peer_slots: *const *mut c_void, peer_flags: *const *mut c_void, n_peers: usize, epoch: u32) -> c_int;
    pub fn capgpu_g1_sum_xyzz_wait_dev(ctx: *mut capgpu_ctx, d_points_xyzz: *const c_void, d_flags: *const c_void, flag_stride_bytes: usize, count: usize, epoch: u32, d_out_xy: *mut c_void) -> c_int;
    // BLS12-381 (curve = 1) / BLS12-377 (curve = 2) G1 over 12-limb base fields (src/config.rs:86-114)
    pub fn capgpu_curve_msm_g1(ctx: *mut capgpu_ctx, curve: c_int, points_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64) -> c_int;
    pub fn capgpu_curve_fq_op(ctx: *mut capgpu_ctx, curve: c_int, op: c_int, a: *const u64, b: *const u64, out: *mut u64, count: usize) -> c_int;
}

