/* capgpu.h -- C ABI of the B200-native PLONK proving backend for CAP (jf-cap 0.0.4).
 *
 * This is the drop-in boundary: the entry points a Rust `capgpu-sys` crate binds (see
 * INTEGRATION.md) in place of the CPU routines that `PlonkKzgSnark::prove` reaches from
 * /root/reference/src/proof/transfer.rs:181, src/proof/mint.rs:113 and
 * src/proof/freeze.rs:151.  Plain pointers and sizes only; no C++/torch types.
 *
 * Data layout (identical to ark-ff 0.3.0 / ark-ec 0.3.0 in-memory values, so Rust slices
 * cross without conversion):
 *   Fr / Fq element : 4 x uint64_t little-endian limbs, MONTGOMERY form (a * 2^256 mod p)
 *   &[Fr] of len N  : contiguous N x 4 uint64_t
 *   G1 affine       : 8 x uint64_t = x[4] || y[4] (Fq, Montgomery); all-zero == infinity
 *                     (Rust `GroupAffine{x,y,infinity}` is repacked to this by the shim)
 * All host buffers are owned by the caller; the library copies in/out.  Handles are opaque
 * and freed by their *_destroy call.  Every function returns 0 on success or a negative
 * CAPGPU_ERR_* code; `capgpu_strerror` maps it to text; the shim maps it to
 * `PlonkError` -> `TxnApiError::FailedSnark` (src/proof/transfer.rs:187).
 * One `capgpu_ctx` per (host thread, GPU): a ctx owns one CUDA stream and its workspace and
 * must not be used from two threads at once; `srs` / `pk` handles are immutable after
 * creation and may be shared by all ctxs of the same device.  There is no CPU fallback:
 * without a CUDA device `capgpu_ctx_create` fails with CAPGPU_ERR_CUDA.
 */
#ifndef CAPGPU_H
#define CAPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPGPU_OK 0
#define CAPGPU_ERR_CUDA (-1)        /* CUDA runtime error (sticky per ctx) */
#define CAPGPU_ERR_ARG (-2)         /* invalid argument / size */
#define CAPGPU_ERR_DEGREE (-3)      /* quotient polynomial has the wrong degree (jf-plonk WrongQuotientPolyDegree) */
#define CAPGPU_ERR_SRS_TOO_SMALL (-4) /* commit key shorter than the polynomial (jf-plonk / KZG10 TooManyCoefficients) */
#define CAPGPU_ERR_STATE (-5)       /* round API called out of order */

#define CAPGPU_NUM_WIRES 5          /* TurboPlonk: /root/reference/src/circuit/transfer.rs:68 */
#define CAPGPU_NUM_SELECTORS 13     /* q_lc[4], q_mul[2], q_hash[4], q_o, q_c, q_ecc */
#define CAPGPU_NUM_BLINDERS 17      /* wires 2x5, permutation product 3, split quotient 4 */

typedef struct capgpu_ctx capgpu_ctx;
typedef struct capgpu_srs capgpu_srs;
typedef struct capgpu_pk capgpu_pk;
typedef struct capgpu_job capgpu_job;

const char* capgpu_strerror(int code);
/* Text of the last CUDA error seen by this ctx ("" if none). */
const char* capgpu_last_error(const capgpu_ctx* ctx);

/* ---- context ------------------------------------------------------------------------- */
int capgpu_ctx_create(int device, capgpu_ctx** out);
void capgpu_ctx_destroy(capgpu_ctx* ctx);
int capgpu_ctx_sync(capgpu_ctx* ctx);
/* Scheduling preference of the prover's internal MSMs: 0 (default) minimises total GPU work per
 * proof (best proofs/s with several contexts per GPU), 1 minimises the depth of each MSM's
 * bucket reduction (best single-proof latency).  Results are identical either way.  The
 * standalone capgpu_msm_g1 / capgpu_msm_g1_dev calls always use the low-latency schedule. */
int capgpu_ctx_set_latency_mode(capgpu_ctx* ctx, int on);
/* Number of notes a context proves in LOCKSTEP when it is handed a batch (capgpu_prove_batch, the
 * proving queue): every round is issued once for the whole group, so the round's NTT / MSM launches
 * carry 5-7 vectors per note and one host thread drives `group` proofs.  1 <= group <= 64, default 8
 * (CAPGPU_GROUP in the environment overrides the default).  Workspace: ~150 MB per note at n = 2^15. */
int capgpu_ctx_set_group(capgpu_ctx* ctx, int group);
/* Raw cudaStream_t of the ctx (for callers that time with CUDA events). */
void* capgpu_ctx_stream(capgpu_ctx* ctx);

/* ---- SRS / commit key ------------------------------------------------------------------
 * Replaces the `powers_of_g` vector of ark-poly-commit 0.3.0 `Powers` / `UniversalParams`
 * (built at /root/reference/src/proof/mod.rs:59-69 or loaded at :74-109).  Uploading also
 * precomputes the window-shifted copies 2^(c*w) * P_i used by the MSM. `window_bits` = 0
 * lets the library choose.  */
int capgpu_srs_upload(capgpu_ctx* ctx, const uint64_t* points_xy, size_t n_points, int window_bits, capgpu_srs** out);
/* Same from ark-serialize 0.3 *compressed* points (32 bytes each: x little-endian, bit 255 = y is
 * the larger root, bit 254 = infinity), the encoding of the `powers_of_g` vector inside the SRS /
 * proving-key files of /root/reference/src/parameters.rs:557-592 and the Aztec CRS that
 * src/proof/mod.rs:74-109 deserialises; decompression (square root in Fq) runs on the device. */
int capgpu_srs_upload_compressed(capgpu_ctx* ctx, const uint8_t* bytes, size_t n_points, int window_bits, capgpu_srs** out);
/* `KZG10::setup` on the device (PlonkKzgSnark::universal_setup, src/proof/mod.rs:59-69, minus
 * the RNG): powers_of_g[i] = tau^i * g for the caller's tau (Fr, Montgomery) and g = (1, 2).
 * Used for synthetic SRS in tests / benches (the Aztec CRS blob is not shipped). */
int capgpu_srs_setup(capgpu_ctx* ctx, const uint64_t* tau, size_t n_points, int window_bits, capgpu_srs** out);
/* Copies the first n_points bases (affine, ABI layout) back to the host. */
int capgpu_srs_export(capgpu_ctx* ctx, const capgpu_srs* srs, uint64_t* points_xy, size_t n_points);
void capgpu_srs_destroy(capgpu_srs* srs);
size_t capgpu_srs_size(const capgpu_srs* srs);

/* ---- G1 multi-scalar multiplication ------------------------------------------------------
 * Replaces ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul(&powers_of_g[base_off..], scalars)`
 * as called by `KZG10::commit`.  `batch` independent scalar vectors of `n` elements each
 * (stride n) over the same bases; `scalars_mont` != 0: scalars are Fr Montgomery values
 * (what a `DensePolynomial` holds), 0: canonical `BigInteger256` (what `into_repr` gives).
 * out_xy: batch x 8 uint64_t affine results.  */
int capgpu_msm_g1(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const uint64_t* scalars, size_t n,
                  size_t batch, int scalars_mont, uint64_t* out_xy);

/* Same, with `scalars` and `out_xy` already resident in device memory of ctx's GPU; enqueued on
 * the ctx stream without synchronising (pair with capgpu_ctx_sync). */
int capgpu_msm_g1_dev(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                      size_t batch, int scalars_mont, void* d_out_xy);

/* One BUCKET-RANGE slice of an MSM, for splitting a single large MSM across the GPUs of a box
 * (BASELINE north_star: "a single large MSM can be split ... with partial sums reduced over NVLink").
 * Every GPU holds the whole commit key and the whole scalar vector; GPU `part` of `parts` keeps only
 * the signed digits whose magnitude falls in its 1/parts of the 2^(c-1) buckets, so its sort,
 * bucket accumulation AND bucket reduction all shrink by `parts`; the slice results are affine points
 * whose sum (capgpu_g1_sum_dev after a gather over NVLink) is the MSM.  parts must divide 2^(c-1)
 * (any power of two up to 8 does for every key of 16 points or more).  Asynchronous on the ctx stream. */
int capgpu_msm_g1_dev_part(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                           int scalars_mont, size_t part, size_t parts, void* d_out_xy);

/* MSM over bases that are not a resident SRS: sum_i scalars[i] * points[i] for n affine points
 * given by the caller (x || y Montgomery, all-zero = infinity).  This is the G1 work of the
 * batched verifier (`txn_batch_verify`, /root/reference/src/lib.rs:517, benches/batch_verification.rs:
 * two aggregated sums over the proofs' and verifying keys' commitments); the pairings stay on the
 * CPU.  Tables are built for the call and released. */
int capgpu_msm_g1_adhoc(capgpu_ctx* ctx, const uint64_t* points_xy, const uint64_t* scalars, size_t n,
                        int scalars_mont, uint64_t* out_xy);

/* Sum of `count` affine points resident on the device (asynchronous on the ctx stream).  Used to
 * fold a point-range-split MSM: each GPU runs capgpu_msm_g1_dev over its slice of the bases, the
 * 64-byte partial results are gathered over NVLink (NCCL all-gather) and added here. */
int capgpu_g1_sum_dev(capgpu_ctx* ctx, const void* d_points_xy, size_t count, void* d_out_xy);
/* Same split with the slice result left in extended Jacobian XYZZ form (X, Y, ZZ, ZZZ: 4 x 4 u64 Montgomery,
 * x = X / ZZ, y = Y / ZZZ, ZZ = 0 for infinity) and its fold: the slices skip their own conversion to affine
 * (one field inversion each) and the fold of the gathered 128-byte results converts once.  Needs at least
 * 512 buckets per slice (windows of >= 10 + log2(parts) bits: every SRS of >= 2^12 points). */
int capgpu_msm_g1_dev_part_xyzz(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                int scalars_mont, size_t part, size_t parts, void* d_out_xyzz);
int capgpu_g1_sum_xyzz_dev(capgpu_ctx* ctx, const void* d_points_xyzz, size_t count, void* d_out_xy);
/* The same split with the exchange FUSED into the kernels, over peer-mapped memory (NVLink / NVSwitch), no
 * collective call on the data path.  Every GPU owns a buffer of `parts` 128-byte slots and `parts` 32-bit flags that
 * its peers can address (CUDA peer access / IPC / symmetric memory; the host side obtains the pointers, e.g. from
 * torch.distributed._symmetric_memory).  capgpu_msm_g1_dev_part_peer computes slice `part` and its last kernel stores
 * the XYZZ sum into peer_slots[i] (slot `part` of GPU i, i < n_peers <= 16; the GPU's own buffer included) and then
 * releases peer_flags[i] = epoch with system scope.  capgpu_g1_sum_xyzz_wait_dev (on each GPU that wants the result)
 * waits until all `count` flags of ITS buffer hold `epoch` (acquire, system scope; it gives up after about two
 * seconds, leaving a wrong result rather than a hung GPU), then folds the slots and converts to affine once.
 * Use a fresh epoch per MSM and alternate between two buffers, or synchronise the GPUs between MSMs: a peer that
 * runs ahead would overwrite slots still being folded. */
int capgpu_msm_g1_dev_part_peer(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                int scalars_mont, size_t part, size_t parts, void* const* peer_slots, void* const* peer_flags,
                                size_t n_peers, uint32_t epoch);
int capgpu_g1_sum_xyzz_wait_dev(capgpu_ctx* ctx, const void* d_points_xyzz, const void* d_flags, size_t flag_stride_bytes,
                                size_t count, uint32_t epoch, void* d_out_xy);

/* ---- radix-2 NTT over Fr -----------------------------------------------------------------
 * Replaces ark-poly 0.3.0 `Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}`
 * (natural order in and out, coset shift = Fr::multiplicative_generator() = 5, ifft includes
 * the 1/n scaling).  `in` holds `batch` vectors of `in_len` <= 2^log_n elements (zero-padded
 * like arkworks), `out` receives batch x 2^log_n elements; in == out is allowed when
 * in_len == 2^log_n.  */
int capgpu_ntt(capgpu_ctx* ctx, const uint64_t* in, size_t in_len, uint64_t* out, unsigned log_n, size_t batch,
               int inverse, int coset);

/* Device-resident variant (d_in != d_out unless in_len == 2^log_n); asynchronous on the ctx stream. */
int capgpu_ntt_dev(capgpu_ctx* ctx, const void* d_in, size_t in_len, void* d_out, unsigned log_n, size_t batch,
                   int inverse, int coset);

/* The 3 * 2^log_n-point domain g <rho> (rho = 5^((r-1) / (3 * 2^log_n))), stored as its three cosets g rho^k H, k-major: value
 * (k, i) = f(g rho^k w^i) sits at index k * 2^log_n + i.  The prover evaluates the quotient (degree 5n + 7) on 6n points this
 * way (log_n = log2(2n)) where jf-plonk's compute_quotient_polynomial uses ark-poly's 8n-point coset; both interpolate the same
 * polynomial.  Forward (inverse == 0): d_in holds `batch` vectors of in_len <= 2^log_n coefficients, d_out receives
 * batch x 3 * 2^log_n values (out of place).  Inverse: in_len == 3 * 2^log_n values per vector -> the 3 * 2^log_n coefficients
 * (d_in == d_out allowed).  Device pointers; asynchronous on the ctx stream. */
int capgpu_ntt3_dev(capgpu_ctx* ctx, const void* d_in, size_t in_len, void* d_out, unsigned log_n, size_t batch, int inverse);

/* ---- proving key --------------------------------------------------------------------------
 * Mirrors jf-plonk 0.1.2 `ProvingKey { sigmas, selectors, commit_key, vk }` as embedded at
 * /root/reference/src/proof/transfer.rs:60 (built by `preprocess`, :124-155).
 * selectors: 13 x n coefficients, sigmas: 5 x n coefficients (each padded to n), k: 5 coset
 * representatives, selector_comms / sigma_comms: the vk commitments (13 + 5 affine points).
 * The upload caches coset-domain evaluations of all 18 polynomials on the device.  */
int capgpu_pk_upload(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs,
                     const uint64_t* selectors, const uint64_t* sigmas, const uint64_t* k,
                     const uint64_t* selector_comms_xy, const uint64_t* sigma_comms_xy, capgpu_pk** out);
/* `PlonkKzgSnark::preprocess` on the device (src/proof/transfer.rs:133): takes selector and
 * sigma EVALUATIONS over the domain (13 x n, 5 x n), interpolates, commits, builds the pk.
 * The computed coefficient polynomials / commitments can be read back with capgpu_pk_export. */
int capgpu_preprocess(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs,
                      const uint64_t* selector_evals, const uint64_t* sigma_evals, const uint64_t* k, capgpu_pk** out);
int capgpu_pk_export(capgpu_ctx* ctx, const capgpu_pk* pk, uint64_t* selectors, uint64_t* sigmas,
                     uint64_t* selector_comms_xy, uint64_t* sigma_comms_xy);
void capgpu_pk_destroy(capgpu_pk* pk);
/* Shape of a key: log2 of the domain size, number of public inputs, the 5 coset representatives k
 * (5 x 4 uint64_t, Montgomery).  Any output pointer may be NULL. */
int capgpu_pk_info(const capgpu_pk* pk, unsigned* log_n, size_t* num_inputs, uint64_t* k);
/* Evaluation-form ("Lagrange basis") wire commitments.  Building a proving key also
 * derives, once, the Lagrange commit key L_j = L_j(tau)·G (inverse DFT of the first n SRS points in
 * the group); round 1 then commits the five wire columns straight from their evaluations
 * `witness[wire_variables[i][j]]` — the same commitments `KZG10::commit` returns for the masked
 * coefficient polynomials (src/proof/transfer.rs:181), but zero witness cells are skipped and
 * boolean / small ones cost one bucket addition instead of 16.
 * capgpu_pk_lagrange: enable != 0 turns the evaluation-form path on (CAPGPU_ERR_STATE if the key
 * was built without the Lagrange commit key, i.e. with CAPGPU_LAGRANGE=0 in the environment),
 * 0 commits from coefficients.  capgpu_pk_lagrange_export reads back the first `count` <= n + 4
 * bases [L_0 .. L_{n-1}, P_0, P_1, P_n, P_{n+1}]. */
int capgpu_pk_lagrange(capgpu_pk* pk, int enable);
int capgpu_pk_lagrange_export(capgpu_ctx* ctx, const capgpu_pk* pk, uint64_t* points_xy, size_t count);

/* ---- proof output --------------------------------------------------------------------------
 * Mirrors jf-plonk `Proof` (embedded at /root/reference/src/transfer.rs:60): 13 G1 + 10 Fr. */
typedef struct capgpu_proof {
  uint64_t wires_poly_comms[5][8];
  uint64_t prod_perm_poly_comm[8];
  uint64_t split_quot_poly_comms[5][8];
  uint64_t opening_proof[8];
  uint64_t shifted_opening_proof[8];
  uint64_t wires_evals[5][4];
  uint64_t wire_sigma_evals[4][4];
  uint64_t perm_next_eval[4];
} capgpu_proof;

/* ---- the reference's on-disk formats ----------------------------------------------
 * `UniversalSrs` and the note proving keys are stored as ark-serialize 0.3 `CanonicalSerialize`
 * blobs (`store_data` / `load_data`, /root/reference/src/parameters.rs:557-592); the Aztec CRS is
 * deserialised behind a SHA-256 gate (/root/reference/src/proof/mod.rs:98-107).  These entry points
 * take the file bytes as they are; the byte grammar is documented at the top of
 * cap_b200/csrc/formats.cu and in INTEGRATION.md.  A blob that does not parse exactly is refused
 * with CAPGPU_ERR_ARG (capgpu_last_error names the field).
 *
 * capgpu_srs_load_serialized: `UniversalSrs::deserialize(bytes)` + `trim`: expect_sha256 (32 bytes or
 *   NULL) is checked first, like the reference's `assert_eq!(hasher.finalize(), hex!(..))`; only the
 *   first max_points powers are kept (0 = all); points are decompressed on the device.
 * capgpu_pk_load_serialized: `ProvingKey::deserialize(bytes)`: builds the key AND its commit key (the
 *   `powers_of_g` embedded in the blob; owned by the key).  `consumed` (optional) receives the number
 *   of bytes the key took -- CAP's `TransferProvingKey` / `MintProvingKey` / `FreezeProvingKey`
 *   (src/proof/transfer.rs:60, mint.rs:55, freeze.rs:47) append n_inputs / n_outputs / tree_depth after
 *   it; with consumed == NULL trailing bytes are an error.
 * capgpu_proof_serialize: the `CanonicalSerialize` bytes of jf-plonk's `Proof` for a capgpu_proof
 *   (out == NULL: size query). */
int capgpu_sha256(const uint8_t* data, size_t len, uint8_t out[32]);
int capgpu_srs_load_serialized(capgpu_ctx* ctx, const uint8_t* bytes, size_t len, const uint8_t* expect_sha256, size_t max_points,
                               int window_bits, capgpu_srs** out);
int capgpu_pk_load_serialized(capgpu_ctx* ctx, const uint8_t* bytes, size_t len, size_t* consumed, capgpu_pk** out);
int capgpu_proof_serialize(const capgpu_proof* proof, uint8_t* out, size_t cap, size_t* len);
/* The blinding scalars a prover draws, from the raw `next_u64` words of its RNG in draw order, the
 * way ark-ff 0.3 `Fr::rand` consumes them (4 words per attempt, top two bits cleared, rejected if
 * >= r; the accepted limbs are the Montgomery representation).  Used to replay a recorded proof
 * (rust/parity-dump): out receives n_out x 4 uint64_t, *used the number of words consumed;
 * CAPGPU_ERR_ARG if the words run out. */
int capgpu_fr_rand_from_words(const uint64_t* words, size_t n_words, uint64_t* out, size_t n_out, size_t* used);

/* ---- whole proof -----------------------------------------------------------------------------
 * Replaces `PlonkKzgSnark::prove::<_, _, SolidityTranscript>(rng, &circuit, &pk, Some(ext_msg))`.
 * wires: 5 x n witness values per wire column (`witness[wire_variables[i][j]]`, Montgomery);
 * pub_inputs: the `num_inputs` public inputs (rows 0..num_inputs-1 after
 * finalize_for_arithmetization); blinders: the 17 field elements the prover draws from its
 * RNG, in draw order (Montgomery, exactly the limbs `Fr::rand` returns);
 * ext_msg: `extra_transcript_init_msg` (may be NULL).  The Fiat-Shamir transcript
 * (jf-plonk SolidityTranscript, Keccak-256) is computed on the host inside the call.
 * COMPATIBILITY NOTE: the transcript conventions of the fused entry points (capgpu_prove*, the batch
 * and queue calls) are this library's restatement of jf-plonk 0.1.2 @ bcd92b2c (vk field order,
 * little-endian lengths, compressed G1, 48-byte challenge reduction); no upstream vector pins them
 * yet (tests/test_replay.py does as soon as a reference-made CAPFIX01 fixture is present).  Until
 * then only the round-level API below, driven by the caller's own `SolidityTranscript`, is
 * upstream-compatible by construction. */
int capgpu_prove(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs,
                 const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out);

/* Same, with the 5 x n wire values already resident in device memory of ctx's GPU (the public
 * inputs and blinders, < 2 KB, still come from the host). */
int capgpu_prove_dev(capgpu_ctx* ctx, const capgpu_pk* pk, const void* d_wires, const uint64_t* pub_inputs,
                     const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out);

/* `count` independent notes over one proving key, spread over `n_ctxs` contexts.  A `capgpu_pk` is
 * bound to ONE device: every context must live on the key's device (CAPGPU_ERR_ARG otherwise); a
 * multi-GPU host uploads the key once per GPU and makes one call (or one queue) per GPU.  One worker
 * thread per context takes groups of up to `group` notes (capgpu_ctx_set_group) and proves each group
 * in lockstep: 2-4 contexts per GPU saturate a B200.  The counterpart of the reference's rayon loop
 * over notes (/root/reference/src/utils/params_builder.rs:195-233).  wires / pub_inputs / blinders /
 * ext_msgs are arrays of `count` pointers (host memory, pageable or pinned); status (optional)
 * receives the per-note code -- a note whose witness does not satisfy the circuit gets
 * CAPGPU_ERR_DEGREE without disturbing the other notes of its group; the return value is the first
 * non-zero code, or 0. */
int capgpu_prove_batch(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count,
                       const uint64_t* const* wires, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                       const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status);
/* Same with every note's 5 x n wire values already resident in device memory of the contexts' GPU. */
int capgpu_prove_batch_dev(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count,
                           const void* const* d_wires, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                           const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status);

/* ---- asynchronous proving queue ---------------------------------------------------
 * Lets the host overlap witness generation of note k+1 (`TransferCircuit::build` +
 * `check_circuit_satisfiability`, /root/reference/src/proof/transfer.rs:167-177) with proving of note
 * k (:181).  capgpu_submit copies the note's 5 x n wire values from the caller's (pageable) buffer
 * into a slot of a pinned staging ring and returns a ticket at once -- every input buffer may be
 * dropped or reused as soon as it returns; it blocks only while the ring is full (back-pressure).
 * One worker thread per context collects up to `group` pending notes and proves them in lockstep.
 * capgpu_poll: *done = 1 once the proof is ready.  capgpu_wait: blocks, copies the proof, releases
 * the ticket and returns the note's status (0, CAPGPU_ERR_DEGREE, ...).  A ticket is waited on once.
 * ring_slots = 0 lets the library choose (2 x group per context).  All contexts must live on the
 * key's device.  submit / poll / wait may be called from any thread. */
typedef struct capgpu_queue capgpu_queue;
int capgpu_queue_create(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t ring_slots, capgpu_queue** out);
void capgpu_queue_destroy(capgpu_queue* q);
int capgpu_submit(capgpu_queue* q, const uint64_t* wires, const uint64_t* pub_inputs, const uint64_t* blinders,
                  const uint8_t* ext_msg, size_t ext_msg_len, uint64_t* ticket);
int capgpu_poll(capgpu_queue* q, uint64_t ticket, int* done);
int capgpu_wait(capgpu_queue* q, uint64_t ticket, capgpu_proof* out);
/* Counters since creation: notes submitted / completed, lockstep groups formed, host milliseconds
 * spent copying wire values into the ring and waiting for a free slot (all submitting threads). */
int capgpu_queue_stats(capgpu_queue* q, uint64_t* submitted, uint64_t* completed, uint64_t* groups, double* copy_ms,
                       double* wait_slot_ms);

/* ---- round-level API ---------------------------------------------------------------------------
 * For a host that keeps its own transcript (the Rust shim calling upstream's
 * `PlonkTranscript`): challenges come from the caller, commitments / evaluations go back.  */
int capgpu_job_begin(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs, capgpu_job** out);
int capgpu_job_round1(capgpu_job* job, const uint64_t* blinders10, uint64_t* wire_comms_xy /*5x8*/);
int capgpu_job_round2(capgpu_job* job, const uint64_t* beta, const uint64_t* gamma, const uint64_t* blinders3, uint64_t* z_comm_xy /*8*/);
int capgpu_job_round3(capgpu_job* job, const uint64_t* alpha, const uint64_t* blinders4, uint64_t* split_comms_xy /*5x8*/);
int capgpu_job_round4(capgpu_job* job, const uint64_t* zeta, uint64_t* evals /*10x4: 5 wires, 4 sigmas, z(zeta*omega)*/);
int capgpu_job_round5(capgpu_job* job, const uint64_t* v, uint64_t* opening_comms_xy /*2x8*/);
/* Releases the job's context for the next proof.  A round that FAILS (CUDA error, CAPGPU_ERR_DEGREE)
 * releases it too: the context accepts a new capgpu_job_begin / capgpu_prove right away, and the
 * failed handle only accepts capgpu_job_end.  An out-of-order call (CAPGPU_ERR_STATE) leaves the job
 * as it was. */
void capgpu_job_end(capgpu_job* job);

/* ---- introspection for tests / benches ---------------------------------------------------------
 * Copies a device-side intermediate of the last proof of this ctx (slot 0 of its last group) to the host.
 * what: 0 wire polys (5 x (n+2)), 1 z evals (n), 2 z poly (n+3), 4 quotient poly (8n),
 *       5 linearisation poly (n+3), 6 opening poly (n+3), 7 shifted opening poly (n+3),
 *       8 public-input poly (n), 9 split quotient polys (5 x (n+3)).  */
int capgpu_debug_read(capgpu_ctx* ctx, int what, uint64_t* out, size_t max_elems, size_t* n_elems);
/* Number of kernel launches issued through this ctx since creation. */
uint64_t capgpu_launch_count(const capgpu_ctx* ctx);

/* Per-kernel timing with CUDA events on the ctx stream (serialises the ctx while enabled;
 * enabling resets the counters).  id: 0 MSM bucket accumulation (units = mixed additions),
 * 1 NTT tile passes (units = butterflies), 2 quotient evaluation (units = coset points),
 * 3 MSM recode + sort (units = digits), 4 MSM bucket reduction (units = buckets),
 * 5 grand product (units = rows). */
int capgpu_profile_enable(capgpu_ctx* ctx, int on);
int capgpu_profile_read(const capgpu_ctx* ctx, int id, double* total_ms, uint64_t* launches, double* units);

/* ---- calibration --------------------------------------------------------------------------------
 * Measures the integer multiply-add issue rate of this GPU (the MSM / NTT roofline
 * denominator, which MEASURED_PEAKS.json lacks): returns giga warp-lane operations / s for
 * plain IMAD and for IMAD.WIDE.U32, and the Montgomery multiplication rate (G mul/s). */
int capgpu_calibrate(capgpu_ctx* ctx, double* gimad_per_s, double* gimad_wide_per_s, double* gfmul_per_s);

/* ---- the other pairing curves of the reference (src/config.rs:86-114) ---------------------------------
 * CAP is generic over `CapConfig::PairingCurve`; besides the default BN254 the reference builds with the
 * cargo features `bls12_381` / `bls12_377` (Cargo.toml:71-75), whose G1 lives over 381- / 377-bit base
 * fields (ark-ff `Fp384`: 6 x u64 limbs, Montgomery with R = 2^384).  These two entry points replace
 * `VariableBaseMSM::multi_scalar_mul` on `GroupAffine<ark_bls12_381::g1::Parameters>` /
 * `<ark_bls12_377::g1::Parameters>` and the `Fp384` arithmetic under it.  The prover entry points above are
 * BN254 only.
 *   points_xy : n x 12 u64, x[6] || y[6] little-endian Montgomery limbs, all-zero = infinity
 *   scalars   : n x 4 u64 canonical (`into_repr`) scalars of the curve's Fr (255 / 253 bits)
 *   out_xy    : 12 u64, affine result in the same layout
 * capgpu_curve_fq_op: element-wise base-field operation on `count` elements (6 u64 each, Montgomery):
 *   op 0 a*b, 1 a^2, 2 a^-1 (batched binary GCD), 3 a+b, 4 a-b, 5 a^-1 by Fermat (cross-check); b may be
 *   NULL for the unary ones. */
#define CAPGPU_CURVE_BLS12_381 1
#define CAPGPU_CURVE_BLS12_377 2
int capgpu_curve_msm_g1(capgpu_ctx* ctx, int curve, const uint64_t* points_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy);
int capgpu_curve_fq_op(capgpu_ctx* ctx, int curve, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* CAPGPU_H */
