// Device-resident 5-round TurboPlonk prover: the orchestration jf-plonk 0.1.2 performs in
// `PlonkKzgSnark::batch_prove_internal` + `Prover::{run_1st_round .. compute_opening_proofs}`
// for a single circuit, reached from /root/reference/src/proof/transfer.rs:181 (mint.rs:113,
// freeze.rs:151).  Every vector-sized step is a kernel (ntt.cu, msm.cu, poly.cu); the host
// only hashes the transcript and derives O(1) scalars between rounds, and reads back 64-byte
// commitments / 32-byte evaluations.  [UPSTREAM-RECALL: round structure per SURVEY.md App. A.]
//
// LOCKSTEP GROUPS.  A context proves G independent notes over one proving key at the same time:
// every round is issued ONCE for the group, with all vectors of one kind stored as
// [row kind][proof][elements] so that the round's NTT and MSM launches are single batches of
// 5G / 6G / 7G rows, and the per-proof scalars (challenges, blinders, evaluation points) sit in
// device arrays indexed by the proof's slot.  The commitments of the permutation product, the
// openings and the quotient transform -- batches of 1 or 2 vectors when one proof is proved alone
// -- then fill the GPU like the wire commitments do, and one host thread drives G proofs with the
// five host round trips of one.  G = 1 is the single-proof path (capgpu_prove, the round API).
// The reference's counterpart is the rayon loop over notes of
// /root/reference/src/utils/params_builder.rs:195-233.
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "poly.cuh"
#include "prover.h"
#include "transcript.h"

using namespace capgpu;

static const uint64_t kHostRoot28[4] = {0x636e735580d13d9cull, 0xa22bf3742445ffd6ull, 0x56452ac01eb203d8ull, 0x1860ef942963f9e7ull};
static const uint64_t kHostGen[4] = {0x1b0d0ef99fffffe6ull, 0xeaba68a3a32a913full, 0x47d8eb76d8dd0689ull, 0x15d0085520f5bbc3ull};

static inline Fr to_dev(const HFr& x) { Fr r; memcpy(r.v, x.v, 32); return r; }
// rho = 5^((r-1) / (3 * 2^28)) in Montgomery form (ntt.cu kRho28): host_rho(L) has order 3 * 2^L and host_rho(L)^3 = host_omega(L)
static const uint64_t kHostRho28[4] = {0x70f4a36d938bb649ull, 0xb81a7492f4178187ull, 0x22e5d04469f2125eull, 0x0a0b84166fdced23ull};
static inline HFr host_rho(unsigned log_n) {
  HFr w = HFr::from_limbs(kHostRho28);
  for (unsigned i = 0; i < 28 - log_n; i++) w = w.sqr();
  return w;
}
static inline QuotDomain quot_domain(const capgpu_pk* pk) { return QuotDomain{pk->m, pk->qsub, pk->qlog_sub, pk->qstep}; }

static inline HFr host_omega(unsigned log_n) {
  HFr w = HFr::from_limbs(kHostRoot28);
  for (unsigned i = 0; i < 28 - log_n; i++) w = w.sqr();
  return w;
}

struct ProofState {
  HFr beta, gamma, alpha, zeta, v;
  HFr evals[10];
};

struct capgpu_job {
  capgpu_ctx* ctx = nullptr;
  const capgpu_pk* pk = nullptr;
  int round = 0;
  bool busy = false;
  int cap = 0;  // proofs the workspace can hold
  int G = 0;    // proofs of the group being proved (row (r, g) of a [rows][G] block sits at (r * G + g) * stride)
  size_t n = 0, m = 0, NP = 0, num_inputs = 0, pub_stride = 0, cstride = 0;
  DevBuf buf;  // one allocation, carved below
  Fr *wires_eval = nullptr, *polys = nullptr, *z_eval = nullptr, *coset = nullptr, *t = nullptr, *split = nullptr;
  Fr *lin = nullptr, *batch = nullptr, *open = nullptr;  // open: [2][G] rows (opening, shifted opening)
  Fr *num = nullptr, *den = nullptr, *cn = nullptr, *cd = nullptr, *ntt_tmp = nullptr, *evals_dev = nullptr, *pub_dev = nullptr;
  Fr *eval_scratch = nullptr, *div_scratch = nullptr;
  size_t div_tmax = 0;
  G1Affine* comms_dev = nullptr;
  uint32_t* flag = nullptr;
  // per-proof kernel arguments: device arrays and their pinned host mirrors
  BlindArgs *d_blind = nullptr, *h_blind = nullptr;
  GpArgs *d_gp = nullptr, *h_gp = nullptr;
  QuotArgs *d_quot = nullptr, *h_quot = nullptr;
  EvalArgs *d_eval = nullptr, *h_eval = nullptr;
  LinArgs *d_lin = nullptr, *h_lin = nullptr;
  DivArgs *d_div = nullptr, *h_div = nullptr;
  void* pinned = nullptr;  // host mirrors + read-back area
  G1Affine* h_comms = nullptr;
  Fr* h_evals = nullptr;
  Fr* h_pub = nullptr;
  uint32_t* h_flag = nullptr;
  std::vector<ProofState> st;
};

void capgpu_job_free_internal(capgpu_job* job) {
  if (!job) return;
  job->buf.release();
  if (job->pinned) cudaFreeHost(job->pinned);
  delete job;
}

namespace {

struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cap_hint: capacity to allocate if the workspace has to be (re)built -- the batch paths pass the
// context's group size so that a first, smaller group does not cause a second allocation later
capgpu_job* job_acquire(capgpu_ctx* ctx, const capgpu_pk* pk, int G, int cap_hint = 1) {
  CAPGPU_REQUIRE(pk->device == ctx->device, "proving key lives on another device");
  CAPGPU_REQUIRE(G >= 1 && G <= CAPGPU_MAX_GROUP, "group size out of range");
  capgpu_job* job = ctx->cached_job;
  if (job && job->busy) throw CodeError{CAPGPU_ERR_STATE};
  if (job && (job->n != pk->n || job->m != pk->m || job->num_inputs < pk->num_inputs || job->cap < G)) {
    capgpu_job_free_internal(job);
    ctx->cached_job = job = nullptr;
  }
  if (!job) {
    std::unique_ptr<capgpu_job, void (*)(capgpu_job*)> holder(new capgpu_job(), capgpu_job_free_internal);
    job = holder.get();
    job->ctx = ctx;
    job->cap = G > cap_hint ? G : cap_hint;
    job->n = pk->n;
    job->m = pk->m;
    job->NP = pk->n + 8;
    job->num_inputs = pk->num_inputs;
    job->pub_stride = align_up(pk->num_inputs + 1, 8);
    const size_t n = job->n, m = job->m, NP = job->NP, C = (size_t)job->cap;
    job->cstride = n / 8 + 8;
    job->div_tmax = (NP + 15) / 16 + 1;
    const size_t elems = C * (6 * NP + 7 * NP + n + 7 * m + m + 5 * NP + 4 * NP + 2 * n + 2 * job->cstride + 7 * m + 16 + 160 +
                              2 * job->div_tmax + job->pub_stride);
    const size_t arg_bytes = C * (align_up(sizeof(BlindArgs), 32) + align_up(sizeof(GpArgs), 32) + align_up(sizeof(QuotArgs), 32) +
                                  align_up(sizeof(EvalArgs), 32) + align_up(sizeof(LinArgs), 32) + align_up(sizeof(DivArgs), 32));
    const size_t bytes = elems * sizeof(Fr) + arg_bytes + 5 * C * sizeof(G1Affine) + align_up(C * sizeof(uint32_t), 32) + 256;
    job->buf.reserve(bytes);
    Fr* p = job->buf.as<Fr>();
    auto take = [&](size_t cnt) { Fr* r = p; p += cnt; return r; };
    job->wires_eval = take(C * 6 * NP);  // [6][G] rows of n evaluations (tail: blinding scalars of the Lagrange commit)
    job->polys = take(C * 7 * NP);       // [7][G]: w0..w4, PI, z
    job->z_eval = take(C * n);
    job->coset = take(C * 7 * m);
    job->t = take(C * m);
    job->split = take(C * 5 * NP);
    job->lin = take(C * NP); job->batch = take(C * NP); job->open = take(C * 2 * NP);
    job->num = take(C * n); job->den = take(C * n);
    job->cn = take(C * job->cstride); job->cd = take(C * job->cstride);
    job->ntt_tmp = take(C * 7 * m);
    job->evals_dev = take(C * 16);
    job->eval_scratch = take(C * 160);
    job->div_scratch = take(C * 2 * job->div_tmax);
    job->pub_dev = take(C * job->pub_stride);
    char* q = reinterpret_cast<char*>(p);
    auto take_bytes = [&](size_t b) { char* r = q; q += align_up(b, 32); return r; };
    job->d_blind = reinterpret_cast<BlindArgs*>(take_bytes(C * sizeof(BlindArgs)));
    job->d_gp = reinterpret_cast<GpArgs*>(take_bytes(C * sizeof(GpArgs)));
    job->d_quot = reinterpret_cast<QuotArgs*>(take_bytes(C * sizeof(QuotArgs)));
    job->d_eval = reinterpret_cast<EvalArgs*>(take_bytes(C * sizeof(EvalArgs)));
    job->d_lin = reinterpret_cast<LinArgs*>(take_bytes(C * sizeof(LinArgs)));
    job->d_div = reinterpret_cast<DivArgs*>(take_bytes(C * sizeof(DivArgs)));
    job->comms_dev = reinterpret_cast<G1Affine*>(take_bytes(5 * C * sizeof(G1Affine)));
    job->flag = reinterpret_cast<uint32_t*>(take_bytes(C * sizeof(uint32_t)));
    // pinned mirrors
    const size_t pin_bytes = arg_bytes + 5 * C * sizeof(G1Affine) + C * 16 * sizeof(Fr) + C * job->pub_stride * sizeof(Fr) +
                             align_up(C * sizeof(uint32_t), 32) + 256;
    CAPGPU_CUDA(cudaMallocHost(&job->pinned, pin_bytes));
    q = static_cast<char*>(job->pinned);
    job->h_blind = reinterpret_cast<BlindArgs*>(take_bytes(C * sizeof(BlindArgs)));
    job->h_gp = reinterpret_cast<GpArgs*>(take_bytes(C * sizeof(GpArgs)));
    job->h_quot = reinterpret_cast<QuotArgs*>(take_bytes(C * sizeof(QuotArgs)));
    job->h_eval = reinterpret_cast<EvalArgs*>(take_bytes(C * sizeof(EvalArgs)));
    job->h_lin = reinterpret_cast<LinArgs*>(take_bytes(C * sizeof(LinArgs)));
    job->h_div = reinterpret_cast<DivArgs*>(take_bytes(C * sizeof(DivArgs)));
    job->h_comms = reinterpret_cast<G1Affine*>(take_bytes(5 * C * sizeof(G1Affine)));
    job->h_evals = reinterpret_cast<Fr*>(take_bytes(C * 16 * sizeof(Fr)));
    job->h_pub = reinterpret_cast<Fr*>(take_bytes(C * job->pub_stride * sizeof(Fr)));
    job->h_flag = reinterpret_cast<uint32_t*>(take_bytes(C * sizeof(uint32_t)));
    job->st.resize(C);
    ctx->cached_job = holder.release();
  }
  job->pk = pk;
  job->round = 0;
  job->G = G;
  job->busy = true;
  return job;
}

template <class T>
void upload_args(capgpu_job* job, T* dev, const T* host) {
  CAPGPU_CUDA(cudaMemcpyAsync(dev, host, job->G * sizeof(T), cudaMemcpyHostToDevice, job->ctx->stream));
}

// notes[g]: wire values (5 x n, host or device) and public inputs of proof g
void job_begin(capgpu_job* job, const NoteIn* notes) {
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  const int G = job->G;
  for (int g = 0; g < G; g++) {
    // rows (i, g), i < 5: destination pitch G * NP
    CAPGPU_CUDA(cudaMemcpy2DAsync(job->wires_eval + (size_t)g * NP, (size_t)G * NP * sizeof(Fr), notes[g].wires, n * sizeof(Fr),
                                  n * sizeof(Fr), 5, notes[g].wires_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                  ctx->stream));
    if (pk->num_inputs) memcpy(job->h_pub + (size_t)g * job->pub_stride, notes[g].pub_inputs, pk->num_inputs * sizeof(Fr));
  }
  if (pk->num_inputs)
    CAPGPU_CUDA(cudaMemcpyAsync(job->pub_dev, job->h_pub, (size_t)G * job->pub_stride * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  fill_pi(ctx, job->wires_eval + (size_t)5 * G * NP, NP, n, G, job->pub_dev, job->pub_stride, pk->num_inputs);
  job->round = 1;
}

// commitments of `rows` row kinds: device order (r, g) -> out[g] + r * 8
void read_points(capgpu_job* job, int rows, uint64_t* const* out) {
  capgpu_ctx* ctx = job->ctx;
  const int G = job->G;
  CAPGPU_CUDA(cudaMemcpyAsync(job->h_comms, job->comms_dev, (size_t)rows * G * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  ctx_wait(ctx);
  for (int g = 0; g < G; g++)
    for (int r = 0; r < rows; r++) memcpy(out[g] + 8 * r, &job->h_comms[r * G + g], sizeof(G1Affine));
}

void round1(capgpu_job* job, const uint64_t* const* blinders10, uint64_t* const* wire_comms) {
  if (job->round != 1) throw CodeError{CAPGPU_ERR_STATE};
  NvtxRange nv("capgpu round 1: wire polynomials + commitments");
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  const int G = job->G;
  // 5 wire polynomials + the public-input polynomial of every proof: one batched INTT
  ntt_device(ctx, pk->log_n, job->wires_eval, n, NP, job->polys, NP, job->ntt_tmp, 6 * G, true, false);
  for (int g = 0; g < G; g++) {
    memcpy(job->h_blind[g].b, blinders10[g], 10 * sizeof(Fr));
    job->h_blind[g].rows_blinded = 5;
  }
  upload_args(job, job->d_blind, job->h_blind);
  blind(ctx, job->polys, NP, n, 6, G, 2, job->d_blind);
  if (pk->lag && pk->use_lag) {
    // commit from the evaluations: sum_j w_j L_j + (b0 + b1 X)(X^n - 1) at tau — the same group element as
    // the coefficient-form MSM, but zero / small witness cells cost nothing / one window
    lagrange_tail(ctx, job->wires_eval, NP, n, G, job->d_blind);
    msm_device(ctx, pk->lag, 0, job->wires_eval, n + 4, NP, 5 * G, true, job->comms_dev, ctx->latency_mode);
  } else {
    msm_device(ctx, pk->srs, 0, job->polys, n + 2, NP, 5 * G, true, job->comms_dev, ctx->latency_mode);
  }
  read_points(job, 5, wire_comms);
  job->round = 2;
}

void round2(capgpu_job* job, const uint64_t* const* beta, const uint64_t* const* gamma, const uint64_t* const* blinders3,
            uint64_t* const* z_comm) {
  if (job->round != 2) throw CodeError{CAPGPU_ERR_STATE};
  NvtxRange nv("capgpu round 2: permutation grand product");
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  const int G = job->G;
  for (int g = 0; g < G; g++) {
    ProofState& s = job->st[g];
    s.beta = HFr::from_limbs(beta[g]);
    s.gamma = HFr::from_limbs(gamma[g]);
    GpArgs& ga = job->h_gp[g];
    ga.beta = to_dev(s.beta);
    ga.gamma = to_dev(s.gamma);
    for (int i = 0; i < 5; i++) ga.k[i] = to_dev(pk->k[i]);
    memcpy(job->h_blind[g].b, blinders3[g], 3 * sizeof(Fr));
    job->h_blind[g].rows_blinded = 1;
  }
  upload_args(job, job->d_gp, job->h_gp);
  upload_args(job, job->d_blind, job->h_blind);
  grand_product(ctx, job->wires_eval, NP, pk->sig_eval, pk->omega_n, n, G, job->d_gp, job->num, job->den, job->cn, job->cd, job->cstride,
                job->z_eval);
  Fr* zp = job->polys + (size_t)6 * G * NP;
  ntt_device(ctx, pk->log_n, job->z_eval, n, n, zp, NP, job->ntt_tmp, G, true, false);
  blind(ctx, zp, NP, n, 1, G, 3, job->d_blind);
  msm_device(ctx, pk->srs, 0, zp, n + 3, NP, G, true, job->comms_dev, ctx->latency_mode);
  read_points(job, 1, z_comm);
  job->round = 3;
}

// status[g] receives CAPGPU_ERR_DEGREE for a proof whose quotient has the wrong degree (the
// witness does not satisfy the circuit); the other proofs of the group are unaffected
void round3(capgpu_job* job, const uint64_t* const* alpha, const uint64_t* const* blinders4, uint64_t* const* split_comms, int* status) {
  if (job->round != 3) throw CodeError{CAPGPU_ERR_STATE};
  NvtxRange nv("capgpu round 3: quotient polynomial");
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, m = job->m, NP = job->NP;
  const int G = job->G;
  for (int g = 0; g < G; g++) {
    ProofState& s = job->st[g];
    s.alpha = HFr::from_limbs(alpha[g]);
    QuotArgs& qa = job->h_quot[g];
    qa.alpha = to_dev(s.alpha);
    qa.alpha2 = to_dev(s.alpha.sqr());
    qa.beta = to_dev(s.beta);
    qa.gamma = to_dev(s.gamma);
    for (int i = 0; i < 5; i++) qa.k[i] = to_dev(pk->k[i]);
    memcpy(job->h_blind[g].b, blinders4[g], 4 * sizeof(Fr));
    job->h_blind[g].rows_blinded = 4;
  }
  upload_args(job, job->d_quot, job->h_quot);
  upload_args(job, job->d_blind, job->h_blind);
  // evaluations of the 5 wire polys, PI and z on the quotient domain (selectors / sigmas are cached in the pk): 6n points as three
  // 2n-point cosets, or the 8n-point coset for tiny circuits
  if (pk->q3) ntt3_forward(ctx, pk->log_n + 1, job->polys, n + 3, NP, job->coset, job->ntt_tmp, 7 * G);
  else ntt_device(ctx, pk->log_n + 3, job->polys, n + 3, NP, job->coset, m, job->ntt_tmp, 7 * G, false, true);
  quotient_evals(ctx, job->coset, pk->sel_coset, pk->sig_coset, pk->xs, pk->l1inv, pk->zh_inv, quot_domain(pk), G, job->d_quot, job->t);
  if (pk->q3) ntt3_inverse(ctx, pk->log_n + 1, job->t, job->ntt_tmp, G);
  else ntt_device(ctx, pk->log_n + 3, job->t, m, m, job->t, m, job->ntt_tmp, G, true, true);
  split_quotient(ctx, job->t, n, m, G, job->split, NP, job->d_blind, job->flag);
  msm_device(ctx, pk->srs, 0, job->split, n + 3, NP, 5 * G, true, job->comms_dev, ctx->latency_mode);
  CAPGPU_CUDA(cudaMemcpyAsync(job->h_flag, job->flag, G * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  read_points(job, 5, split_comms);
  for (int g = 0; g < G; g++) status[g] = job->h_flag[g] ? CAPGPU_ERR_DEGREE : CAPGPU_OK;
  job->round = 4;
}

void round4(capgpu_job* job, const uint64_t* const* zeta, uint64_t* const* evals_out) {
  if (job->round != 4) throw CodeError{CAPGPU_ERR_STATE};
  NvtxRange nv("capgpu round 4: evaluations");
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  const int G = job->G;
  const HFr omega = host_omega(pk->log_n);
  for (int g = 0; g < G; g++) {
    ProofState& s = job->st[g];
    s.zeta = HFr::from_limbs(zeta[g]);
    HFr zeta_w = s.zeta * omega;
    EvalArgs& ea = job->h_eval[g];
    for (int i = 0; i < 5; i++) { ea.poly[i] = job->polys + ((size_t)i * G + g) * NP; ea.len[i] = n + 2; ea.x[i] = to_dev(s.zeta); }
    for (int i = 0; i < 4; i++) { ea.poly[5 + i] = pk->sig_coef + (size_t)i * n; ea.len[5 + i] = n; ea.x[5 + i] = to_dev(s.zeta); }
    ea.poly[9] = job->polys + ((size_t)6 * G + g) * NP; ea.len[9] = n + 3; ea.x[9] = to_dev(zeta_w);
  }
  upload_args(job, job->d_eval, job->h_eval);
  evaluate(ctx, job->d_eval, 10, G, job->evals_dev, job->eval_scratch);
  CAPGPU_CUDA(cudaMemcpyAsync(job->h_evals, job->evals_dev, (size_t)G * 16 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  ctx_wait(ctx);
  for (int g = 0; g < G; g++) {
    memcpy(evals_out[g], job->h_evals + (size_t)g * 16, 10 * sizeof(Fr));
    for (int i = 0; i < 10; i++) job->st[g].evals[i] = HFr::from_limbs(evals_out[g] + 4 * i);
  }
  job->round = 5;
}

void round5(capgpu_job* job, const uint64_t* const* v_in, uint64_t* const* opening_comms) {
  if (job->round != 5) throw CodeError{CAPGPU_ERR_STATE};
  NvtxRange nv("capgpu round 5: linearisation + opening proofs");
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  const int G = job->G;
  const HFr one = HFr::one();
  const HFr omega = host_omega(pk->log_n);
  const HFr n_mont = HFr::from_u64(n);
  for (int g = 0; g < G; g++) {
    ProofState& s = job->st[g];
    s.v = HFr::from_limbs(v_in[g]);
    const HFr* w = s.evals;           // wires_evals[5]
    const HFr* se = s.evals + 5;      // wire_sigma_evals[4]
    const HFr zw = s.evals[9];        // perm_next_eval
    const HFr &alpha = s.alpha, &beta = s.beta, &gamma = s.gamma, &zeta = s.zeta;
    HFr zh = zeta.pow_u64(n) - one;
    HFr l1 = zh * (n_mont * (zeta - one)).inv();
    LinArgs& la = job->h_lin[g];
    la.polys = job->polys + (size_t)g * NP;
    la.split = job->split + (size_t)g * NP;
    la.rstride = (size_t)G * NP;
    HFr w01 = w[0] * w[1], w23 = w[2] * w[3];
    auto p5 = [](const HFr& x) { HFr x2 = x.sqr(); return x2.sqr() * x; };
    HFr sel[13] = {w[0], w[1], w[2], w[3], w01, w23, p5(w[0]), p5(w[1]), p5(w[2]), p5(w[3]), w[4].neg(), one, w01 * w23 * w[4]};
    for (int i = 0; i < 13; i++) la.cs_sel[i] = to_dev(sel[i]);
    HFr cz = alpha;
    HFr bz = beta * zeta;
    for (int j = 0; j < 5; j++) cz = cz * (w[j] + pk->k[j] * bz + gamma);
    cz = cz + alpha.sqr() * l1;
    la.cz = to_dev(cz);
    HFr cs = alpha * beta * zw;
    for (int j = 0; j < 4; j++) cs = cs * (w[j] + beta * se[j] + gamma);
    la.csig = to_dev(cs.neg());
    HFr zn2 = (zh + one) * zeta * zeta;
    HFr c = one;
    for (int i = 0; i < 5; i++) { la.ct[i] = to_dev((zh * c).neg()); c = c * zn2; }
    HFr vp = s.v;
    for (int i = 0; i < 9; i++) { la.vp[i] = to_dev(vp); vp = vp * s.v; }
    DivArgs& da = job->h_div[g];
    const HFr zeta_w = zeta * omega;
    da.src[0] = job->batch + (size_t)g * NP; da.dst[0] = job->open + (size_t)g * NP; da.len[0] = n + 3;
    da.x[0] = to_dev(zeta); da.xinv[0] = to_dev(zeta.inv());
    da.src[1] = job->polys + ((size_t)6 * G + g) * NP; da.dst[1] = job->open + ((size_t)G + g) * NP; da.len[1] = n + 3;
    da.x[1] = to_dev(zeta_w); da.xinv[1] = to_dev(zeta_w.inv());
  }
  upload_args(job, job->d_lin, job->h_lin);
  upload_args(job, job->d_div, job->h_div);
  lin_batch(ctx, job->d_lin, pk->sel_coef, pk->sig_coef, n, n + 3, G, job->lin, job->batch, NP);
  divide_linear(ctx, job->d_div, 2, G, n + 3, job->div_scratch, job->div_tmax);
  // rows (0, g) = opening quotient, (1, g) = shifted opening quotient
  msm_device(ctx, pk->srs, 0, job->open, n + 2, NP, 2 * G, true, job->comms_dev, ctx->latency_mode);
  read_points(job, 2, opening_comms);
  job->round = 6;
}

// ---- proving key construction -------------------------------------------------------------
void pk_finish(capgpu_ctx* ctx, capgpu_pk* pk) {
  const size_t n = pk->n, m = pk->m;
  CAPGPU_CUDA(cudaMalloc(&pk->sig_eval, 5 * n * sizeof(Fr)));
  CAPGPU_CUDA(cudaMalloc(&pk->sel_coset, 13 * m * sizeof(Fr)));
  CAPGPU_CUDA(cudaMalloc(&pk->sig_coset, 5 * m * sizeof(Fr)));
  CAPGPU_CUDA(cudaMalloc(&pk->xs, m * sizeof(Fr)));
  CAPGPU_CUDA(cudaMalloc(&pk->l1inv, m * sizeof(Fr)));
  CAPGPU_CUDA(cudaMalloc(&pk->zh_inv, 8 * sizeof(Fr)));
  const unsigned log_sub = pk->qlog_sub;
  CAPGPU_CUDA(cudaMalloc(&pk->omega_n, n * sizeof(Fr)));
  ctx->ntt_tmp.reserve(5 * m * sizeof(Fr));
  Fr* tmp = ctx->ntt_tmp.as<Fr>();
  ntt_device(ctx, pk->log_n, pk->sig_coef, n, n, pk->sig_eval, n, tmp, 5, false, false);
  for (int s = 0; s < 13; s++) {
    if (pk->q3) ntt3_forward(ctx, log_sub, pk->sel_coef + (size_t)s * n, n, n, pk->sel_coset + (size_t)s * m, tmp, 1);
    else ntt_device(ctx, log_sub, pk->sel_coef + (size_t)s * n, n, n, pk->sel_coset + (size_t)s * m, m, tmp, 1, false, true);
  }
  if (pk->q3) ntt3_forward(ctx, log_sub, pk->sig_coef, n, n, pk->sig_coset, tmp, 5);
  else ntt_device(ctx, log_sub, pk->sig_coef, n, n, pk->sig_coset, m, tmp, 5, false, true);
  CAPGPU_CUDA(cudaMemcpyAsync(pk->omega_n, domain_omega_powers(ctx, pk->log_n), n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  {
    const char* e = getenv("CAPGPU_LAGRANGE");
    if (!(e && atoi(e) == 0) && pk->srs->n >= n + 2)
      pk->lag = srs_lagrange(ctx, pk->srs, pk->log_n, pk->omega_n, to_dev(HFr::from_u64(n).inv()));
  }
  // coset shifts s_k = g rho^k (one coset, s_0 = g, on the 8n domain); 1 / Z_H has period `qstep` along a coset
  const HFr gen = HFr::from_limbs(kHostGen);
  const HFr rho = pk->q3 ? host_rho(log_sub) : HFr::one();
  const int cosets = pk->q3 ? 3 : 1;
  CosetShifts shifts;
  HFr sk = gen;
  const HFr wsub = host_omega(log_sub);
  Fr zh[8];
  for (int k = 0; k < 3; k++) {
    shifts.s[k] = to_dev(sk);
    if (k < cosets) {
      HFr x = sk;
      for (unsigned i = 0; i < pk->qstep; i++) {
        zh[k * pk->qstep + i] = to_dev((x.pow_u64(n) - HFr::one()).inv());
        x = x * wsub;
      }
    }
    sk = sk * rho;
  }
  coset_tables(ctx, domain_omega_powers(ctx, log_sub), quot_domain(pk), shifts, to_dev(HFr::from_u64(n)), pk->xs, pk->l1inv);
  CAPGPU_CUDA(cudaMemcpyAsync(pk->zh_inv, zh, cosets * pk->qstep * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  ctx_wait(ctx);
  // transcript bytes of the verifying key (SolidityTranscript::append_vk_and_pub_input, minus the inputs)
  SolidityTranscript t;
  t.append_u64_le(254);
  t.append_u64_le(pk->n);
  t.append_u64_le(pk->num_inputs);
  for (int i = 0; i < 5; i++) t.append_field(pk->k[i]);
  for (int i = 0; i < 13; i++) t.append_commitment(pk->sel_comms[i]);
  for (int i = 0; i < 5; i++) t.append_commitment(pk->sig_comms[i]);
  pk->vk_bytes = t.transcript;
}

capgpu_pk* pk_alloc(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* k) {
  CAPGPU_REQUIRE(log_n >= 2 && log_n <= 17, "domain size must be 2^2 .. 2^17");
  capgpu_pk* pk = new capgpu_pk();
  pk->device = ctx->device;
  pk->log_n = log_n;
  pk->n = (size_t)1 << log_n;
  // Quotient domain: t has degree 5n + 7, so 6n points (three cosets of 2n, ntt.cu) carry it when n >= 8; tiny circuits and
  // CAPGPU_QDOMAIN=8 keep the 8n-point coset of the next power of two (ark-poly's choice — same polynomial either way).
  const char* qenv = getenv("CAPGPU_QDOMAIN");  // read per key, so one process can hold keys of both kinds (tests)
  const bool force8 = qenv && atoi(qenv) == 8;
  pk->q3 = !force8 && log_n >= 6;
  pk->qlog_sub = pk->q3 ? log_n + 1 : log_n + 3;
  pk->qsub = (size_t)1 << pk->qlog_sub;
  pk->qstep = (unsigned)(pk->qsub / pk->n);
  pk->m = pk->q3 ? 3 * pk->qsub : pk->qsub;
  pk->num_inputs = num_inputs;
  pk->srs = srs;
  for (int i = 0; i < 5; i++) pk->k[i] = HFr::from_limbs(k + 4 * i);
  return pk;
}

void pk_free(capgpu_pk* pk) {
  if (!pk) return;
  cudaSetDevice(pk->device);
  Fr* ptrs[] = {pk->sel_coef, pk->sig_coef, pk->sig_eval, pk->sel_coset, pk->sig_coset, pk->xs, pk->l1inv, pk->zh_inv, pk->omega_n};
  for (Fr* p : ptrs) if (p) cudaFree(p);
  if (pk->lag) capgpu_srs_destroy(pk->lag);
  if (pk->owned_srs) capgpu_srs_destroy(pk->owned_srs);
  delete pk;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------
namespace capgpu {

// shared by capgpu_pk_upload and the CanonicalSerialize loader (formats.cu)
int pk_create_from_coefficients(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* selectors,
                                const uint64_t* sigmas, const uint64_t* k, const uint64_t* selector_comms_xy,
                                const uint64_t* sigma_comms_xy, capgpu_pk** out) {
  *out = nullptr;
  capgpu_pk* pk = nullptr;
  int rc = guarded(ctx, [&] {
    pk = pk_alloc(ctx, srs, log_n, num_inputs, k);
    CAPGPU_REQUIRE(num_inputs < pk->n, "more public inputs than rows");
    if (srs->n < pk->n + 3) throw CodeError{CAPGPU_ERR_SRS_TOO_SMALL};
    const size_t n = pk->n;
    CAPGPU_CUDA(cudaMalloc(&pk->sel_coef, 13 * n * sizeof(Fr)));
    CAPGPU_CUDA(cudaMalloc(&pk->sig_coef, 5 * n * sizeof(Fr)));
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sel_coef, selectors, 13 * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sig_coef, sigmas, 5 * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    memcpy(pk->sel_comms, selector_comms_xy, sizeof pk->sel_comms);
    memcpy(pk->sig_comms, sigma_comms_xy, sizeof pk->sig_comms);
    pk_finish(ctx, pk);
  });
  if (rc != CAPGPU_OK) { pk_free(pk); return rc; }
  *out = pk;
  return CAPGPU_OK;
}

}  // namespace capgpu

extern "C" int capgpu_pk_upload(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* selectors,
                                const uint64_t* sigmas, const uint64_t* k, const uint64_t* selector_comms_xy,
                                const uint64_t* sigma_comms_xy, capgpu_pk** out) {
  if (!ctx || !srs || !selectors || !sigmas || !k || !selector_comms_xy || !sigma_comms_xy || !out) return CAPGPU_ERR_ARG;
  return pk_create_from_coefficients(ctx, srs, log_n, num_inputs, selectors, sigmas, k, selector_comms_xy, sigma_comms_xy, out);
}

extern "C" int capgpu_preprocess(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs,
                                 const uint64_t* selector_evals, const uint64_t* sigma_evals, const uint64_t* k, capgpu_pk** out) {
  if (!ctx || !srs || !selector_evals || !sigma_evals || !k || !out) return CAPGPU_ERR_ARG;
  *out = nullptr;
  capgpu_pk* pk = nullptr;
  int rc = guarded(ctx, [&] {
    pk = pk_alloc(ctx, srs, log_n, num_inputs, k);
    CAPGPU_REQUIRE(num_inputs < pk->n, "more public inputs than rows");
    if (srs->n < pk->n + 3) throw CodeError{CAPGPU_ERR_SRS_TOO_SMALL};
    const size_t n = pk->n;
    CAPGPU_CUDA(cudaMalloc(&pk->sel_coef, 13 * n * sizeof(Fr)));
    CAPGPU_CUDA(cudaMalloc(&pk->sig_coef, 5 * n * sizeof(Fr)));
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sel_coef, selector_evals, 13 * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sig_coef, sigma_evals, 5 * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    ctx->ntt_tmp.reserve(13 * n * sizeof(Fr));
    ntt_device(ctx, log_n, pk->sel_coef, n, n, pk->sel_coef, n, ctx->ntt_tmp.as<Fr>(), 13, true, false);
    ntt_device(ctx, log_n, pk->sig_coef, n, n, pk->sig_coef, n, ctx->ntt_tmp.as<Fr>(), 5, true, false);
    ctx->msm_out.reserve(18 * sizeof(G1Affine));
    G1Affine* outp = ctx->msm_out.as<G1Affine>();
    // the selector and sigma coefficient blocks are separate allocations: two batched MSMs
    msm_device(ctx, srs, 0, pk->sel_coef, n, n, 13, true, outp);
    msm_device(ctx, srs, 0, pk->sig_coef, n, n, 5, true, outp + 13);
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sel_comms, outp, 13 * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaMemcpyAsync(pk->sig_comms, outp + 13, 5 * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    pk_finish(ctx, pk);
  });
  if (rc != CAPGPU_OK) { pk_free(pk); return rc; }
  *out = pk;
  return CAPGPU_OK;
}

extern "C" int capgpu_pk_export(capgpu_ctx* ctx, const capgpu_pk* pk, uint64_t* selectors, uint64_t* sigmas,
                                uint64_t* selector_comms_xy, uint64_t* sigma_comms_xy) {
  if (!ctx || !pk) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    const size_t n = pk->n;
    if (selectors) CAPGPU_CUDA(cudaMemcpyAsync(selectors, pk->sel_coef, 13 * n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (sigmas) CAPGPU_CUDA(cudaMemcpyAsync(sigmas, pk->sig_coef, 5 * n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    if (selector_comms_xy) memcpy(selector_comms_xy, pk->sel_comms, sizeof pk->sel_comms);
    if (sigma_comms_xy) memcpy(sigma_comms_xy, pk->sig_comms, sizeof pk->sig_comms);
  });
}

extern "C" void capgpu_pk_destroy(capgpu_pk* pk) { pk_free(pk); }

extern "C" int capgpu_pk_info(const capgpu_pk* pk, unsigned* log_n, size_t* num_inputs, uint64_t* k) {
  if (!pk) return CAPGPU_ERR_ARG;
  if (log_n) *log_n = pk->log_n;
  if (num_inputs) *num_inputs = pk->num_inputs;
  if (k) for (int i = 0; i < 5; i++) memcpy(k + 4 * i, pk->k[i].v, 32);
  return CAPGPU_OK;
}

extern "C" int capgpu_pk_lagrange(capgpu_pk* pk, int enable) {
  if (!pk) return CAPGPU_ERR_ARG;
  if (enable && !pk->lag) return CAPGPU_ERR_STATE;
  pk->use_lag = enable != 0;
  return CAPGPU_OK;
}

extern "C" int capgpu_pk_lagrange_export(capgpu_ctx* ctx, const capgpu_pk* pk, uint64_t* points_xy, size_t count) {
  if (!ctx || !pk || !points_xy) return CAPGPU_ERR_ARG;
  if (!pk->lag) return CAPGPU_ERR_STATE;
  return capgpu_srs_export(ctx, pk->lag, points_xy, count);
}

extern "C" int capgpu_job_begin(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs, capgpu_job** out) {
  if (!ctx || !pk || !wires || !out || (!pub_inputs && pk->num_inputs)) return CAPGPU_ERR_ARG;
  *out = nullptr;
  return guarded(ctx, [&] {
    capgpu_job* job = job_acquire(ctx, pk, 1);
    try {
      NoteIn in{wires, false, pub_inputs, nullptr, nullptr, 0};
      job_begin(job, &in);
    } catch (...) {
      job->busy = false;
      throw;
    }
    *out = job;
  });
}

// A round that fails (CUDA error, wrong quotient degree, out-of-order call) releases the job: the
// context is immediately usable for a new capgpu_job_begin / capgpu_prove, and the failed job
// handle only accepts capgpu_job_end (a no-op then).
template <class F>
static int job_round(capgpu_job* job, F&& body) {
  if (!job || !job->busy) return CAPGPU_ERR_STATE;
  int rc = guarded(job->ctx, body);
  if (rc != CAPGPU_OK && rc != CAPGPU_ERR_STATE) job->busy = false;
  return rc;
}

extern "C" int capgpu_job_round1(capgpu_job* job, const uint64_t* blinders10, uint64_t* wire_comms_xy) {
  if (!blinders10 || !wire_comms_xy) return CAPGPU_ERR_ARG;
  return job_round(job, [&] { round1(job, &blinders10, &wire_comms_xy); });
}
extern "C" int capgpu_job_round2(capgpu_job* job, const uint64_t* beta, const uint64_t* gamma, const uint64_t* blinders3, uint64_t* z_comm_xy) {
  if (!beta || !gamma || !blinders3 || !z_comm_xy) return CAPGPU_ERR_ARG;
  return job_round(job, [&] { round2(job, &beta, &gamma, &blinders3, &z_comm_xy); });
}
extern "C" int capgpu_job_round3(capgpu_job* job, const uint64_t* alpha, const uint64_t* blinders4, uint64_t* split_comms_xy) {
  if (!alpha || !blinders4 || !split_comms_xy) return CAPGPU_ERR_ARG;
  return job_round(job, [&] {
    int st = CAPGPU_OK;
    round3(job, &alpha, &blinders4, &split_comms_xy, &st);
    if (st != CAPGPU_OK) throw CodeError{st};
  });
}
extern "C" int capgpu_job_round4(capgpu_job* job, const uint64_t* zeta, uint64_t* evals) {
  if (!zeta || !evals) return CAPGPU_ERR_ARG;
  return job_round(job, [&] { round4(job, &zeta, &evals); });
}
extern "C" int capgpu_job_round5(capgpu_job* job, const uint64_t* v, uint64_t* opening_comms_xy) {
  if (!v || !opening_comms_xy) return CAPGPU_ERR_ARG;
  return job_round(job, [&] { round5(job, &v, &opening_comms_xy); });
}
extern "C" void capgpu_job_end(capgpu_job* job) {
  if (job) job->busy = false;
}

// ---- whole proofs: one lockstep group --------------------------------------------------------
namespace capgpu {

// Proves notes[0..G) in lockstep on ctx.  status[g]: per-proof result (CAPGPU_ERR_DEGREE for an
// unsatisfied circuit; the rest of the group still completes).  Throws on CUDA / argument errors
// (the whole group fails).  `inputs_consumed` (optional) is called once the wire values have been
// read from the callers' buffers (end of round 1).
void prove_group(capgpu_ctx* ctx, const capgpu_pk* pk, int G, const NoteIn* notes, capgpu_proof* const* out, int* status,
                 const std::function<void()>& inputs_consumed, int cap_hint) {
  capgpu_job* job = job_acquire(ctx, pk, G, cap_hint);
  struct Release { capgpu_job* j; ~Release() { j->busy = false; } } release{job};
  job_begin(job, notes);
  std::vector<SolidityTranscript> tr(G);
  std::vector<HFr> ch(2 * G);
  std::vector<const uint64_t*> p0(G), p1(G), p2(G);
  std::vector<uint64_t*> o0(G);
  for (int g = 0; g < G; g++) {
    if (notes[g].ext_msg_len) tr[g].append_message(notes[g].ext_msg, notes[g].ext_msg_len);
    tr[g].append_message(pk->vk_bytes.data(), pk->vk_bytes.size());
    for (size_t i = 0; i < pk->num_inputs; i++) tr[g].append_field(HFr::from_limbs(notes[g].pub_inputs + 4 * i));
    status[g] = CAPGPU_OK;
  }
  // Round 1
  for (int g = 0; g < G; g++) { p0[g] = notes[g].blinders; o0[g] = &out[g]->wires_poly_comms[0][0]; }
  round1(job, p0.data(), o0.data());
  if (inputs_consumed) inputs_consumed();
  // Round 2
  for (int g = 0; g < G; g++) {
    for (int i = 0; i < 5; i++) tr[g].append_commitment(out[g]->wires_poly_comms[i]);
    ch[2 * g] = tr[g].get_and_append_challenge();
    ch[2 * g + 1] = tr[g].get_and_append_challenge();
    p0[g] = ch[2 * g].v; p1[g] = ch[2 * g + 1].v; p2[g] = notes[g].blinders + 4 * 10;
    o0[g] = out[g]->prod_perm_poly_comm;
  }
  round2(job, p0.data(), p1.data(), p2.data(), o0.data());
  // Round 3
  for (int g = 0; g < G; g++) {
    tr[g].append_commitment(out[g]->prod_perm_poly_comm);
    ch[g] = tr[g].get_and_append_challenge();
    p0[g] = ch[g].v; p1[g] = notes[g].blinders + 4 * 13;
    o0[g] = &out[g]->split_quot_poly_comms[0][0];
  }
  round3(job, p0.data(), p1.data(), o0.data(), status);
  // Round 4
  std::vector<uint64_t> evals((size_t)G * 40);
  for (int g = 0; g < G; g++) {
    for (int i = 0; i < 5; i++) tr[g].append_commitment(out[g]->split_quot_poly_comms[i]);
    ch[g] = tr[g].get_and_append_challenge();
    p0[g] = ch[g].v;
    o0[g] = evals.data() + (size_t)g * 40;
  }
  round4(job, p0.data(), o0.data());
  // Round 5
  for (int g = 0; g < G; g++) {
    const uint64_t* ev = evals.data() + (size_t)g * 40;
    memcpy(out[g]->wires_evals, ev, 5 * 32);
    memcpy(out[g]->wire_sigma_evals, ev + 20, 4 * 32);
    memcpy(out[g]->perm_next_eval, ev + 36, 32);
    for (int i = 0; i < 10; i++) tr[g].append_field(HFr::from_limbs(ev + 4 * i));
    ch[g] = tr[g].get_and_append_challenge();
    p0[g] = ch[g].v;
  }
  std::vector<uint64_t> open2((size_t)G * 16);
  for (int g = 0; g < G; g++) o0[g] = open2.data() + (size_t)g * 16;
  round5(job, p0.data(), o0.data());
  for (int g = 0; g < G; g++) {
    memcpy(out[g]->opening_proof, open2.data() + (size_t)g * 16, 64);
    memcpy(out[g]->shifted_opening_proof, open2.data() + (size_t)g * 16 + 8, 64);
  }
}

}  // namespace capgpu

static int prove_impl(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, bool wires_on_device, const uint64_t* pub_inputs,
                      const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  if (!ctx || !pk || !wires || !blinders || !out || (!pub_inputs && pk->num_inputs) || (!ext_msg && ext_msg_len)) return CAPGPU_ERR_ARG;
  int status = CAPGPU_OK;
  int rc = guarded(ctx, [&] {
    NoteIn in{wires, wires_on_device, pub_inputs, blinders, ext_msg, ext_msg_len};
    prove_group(ctx, pk, 1, &in, &out, &status, nullptr, 1);
  });
  return rc != CAPGPU_OK ? rc : status;
}

extern "C" int capgpu_prove(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs,
                            const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  return prove_impl(ctx, pk, wires, false, pub_inputs, blinders, ext_msg, ext_msg_len, out);
}

extern "C" int capgpu_prove_dev(capgpu_ctx* ctx, const capgpu_pk* pk, const void* d_wires, const uint64_t* pub_inputs,
                                const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  return prove_impl(ctx, pk, (const uint64_t*)d_wires, true, pub_inputs, blinders, ext_msg, ext_msg_len, out);
}

extern "C" int capgpu_ctx_set_group(capgpu_ctx* ctx, int group) {
  if (!ctx || group < 1 || group > CAPGPU_MAX_GROUP) return CAPGPU_ERR_ARG;
  ctx->group = group;
  return CAPGPU_OK;
}

// Batch of independent notes over one proving key: the B200 counterpart of the reference's rayon
// loop over builders (/root/reference/src/utils/params_builder.rs:195-233).  One worker thread per
// context pulls GROUPS of notes from a shared counter and proves each group in lockstep; two to
// four contexts per GPU keep it busy while one group waits on a host round trip.
static int prove_batch_impl(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count, const uint64_t* const* wires,
                            bool wires_on_device, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                            const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status) {
  if (!ctxs || !n_ctxs || !pk || (count && (!wires || !blinders || !out))) return CAPGPU_ERR_ARG;
  if (!pub_inputs && pk->num_inputs && count) return CAPGPU_ERR_ARG;
  for (size_t i = 0; i < n_ctxs; i++) if (!ctxs[i]) return CAPGPU_ERR_ARG;
  for (size_t i = 0; i < count; i++)
    if (!wires[i] || !blinders[i] || (pk->num_inputs && !pub_inputs[i])) return CAPGPU_ERR_ARG;
  std::mutex mu;
  size_t next = 0;
  std::atomic<int> first_error{CAPGPU_OK};
  // groups are dealt so that every context gets work: never more than an even share of what is left
  auto worker = [&](capgpu_ctx* ctx) {
    for (;;) {
      size_t i0, take;
      {
        std::lock_guard<std::mutex> lk(mu);
        if (next >= count) break;
        // an even share of what is left, but no group smaller than half the lockstep size (small groups
        // pay the five host round trips for little work)
        const size_t left = count - next, share = (left + n_ctxs - 1) / n_ctxs, G = (size_t)ctx->group;
        take = share < (G + 1) / 2 ? (G + 1) / 2 : share;
        if (take > G) take = G;
        if (take > left) take = left;
        i0 = next;
        next += take;
      }
      const int g_n = (int)take;
      std::vector<NoteIn> notes(g_n);
      std::vector<capgpu_proof*> outs(g_n);
      std::vector<int> st(g_n, CAPGPU_OK);
      for (int g = 0; g < g_n; g++) {
        const size_t i = i0 + g;
        notes[g] = NoteIn{wires[i], wires_on_device, pub_inputs ? pub_inputs[i] : nullptr, blinders[i], ext_msgs ? ext_msgs[i] : nullptr,
                          (ext_msgs && ext_msg_lens && ext_msgs[i]) ? ext_msg_lens[i] : 0};
        outs[g] = &out[i];
      }
      int rc = guarded(ctx, [&] { prove_group(ctx, pk, g_n, notes.data(), outs.data(), st.data(), nullptr, ctx->group); });
      for (int g = 0; g < g_n; g++) {
        const int code = rc != CAPGPU_OK ? rc : st[g];
        if (status) status[i0 + g] = code;
        if (code != CAPGPU_OK) {
          int expected = CAPGPU_OK;
          first_error.compare_exchange_strong(expected, code);
        }
      }
    }
  };
  std::vector<std::thread> threads;
  for (size_t t = 1; t < n_ctxs; t++) threads.emplace_back(worker, ctxs[t]);
  worker(ctxs[0]);
  for (auto& th : threads) th.join();
  return first_error.load();
}

extern "C" int capgpu_prove_batch(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count,
                                  const uint64_t* const* wires, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                                  const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status) {
  return prove_batch_impl(ctxs, n_ctxs, pk, count, wires, false, pub_inputs, blinders, ext_msgs, ext_msg_lens, out, status);
}

extern "C" int capgpu_prove_batch_dev(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count,
                                      const void* const* d_wires, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                                      const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status) {
  return prove_batch_impl(ctxs, n_ctxs, pk, count, reinterpret_cast<const uint64_t* const*>(d_wires), true, pub_inputs, blinders,
                          ext_msgs, ext_msg_lens, out, status);
}

// ---- asynchronous proving queue ----------------------------------------------------------------
// the host overlaps witness generation for note k+1 (TransferCircuit::build,
// /root/reference/src/proof/transfer.rs:167-177) with proving of note k (transfer.rs:181).
// capgpu_submit copies the note's wire values into a slot of a pinned staging ring (so the
// caller's pageable Vec<Fr> can be dropped at once) and returns a ticket; one worker thread per
// context collects up to `group` pending notes and proves them in lockstep; capgpu_poll /
// capgpu_wait deliver the proof.
struct capgpu_queue {
  std::vector<capgpu_ctx*> ctxs;
  const capgpu_pk* pk = nullptr;
  size_t n = 0, num_inputs = 0;
  struct Note {
    uint64_t ticket = 0;
    int slot = -1;
    std::vector<uint64_t> pub, blinders;
    std::vector<uint8_t> ext;
    int state = 0;  // 0 queued, 1 running, 2 done
    int status = CAPGPU_OK;
    capgpu_proof proof;
  };
  std::mutex mu;
  std::condition_variable cv_work, cv_done, cv_slot;
  std::deque<Note*> pending;
  std::unordered_map<uint64_t, Note*> notes;
  uint64_t next_ticket = 1;
  char* ring = nullptr;
  size_t slot_bytes = 0;
  std::vector<int> free_slots;
  bool stop = false;
  unsigned linger_us = 300;
  std::vector<std::thread> workers;
  // statistics
  uint64_t submitted = 0, completed = 0, groups = 0;
  double copy_ms = 0, wait_slot_ms = 0;

  void run(capgpu_ctx* ctx) {
    const size_t G = (size_t)ctx->group;
    for (;;) {
      std::vector<Note*> grp;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || !pending.empty(); });
        if (stop && pending.empty()) return;
        // linger briefly for a fuller group while notes are still arriving
        if (pending.size() < G && !stop) {
          auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(linger_us);
          cv_work.wait_until(lk, deadline, [&] { return stop || pending.size() >= G; });
        }
        // leave work for the other contexts: never more than an even share of what is pending
        const size_t share = (pending.size() + ctxs.size() - 1) / ctxs.size();
        size_t take = share < (G + 1) / 2 ? (G + 1) / 2 : share;
        if (take > G) take = G;
        if (take > pending.size()) take = pending.size();
        for (size_t i = 0; i < take; i++) { grp.push_back(pending.front()); pending.pop_front(); grp.back()->state = 1; }
        groups++;
      }
      if (grp.empty()) continue;
      const int g_n = (int)grp.size();
      std::vector<NoteIn> in(g_n);
      std::vector<capgpu_proof*> outs(g_n);
      std::vector<int> st(g_n, CAPGPU_OK);
      for (int g = 0; g < g_n; g++) {
        Note* nt = grp[g];
        in[g] = NoteIn{reinterpret_cast<const uint64_t*>(ring + (size_t)nt->slot * slot_bytes), false,
                       nt->pub.empty() ? nullptr : nt->pub.data(), nt->blinders.data(), nt->ext.empty() ? nullptr : nt->ext.data(),
                       nt->ext.size()};
        outs[g] = &nt->proof;
      }
      bool released = false;
      auto release_slots = [&] {
        if (released) return;
        released = true;
        std::lock_guard<std::mutex> lk(mu);
        for (Note* nt : grp) { free_slots.push_back(nt->slot); nt->slot = -1; }
        cv_slot.notify_all();
      };
      int rc = guarded(ctx, [&] { prove_group(ctx, pk, g_n, in.data(), outs.data(), st.data(), release_slots, ctx->group); });
      release_slots();
      {
        std::lock_guard<std::mutex> lk(mu);
        for (int g = 0; g < g_n; g++) {
          grp[g]->status = rc != CAPGPU_OK ? rc : st[g];
          grp[g]->state = 2;
          completed++;
        }
      }
      cv_done.notify_all();
    }
  }
};

extern "C" int capgpu_queue_create(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t ring_slots, capgpu_queue** out) {
  if (!ctxs || !n_ctxs || !pk || !out) return CAPGPU_ERR_ARG;
  *out = nullptr;
  for (size_t i = 0; i < n_ctxs; i++)
    if (!ctxs[i] || ctxs[i]->device != pk->device) return CAPGPU_ERR_ARG;
  capgpu_queue* q = new capgpu_queue();
  q->pk = pk;
  q->n = pk->n;
  q->num_inputs = pk->num_inputs;
  q->ctxs.assign(ctxs, ctxs + n_ctxs);
  size_t groups = 0;
  for (size_t i = 0; i < n_ctxs; i++) groups += (size_t)ctxs[i]->group;
  if (ring_slots == 0) ring_slots = 2 * groups;  // one group in flight + one being filled, per context
  q->slot_bytes = 5 * q->n * sizeof(Fr);
  if (const char* e = getenv("CAPGPU_QUEUE_LINGER_US")) q->linger_us = (unsigned)atoi(e);
  int rc = guarded(ctxs[0], [&] { CAPGPU_CUDA(cudaHostAlloc((void**)&q->ring, ring_slots * q->slot_bytes, cudaHostAllocPortable)); });
  if (rc != CAPGPU_OK) { delete q; return rc; }
  for (size_t s = ring_slots; s-- > 0;) q->free_slots.push_back((int)s);
  for (size_t i = 0; i < n_ctxs; i++) q->workers.emplace_back([q, i] { q->run(q->ctxs[i]); });
  *out = q;
  return CAPGPU_OK;
}

extern "C" void capgpu_queue_destroy(capgpu_queue* q) {
  if (!q) return;
  {
    std::lock_guard<std::mutex> lk(q->mu);
    q->stop = true;
  }
  q->cv_work.notify_all();
  for (auto& t : q->workers) t.join();
  for (auto& kv : q->notes) delete kv.second;
  if (q->ring) {
    cudaSetDevice(q->pk->device);
    cudaFreeHost(q->ring);
  }
  delete q;
}

extern "C" int capgpu_submit(capgpu_queue* q, const uint64_t* wires, const uint64_t* pub_inputs, const uint64_t* blinders,
                             const uint8_t* ext_msg, size_t ext_msg_len, uint64_t* ticket) {
  if (!q || !wires || !blinders || !ticket || (!pub_inputs && q->num_inputs) || (!ext_msg && ext_msg_len)) return CAPGPU_ERR_ARG;
  auto* nt = new capgpu_queue::Note();
  if (q->num_inputs) nt->pub.assign(pub_inputs, pub_inputs + 4 * q->num_inputs);
  nt->blinders.assign(blinders, blinders + 4 * CAPGPU_NUM_BLINDERS);
  if (ext_msg_len) nt->ext.assign(ext_msg, ext_msg + ext_msg_len);
  auto t0 = std::chrono::steady_clock::now();
  {
    std::unique_lock<std::mutex> lk(q->mu);
    if (q->stop) { delete nt; return CAPGPU_ERR_STATE; }
    q->cv_slot.wait(lk, [&] { return !q->free_slots.empty(); });  // back-pressure: the ring is full
    nt->slot = q->free_slots.back();
    q->free_slots.pop_back();
  }
  auto t1 = std::chrono::steady_clock::now();
  memcpy(q->ring + (size_t)nt->slot * q->slot_bytes, wires, q->slot_bytes);  // pageable -> pinned, on the caller's thread
  auto t2 = std::chrono::steady_clock::now();
  {
    std::lock_guard<std::mutex> lk(q->mu);
    nt->ticket = q->next_ticket++;
    q->notes[nt->ticket] = nt;
    q->pending.push_back(nt);
    q->submitted++;
    q->wait_slot_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    q->copy_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
    *ticket = nt->ticket;
  }
  q->cv_work.notify_one();
  return CAPGPU_OK;
}

extern "C" int capgpu_poll(capgpu_queue* q, uint64_t ticket, int* done) {
  if (!q || !done) return CAPGPU_ERR_ARG;
  std::lock_guard<std::mutex> lk(q->mu);
  auto it = q->notes.find(ticket);
  if (it == q->notes.end()) return CAPGPU_ERR_ARG;
  *done = it->second->state == 2;
  return CAPGPU_OK;
}

extern "C" int capgpu_wait(capgpu_queue* q, uint64_t ticket, capgpu_proof* out) {
  if (!q || !out) return CAPGPU_ERR_ARG;
  std::unique_lock<std::mutex> lk(q->mu);
  auto it = q->notes.find(ticket);
  if (it == q->notes.end()) return CAPGPU_ERR_ARG;
  capgpu_queue::Note* nt = it->second;
  q->cv_done.wait(lk, [&] { return nt->state == 2; });
  *out = nt->proof;
  int status = nt->status;
  q->notes.erase(it);
  delete nt;
  return status;
}

extern "C" int capgpu_queue_stats(capgpu_queue* q, uint64_t* submitted, uint64_t* completed, uint64_t* groups, double* copy_ms,
                                  double* wait_slot_ms) {
  if (!q) return CAPGPU_ERR_ARG;
  std::lock_guard<std::mutex> lk(q->mu);
  if (submitted) *submitted = q->submitted;
  if (completed) *completed = q->completed;
  if (groups) *groups = q->groups;
  if (copy_ms) *copy_ms = q->copy_ms;
  if (wait_slot_ms) *wait_slot_ms = q->wait_slot_ms;
  return CAPGPU_OK;
}

extern "C" int capgpu_debug_read(capgpu_ctx* ctx, int what, uint64_t* out, size_t max_elems, size_t* n_elems) {
  if (!ctx || !out || !n_elems) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    capgpu_job* job = ctx->cached_job;
    CAPGPU_REQUIRE(job != nullptr, "no proof has run on this ctx");
    const size_t n = job->n, m = job->m, NP = job->NP, G = (size_t)job->G;
    const Fr* src = nullptr;
    size_t rows = 1, len = 0, stride = 0;
    // proof slot 0 of the last group: row (r, 0) sits at r * G * NP
    switch (what) {
      case 0: src = job->polys; rows = 5; len = n + 2; stride = G * NP; break;
      case 1: src = job->z_eval; len = n; break;
      case 2: src = job->polys + 6 * G * NP; len = n + 3; break;
      case 3: throw ArgError{"quotient evaluations are overwritten in place by the coset INTT"};
      case 4: src = job->t; len = m; break;
      case 5: src = job->lin; len = n + 3; break;
      case 6: src = job->open; len = n + 3; break;
      case 7: src = job->open + G * NP; len = n + 3; break;
      case 8: src = job->polys + 5 * G * NP; len = n; break;
      case 9: src = job->split; rows = 5; len = n + 3; stride = G * NP; break;
      default: throw ArgError{"unknown debug item"};
    }
    CAPGPU_REQUIRE(rows * len <= max_elems, "debug buffer too small");
    if (rows == 1) {
      CAPGPU_CUDA(cudaMemcpyAsync(out, src, len * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      CAPGPU_CUDA(cudaMemcpy2DAsync(out, len * sizeof(Fr), src, stride * sizeof(Fr), len * sizeof(Fr), rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_elems = rows * len;
  });
}
