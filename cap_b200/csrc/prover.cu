// Device-resident 5-round TurboPlonk prover: the orchestration jf-plonk 0.1.2 performs in
// `PlonkKzgSnark::batch_prove_internal` + `Prover::{run_1st_round .. compute_opening_proofs}`
// for a single circuit, reached from /root/reference/src/proof/transfer.rs:181 (mint.rs:113,
// freeze.rs:151).  Every vector-sized step is a kernel (ntt.cu, msm.cu, poly.cu); the host
// only hashes the transcript and derives O(1) scalars between rounds, and reads back 64-byte
// commitments / 32-byte evaluations.  [UPSTREAM-RECALL: round structure per SURVEY.md App. A.]
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "poly.cuh"
#include "transcript.h"

using namespace capgpu;

static const uint64_t kHostRoot28[4] = {0x636e735580d13d9cull, 0xa22bf3742445ffd6ull, 0x56452ac01eb203d8ull, 0x1860ef942963f9e7ull};
static const uint64_t kHostGen[4] = {0x1b0d0ef99fffffe6ull, 0xeaba68a3a32a913full, 0x47d8eb76d8dd0689ull, 0x15d0085520f5bbc3ull};

static inline Fr to_dev(const HFr& x) { Fr r; memcpy(r.v, x.v, 32); return r; }
static inline HFr host_omega(unsigned log_n) {
  HFr w = HFr::from_limbs(kHostRoot28);
  for (unsigned i = 0; i < 28 - log_n; i++) w = w.sqr();
  return w;
}

struct capgpu_pk {
  int device = 0;
  unsigned log_n = 0;
  size_t n = 0, m = 0, num_inputs = 0;
  const capgpu_srs* srs = nullptr;
  capgpu_srs* lag = nullptr;  // Lagrange-basis commit key [L_0..L_{n-1}, P_0, P_1, P_n, P_{n+1}] (owned)
  bool use_lag = true;
  Fr *sel_coef = nullptr, *sig_coef = nullptr, *sig_eval = nullptr, *sel_coset = nullptr, *sig_coset = nullptr;
  Fr *xs = nullptr, *l1inv = nullptr, *zh_inv = nullptr, *omega_n = nullptr;
  HFr k[5];
  uint64_t sel_comms[13][8], sig_comms[5][8];
  std::vector<uint8_t> vk_bytes;
};

struct capgpu_job {
  capgpu_ctx* ctx = nullptr;
  const capgpu_pk* pk = nullptr;
  int round = 0;
  bool busy = false;
  size_t n = 0, m = 0, NP = 0, num_inputs = 0;
  DevBuf buf;  // one allocation, carved below
  Fr *wires_eval = nullptr, *polys = nullptr, *z_eval = nullptr, *coset = nullptr, *t = nullptr, *split = nullptr;
  Fr *lin = nullptr, *batch = nullptr, *open = nullptr, *shifted = nullptr;
  Fr *num = nullptr, *den = nullptr, *cn = nullptr, *cd = nullptr, *ntt_tmp = nullptr, *evals_dev = nullptr, *pub_dev = nullptr;
  Fr *eval_scratch = nullptr, *div_scratch = nullptr;
  size_t div_tmax = 0;
  G1Affine* comms_dev = nullptr;
  uint32_t* flag = nullptr;
  HFr beta, gamma, alpha, zeta, v;
  HFr evals[10];
};

void capgpu_job_free_internal(capgpu_job* job) {
  if (!job) return;
  job->buf.release();
  delete job;
}

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

capgpu_job* job_acquire(capgpu_ctx* ctx, const capgpu_pk* pk) {
  CAPGPU_REQUIRE(pk->device == ctx->device, "proving key lives on another device");
  capgpu_job* job = ctx->cached_job;
  if (job && job->busy) throw CodeError{CAPGPU_ERR_STATE};
  if (job && (job->n != pk->n || job->num_inputs < pk->num_inputs)) {
    capgpu_job_free_internal(job);
    ctx->cached_job = job = nullptr;
  }
  if (!job) {
    job = new capgpu_job();
    job->ctx = ctx;
    job->n = pk->n;
    job->m = pk->m;
    job->NP = pk->n + 8;
    job->num_inputs = pk->num_inputs;
    const size_t n = job->n, m = job->m, NP = job->NP;
    job->div_tmax = (NP + 15) / 16 + 1;
    size_t elems = 6 * NP + 7 * NP + n + 7 * m + m + 5 * NP + 4 * NP + 2 * n + 2 * (n / 8 + 8) + 7 * m + 16 + 160 + 2 * job->div_tmax +
                   align_up(pk->num_inputs + 1, 8);
    size_t bytes = elems * sizeof(Fr) + 8 * sizeof(G1Affine) + 256;
    job->buf.reserve(bytes);
    Fr* p = job->buf.as<Fr>();
    auto take = [&](size_t cnt) { Fr* r = p; p += cnt; return r; };
    job->wires_eval = take(6 * NP);  // rows of n evaluations, stride NP (tail: blinding scalars of the Lagrange commit)
    job->polys = take(7 * NP);
    job->z_eval = take(n);
    job->coset = take(7 * m);
    job->t = take(m);
    job->split = take(5 * NP);
    job->lin = take(NP); job->batch = take(NP); job->open = take(NP); job->shifted = take(NP);
    job->num = take(n); job->den = take(n);
    job->cn = take(n / 8 + 8); job->cd = take(n / 8 + 8);
    job->ntt_tmp = take(7 * m);
    job->evals_dev = take(16);
    job->eval_scratch = take(160);
    job->div_scratch = take(2 * job->div_tmax);
    job->pub_dev = take(align_up(pk->num_inputs + 1, 8));
    job->comms_dev = reinterpret_cast<G1Affine*>(p);
    job->flag = reinterpret_cast<uint32_t*>(job->comms_dev + 8);
    ctx->cached_job = job;
  }
  job->pk = pk;
  job->round = 0;
  job->busy = true;
  return job;
}

void job_begin(capgpu_job* job, const uint64_t* wires, const uint64_t* pub_inputs, bool wires_on_device = false) {
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n;
  CAPGPU_CUDA(cudaMemcpy2DAsync(job->wires_eval, job->NP * sizeof(Fr), wires, n * sizeof(Fr), n * sizeof(Fr), 5,
                                wires_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
  if (pk->num_inputs)
    CAPGPU_CUDA(cudaMemcpyAsync(job->pub_dev, pub_inputs, pk->num_inputs * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  fill_pi(ctx, job->wires_eval + 5 * job->NP, n, job->pub_dev, pk->num_inputs);
  job->round = 1;
}

void read_points(capgpu_job* job, int count, uint64_t* out_xy) {
  capgpu_ctx* ctx = job->ctx;
  CAPGPU_CUDA(cudaMemcpyAsync(ctx->pinned, job->comms_dev, count * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
  ctx_wait(ctx);
  memcpy(out_xy, ctx->pinned, count * sizeof(G1Affine));
}

void round1(capgpu_job* job, const uint64_t* blinders10, uint64_t* wire_comms) {
  if (job->round != 1) throw CodeError{CAPGPU_ERR_STATE};
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  // 5 wire polynomials + the public-input polynomial: one batched INTT
  ntt_device(ctx, pk->log_n, job->wires_eval, n, NP, job->polys, NP, job->ntt_tmp, 6, true, false);
  BlindArgs ba;
  memcpy(ba.b, blinders10, 10 * sizeof(Fr));
  ba.rows_blinded = 5;
  blind(ctx, job->polys, NP, n, 6, 2, ba);
  if (pk->lag && pk->use_lag) {
    // commit from the evaluations: sum_j w_j L_j + (b0 + b1 X)(X^n - 1) at tau — the same group element as
    // the coefficient-form MSM, but zero / small witness cells cost nothing / one window
    lagrange_tail(ctx, job->wires_eval, NP, n, ba);
    msm_device(ctx, pk->lag, 0, job->wires_eval, n + 4, NP, 5, true, job->comms_dev, ctx->latency_mode);
  } else {
    msm_device(ctx, pk->srs, 0, job->polys, n + 2, NP, 5, true, job->comms_dev, ctx->latency_mode);
  }
  read_points(job, 5, wire_comms);
  job->round = 2;
}

void round2(capgpu_job* job, const uint64_t* beta, const uint64_t* gamma, const uint64_t* blinders3, uint64_t* z_comm) {
  if (job->round != 2) throw CodeError{CAPGPU_ERR_STATE};
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  job->beta = HFr::from_limbs(beta);
  job->gamma = HFr::from_limbs(gamma);
  GpArgs ga;
  ga.beta = to_dev(job->beta);
  ga.gamma = to_dev(job->gamma);
  for (int i = 0; i < 5; i++) ga.k[i] = to_dev(pk->k[i]);
  grand_product(ctx, job->wires_eval, NP, pk->sig_eval, pk->omega_n, n, ga, job->num, job->den, job->cn, job->cd, job->z_eval);
  Fr* zp = job->polys + 6 * NP;
  ntt_device(ctx, pk->log_n, job->z_eval, n, n, zp, NP, job->ntt_tmp, 1, true, false);
  BlindArgs ba;
  memcpy(ba.b, blinders3, 3 * sizeof(Fr));
  ba.rows_blinded = 1;
  blind(ctx, zp, NP, n, 1, 3, ba);
  msm_device(ctx, pk->srs, 0, zp, n + 3, NP, 1, true, job->comms_dev, ctx->latency_mode);
  read_points(job, 1, z_comm);
  job->round = 3;
}

void round3(capgpu_job* job, const uint64_t* alpha, const uint64_t* blinders4, uint64_t* split_comms) {
  if (job->round != 3) throw CodeError{CAPGPU_ERR_STATE};
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, m = job->m, NP = job->NP;
  job->alpha = HFr::from_limbs(alpha);
  // coset evaluations of the 5 wire polys, PI and z on the 8n domain (selectors / sigmas are cached in the pk)
  ntt_device(ctx, pk->log_n + 3, job->polys, n + 3, NP, job->coset, m, job->ntt_tmp, 7, false, true);
  QuotArgs qa;
  qa.alpha = to_dev(job->alpha);
  qa.alpha2 = to_dev(job->alpha.sqr());
  qa.beta = to_dev(job->beta);
  qa.gamma = to_dev(job->gamma);
  for (int i = 0; i < 5; i++) qa.k[i] = to_dev(pk->k[i]);
  qa.zh_inv = pk->zh_inv;
  quotient_evals(ctx, job->coset, pk->sel_coset, pk->sig_coset, pk->xs, pk->l1inv, m, qa, job->t);
  ntt_device(ctx, pk->log_n + 3, job->t, m, m, job->t, m, job->ntt_tmp, 1, true, true);
  BlindArgs ba;
  memcpy(ba.b, blinders4, 4 * sizeof(Fr));
  ba.rows_blinded = 4;
  split_quotient(ctx, job->t, n, m, job->split, NP, ba, job->flag);
  msm_device(ctx, pk->srs, 0, job->split, n + 3, NP, 5, true, job->comms_dev, ctx->latency_mode);
  uint8_t* pin = static_cast<uint8_t*>(ctx->pinned);
  CAPGPU_CUDA(cudaMemcpyAsync(pin + 1024, job->flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  read_points(job, 5, split_comms);
  uint32_t flag;
  memcpy(&flag, pin + 1024, sizeof flag);
  if (flag) throw CodeError{CAPGPU_ERR_DEGREE};
  job->round = 4;
}

void round4(capgpu_job* job, const uint64_t* zeta, uint64_t* evals_out) {
  if (job->round != 4) throw CodeError{CAPGPU_ERR_STATE};
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  job->zeta = HFr::from_limbs(zeta);
  HFr zeta_w = job->zeta * host_omega(pk->log_n);
  EvalArgs ea;
  for (int i = 0; i < 5; i++) { ea.poly[i] = job->polys + (size_t)i * NP; ea.len[i] = n + 2; ea.x[i] = to_dev(job->zeta); }
  for (int i = 0; i < 4; i++) { ea.poly[5 + i] = pk->sig_coef + (size_t)i * n; ea.len[5 + i] = n; ea.x[5 + i] = to_dev(job->zeta); }
  ea.poly[9] = job->polys + 6 * NP; ea.len[9] = n + 3; ea.x[9] = to_dev(zeta_w);
  evaluate(ctx, ea, 10, job->evals_dev, job->eval_scratch);
  CAPGPU_CUDA(cudaMemcpyAsync(ctx->pinned, job->evals_dev, 10 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
  ctx_wait(ctx);
  memcpy(evals_out, ctx->pinned, 10 * sizeof(Fr));
  for (int i = 0; i < 10; i++) job->evals[i] = HFr::from_limbs(evals_out + 4 * i);
  job->round = 5;
}

void round5(capgpu_job* job, const uint64_t* v_in, uint64_t* opening_comms) {
  if (job->round != 5) throw CodeError{CAPGPU_ERR_STATE};
  capgpu_ctx* ctx = job->ctx;
  const capgpu_pk* pk = job->pk;
  const size_t n = job->n, NP = job->NP;
  job->v = HFr::from_limbs(v_in);
  const HFr* w = job->evals;           // wires_evals[5]
  const HFr* se = job->evals + 5;      // wire_sigma_evals[4]
  const HFr zw = job->evals[9];        // perm_next_eval
  const HFr &alpha = job->alpha, &beta = job->beta, &gamma = job->gamma, &zeta = job->zeta;
  const HFr one = HFr::one();
  HFr zh = zeta.pow_u64(n) - one;
  HFr l1 = zh * (HFr::from_u64(n) * (zeta - one)).inv();
  LinArgs la;
  la.polys = job->polys; la.split = job->split; la.sel = pk->sel_coef; la.sig = pk->sig_coef;
  la.pstride = NP; la.n = n; la.len = n + 3;
  HFr w01 = w[0] * w[1], w23 = w[2] * w[3];
  auto p5 = [](const HFr& x) { HFr x2 = x.sqr(); return x2.sqr() * x; };
  HFr sel[13] = {w[0], w[1], w[2], w[3], w01, w23, p5(w[0]), p5(w[1]), p5(w[2]), p5(w[3]), w[4].neg(), one, w01 * w23 * w[4]};
  for (int s = 0; s < 13; s++) la.cs_sel[s] = to_dev(sel[s]);
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
  HFr vp = job->v;
  for (int i = 0; i < 9; i++) { la.vp[i] = to_dev(vp); vp = vp * job->v; }
  lin_batch(ctx, la, job->lin, job->batch);
  DivArgs da;
  const HFr zeta_w = zeta * host_omega(pk->log_n);
  da.src[0] = job->batch; da.dst[0] = job->open; da.len[0] = n + 3; da.x[0] = to_dev(zeta); da.xinv[0] = to_dev(zeta.inv());
  da.src[1] = job->polys + 6 * NP; da.dst[1] = job->shifted; da.len[1] = n + 3; da.x[1] = to_dev(zeta_w); da.xinv[1] = to_dev(zeta_w.inv());
  divide_linear(ctx, da, 2, job->div_scratch, job->div_tmax);
  // open and shifted are adjacent rows of stride NP
  msm_device(ctx, pk->srs, 0, job->open, n + 2, NP, 2, true, job->comms_dev, ctx->latency_mode);
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
  CAPGPU_CUDA(cudaMalloc(&pk->omega_n, n * sizeof(Fr)));
  ctx->ntt_tmp.reserve(5 * m * sizeof(Fr));
  Fr* tmp = ctx->ntt_tmp.as<Fr>();
  ntt_device(ctx, pk->log_n, pk->sig_coef, n, n, pk->sig_eval, n, tmp, 5, false, false);
  for (int s = 0; s < 13; s++)
    ntt_device(ctx, pk->log_n + 3, pk->sel_coef + (size_t)s * n, n, n, pk->sel_coset + (size_t)s * m, m, tmp, 1, false, true);
  ntt_device(ctx, pk->log_n + 3, pk->sig_coef, n, n, pk->sig_coset, m, tmp, 5, false, true);
  CAPGPU_CUDA(cudaMemcpyAsync(pk->omega_n, domain_omega_powers(ctx, pk->log_n), n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
  {
    const char* e = getenv("CAPGPU_LAGRANGE");
    if (!(e && atoi(e) == 0) && pk->srs->n >= n + 2)
      pk->lag = srs_lagrange(ctx, pk->srs, pk->log_n, pk->omega_n, to_dev(HFr::from_u64(n).inv()));
  }
  HFr gen = HFr::from_limbs(kHostGen);
  coset_tables(ctx, domain_omega_powers(ctx, pk->log_n + 3), m, to_dev(gen), to_dev(HFr::from_u64(n)), pk->xs, pk->l1inv);
  HFr wm = host_omega(pk->log_n + 3);
  Fr zh[8];
  HFr x = gen;
  for (int i = 0; i < 8; i++) {
    zh[i] = to_dev((x.pow_u64(n) - HFr::one()).inv());
    x = x * wm;
  }
  CAPGPU_CUDA(cudaMemcpyAsync(pk->zh_inv, zh, sizeof zh, cudaMemcpyHostToDevice, ctx->stream));
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
  pk->m = pk->n * 8;
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
  delete pk;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------
extern "C" int capgpu_pk_upload(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* selectors,
                                const uint64_t* sigmas, const uint64_t* k, const uint64_t* selector_comms_xy,
                                const uint64_t* sigma_comms_xy, capgpu_pk** out) {
  if (!ctx || !srs || !selectors || !sigmas || !k || !selector_comms_xy || !sigma_comms_xy || !out) return CAPGPU_ERR_ARG;
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
    for (int s = 0; s < 13; s += 5) {
      int cnt = 13 - s < 5 ? 13 - s : 5;
      msm_device(ctx, srs, 0, pk->sel_coef + (size_t)s * n, n, n, cnt, true, outp + s);
    }
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
    capgpu_job* job = job_acquire(ctx, pk);
    try {
      job_begin(job, wires, pub_inputs);
    } catch (...) {
      job->busy = false;
      throw;
    }
    *out = job;
  });
}

#define CAPGPU_JOB_GUARD(job, body)                    \
  if (!(job) || !(job)->busy) return CAPGPU_ERR_STATE; \
  return guarded((job)->ctx, [&] { body; });

extern "C" int capgpu_job_round1(capgpu_job* job, const uint64_t* blinders10, uint64_t* wire_comms_xy) {
  if (!blinders10 || !wire_comms_xy) return CAPGPU_ERR_ARG;
  CAPGPU_JOB_GUARD(job, round1(job, blinders10, wire_comms_xy));
}
extern "C" int capgpu_job_round2(capgpu_job* job, const uint64_t* beta, const uint64_t* gamma, const uint64_t* blinders3, uint64_t* z_comm_xy) {
  if (!beta || !gamma || !blinders3 || !z_comm_xy) return CAPGPU_ERR_ARG;
  CAPGPU_JOB_GUARD(job, round2(job, beta, gamma, blinders3, z_comm_xy));
}
extern "C" int capgpu_job_round3(capgpu_job* job, const uint64_t* alpha, const uint64_t* blinders4, uint64_t* split_comms_xy) {
  if (!alpha || !blinders4 || !split_comms_xy) return CAPGPU_ERR_ARG;
  CAPGPU_JOB_GUARD(job, round3(job, alpha, blinders4, split_comms_xy));
}
extern "C" int capgpu_job_round4(capgpu_job* job, const uint64_t* zeta, uint64_t* evals) {
  if (!zeta || !evals) return CAPGPU_ERR_ARG;
  CAPGPU_JOB_GUARD(job, round4(job, zeta, evals));
}
extern "C" int capgpu_job_round5(capgpu_job* job, const uint64_t* v, uint64_t* opening_comms_xy) {
  if (!v || !opening_comms_xy) return CAPGPU_ERR_ARG;
  CAPGPU_JOB_GUARD(job, round5(job, v, opening_comms_xy));
}
extern "C" void capgpu_job_end(capgpu_job* job) {
  if (job) job->busy = false;
}

static int prove_impl(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, bool wires_on_device, const uint64_t* pub_inputs,
                      const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out);

extern "C" int capgpu_prove(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, const uint64_t* pub_inputs,
                            const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  return prove_impl(ctx, pk, wires, false, pub_inputs, blinders, ext_msg, ext_msg_len, out);
}

extern "C" int capgpu_prove_dev(capgpu_ctx* ctx, const capgpu_pk* pk, const void* d_wires, const uint64_t* pub_inputs,
                                const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  return prove_impl(ctx, pk, (const uint64_t*)d_wires, true, pub_inputs, blinders, ext_msg, ext_msg_len, out);
}

static int prove_impl(capgpu_ctx* ctx, const capgpu_pk* pk, const uint64_t* wires, bool wires_on_device, const uint64_t* pub_inputs,
                      const uint64_t* blinders, const uint8_t* ext_msg, size_t ext_msg_len, capgpu_proof* out) {
  if (!ctx || !pk || !wires || !blinders || !out || (!pub_inputs && pk->num_inputs) || (!ext_msg && ext_msg_len)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    capgpu_job* job = job_acquire(ctx, pk);
    struct Release { capgpu_job* j; ~Release() { j->busy = false; } } release{job};
    job_begin(job, wires, pub_inputs, wires_on_device);
    SolidityTranscript tr;
    if (ext_msg_len) tr.append_message(ext_msg, ext_msg_len);
    tr.append_message(pk->vk_bytes.data(), pk->vk_bytes.size());
    for (size_t i = 0; i < pk->num_inputs; i++) tr.append_field(HFr::from_limbs(pub_inputs + 4 * i));
    // Round 1
    round1(job, blinders, &out->wires_poly_comms[0][0]);
    for (int i = 0; i < 5; i++) tr.append_commitment(out->wires_poly_comms[i]);
    // Round 2
    HFr beta = tr.get_and_append_challenge();
    HFr gamma = tr.get_and_append_challenge();
    round2(job, beta.v, gamma.v, blinders + 4 * 10, out->prod_perm_poly_comm);
    tr.append_commitment(out->prod_perm_poly_comm);
    // Round 3
    HFr alpha = tr.get_and_append_challenge();
    round3(job, alpha.v, blinders + 4 * 13, &out->split_quot_poly_comms[0][0]);
    for (int i = 0; i < 5; i++) tr.append_commitment(out->split_quot_poly_comms[i]);
    // Round 4
    HFr zeta = tr.get_and_append_challenge();
    uint64_t evals[40];
    round4(job, zeta.v, evals);
    memcpy(out->wires_evals, evals, 5 * 32);
    memcpy(out->wire_sigma_evals, evals + 20, 4 * 32);
    memcpy(out->perm_next_eval, evals + 36, 32);
    for (int i = 0; i < 10; i++) tr.append_field(HFr::from_limbs(evals + 4 * i));
    // Round 5
    HFr v = tr.get_and_append_challenge();
    uint64_t open2[16];
    round5(job, v.v, open2);
    memcpy(out->opening_proof, open2, 64);
    memcpy(out->shifted_opening_proof, open2 + 8, 64);
  });
}

// Batch of independent notes over one proving key: the B200 counterpart of the reference's rayon
// loop over builders (/root/reference/src/utils/params_builder.rs:195-233).  One worker thread per
// context pulls note indices from a shared counter; each context's stream carries one proof at a
// time, so latency-bound kernels of one proof overlap with throughput-bound kernels of another.
extern "C" int capgpu_prove_batch(capgpu_ctx* const* ctxs, size_t n_ctxs, const capgpu_pk* pk, size_t count,
                                  const uint64_t* const* wires, const uint64_t* const* pub_inputs, const uint64_t* const* blinders,
                                  const uint8_t* const* ext_msgs, const size_t* ext_msg_lens, capgpu_proof* out, int* status) {
  if (!ctxs || !n_ctxs || !pk || (count && (!wires || !blinders || !out))) return CAPGPU_ERR_ARG;
  for (size_t i = 0; i < n_ctxs; i++) if (!ctxs[i]) return CAPGPU_ERR_ARG;
  std::atomic<size_t> next{0};
  std::atomic<int> first_error{CAPGPU_OK};
  auto worker = [&](capgpu_ctx* ctx) {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= count) break;
      const uint64_t* pi = pub_inputs ? pub_inputs[i] : nullptr;
      const uint8_t* msg = ext_msgs ? ext_msgs[i] : nullptr;
      size_t len = (ext_msgs && ext_msg_lens) ? ext_msg_lens[i] : 0;
      int rc = capgpu_prove(ctx, pk, wires[i], pi, blinders[i], msg, len, &out[i]);
      if (status) status[i] = rc;
      if (rc != CAPGPU_OK) {
        int expected = CAPGPU_OK;
        first_error.compare_exchange_strong(expected, rc);
      }
    }
  };
  std::vector<std::thread> threads;
  for (size_t t = 1; t < n_ctxs; t++) threads.emplace_back(worker, ctxs[t]);
  worker(ctxs[0]);
  for (auto& th : threads) th.join();
  return first_error.load();
}

extern "C" int capgpu_debug_read(capgpu_ctx* ctx, int what, uint64_t* out, size_t max_elems, size_t* n_elems) {
  if (!ctx || !out || !n_elems) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    capgpu_job* job = ctx->cached_job;
    CAPGPU_REQUIRE(job != nullptr, "no proof has run on this ctx");
    const size_t n = job->n, m = job->m, NP = job->NP;
    const Fr* src = nullptr;
    size_t rows = 1, len = 0, stride = 0;
    switch (what) {
      case 0: src = job->polys; rows = 5; len = n + 2; stride = NP; break;
      case 1: src = job->z_eval; len = n; break;
      case 2: src = job->polys + 6 * NP; len = n + 3; break;
      case 3: throw ArgError{"quotient evaluations are overwritten in place by the coset INTT"};
      case 4: src = job->t; len = m; break;
      case 5: src = job->lin; len = n + 3; break;
      case 6: src = job->open; len = n + 3; break;
      case 7: src = job->shifted; len = n + 3; break;
      case 8: src = job->polys + 5 * NP; len = n; break;
      case 9: src = job->split; rows = 5; len = n + 3; stride = NP; break;
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
