// Polynomial-side kernels of the TurboPlonk prover (everything between the NTTs and MSMs):
// blinding, permutation grand product with batch inversion, point-wise quotient evaluation on
// the quotient domain (8n coset or 3 x 2n cosets), quotient splitting, Horner evaluation, linearisation / batching, division by
// (X - z).  Each replaces a CPU loop of jf-plonk 0.1.2 `Prover` / jf-relation 0.1.2
// `PlonkCircuit` behind /root/reference/src/proof/transfer.rs:181 (names cited per kernel).
#include "poly.cuh"
#include <stdlib.h>

namespace capgpu {

// ------------------------------------------------------------------------------------------
// Prover::mask_polynomial: p(X) + (b0 + b1 X + ..)(X^n - 1); also clears the padding tail.
// polys: rows of `stride` elements holding n coefficients; row r gets blinders[boff[r] ..].
// ------------------------------------------------------------------------------------------
__global__ void blind_kernel(Fr* polys, size_t stride, size_t n, int nrows, int G, int nb, const BlindArgs* __restrict__ argv) {
  const int r = blockIdx.x, g = blockIdx.y;
  if (r >= nrows) return;
  const BlindArgs& args = argv[g];
  Fr* p = polys + ((size_t)r * G + g) * stride;
  for (size_t t = threadIdx.x; t < stride - n; t += blockDim.x) {
    Fr v = Fr::zero();
    if ((int)t < nb && r < args.rows_blinded) {
      Fr b = args.b[r * nb + t];
      v = b;
      p[t] = fp_sub(p[t], b);
    }
    p[n + t] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Permutation grand product (jf-relation compute_prod_permutation_polynomial):
//   z_0 = 1, z_{j+1} = z_j * prod_i (w_ij + beta*k_i*omega^j + gamma) / prod_i (w_ij + beta*sigma_ij + gamma)
// Serial with one field division per row in the reference; here: per-row numerator and
// denominator (gp_terms), chunk products, one block-wide scan over <= 1024 chunk products with
// a single inversion (gp_scan), then each chunk rebuilds its prefix numerators forward and the
// inverse prefix denominators backward (gp_finish).  15 products per row, 1 inversion total.
// ------------------------------------------------------------------------------------------
__global__ void gp_terms(const Fr* __restrict__ wires, size_t wstride, const Fr* __restrict__ sig_eval, const Fr* __restrict__ omega_pows,
                         size_t n, int G, const GpArgs* __restrict__ argv, Fr* num, Fr* den) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int g = blockIdx.y;
  const GpArgs& a = argv[g];
  wires += (size_t)g * wstride;
  wstride *= G;  // row (i, g) of the [5][G] block
  num += (size_t)g * n;
  den += (size_t)g * n;
  Fr bw = fp_mul(a.beta, omega_pows[j]);
  Fr nu, de;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    Fr w = fp_add(wires[(size_t)i * wstride + j], a.gamma);
    Fr id = i == 0 ? bw : fp_mul(bw, a.k[i]);
    Fr tn = fp_add(w, id);
    Fr td = fp_add(w, fp_mul(a.beta, sig_eval[(size_t)i * n + j]));
    nu = i == 0 ? tn : fp_mul(nu, tn);
    de = i == 0 ? td : fp_mul(de, td);
  }
  num[j] = nu;
  den[j] = de;
}

__global__ void gp_chunk_prod(const Fr* __restrict__ num, const Fr* __restrict__ den, size_t n, size_t L, Fr* cn, Fr* cd, size_t cstride) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = n / L;
  if (t >= T) return;
  num += (size_t)blockIdx.y * n; den += (size_t)blockIdx.y * n;
  cn += (size_t)blockIdx.y * cstride; cd += (size_t)blockIdx.y * cstride;
  Fr pn = num[t * L], pd = den[t * L];
  for (size_t j = t * L + 1; j < (t + 1) * L; j++) {
    pn = fp_mul(pn, num[j]);
    pd = fp_mul(pd, den[j]);
  }
  cn[t] = pn;
  cd[t] = pd;
}

// One block of 1024 threads over T chunk products (each thread owns `per` consecutive chunks).
// Out: cn[t] = prod_{u<t} cn_in[u] (numerator prefix at the chunk start),
//      cd[t] = 1 / prod_{u<=t} cd_in[u] (inverse denominator prefix at the chunk end).
__global__ void __launch_bounds__(1024) gp_scan(Fr* cn, Fr* cd, int T, size_t cstride) {
  extern __shared__ uint32_t gp_sm[];
  cn += (size_t)blockIdx.x * cstride; cd += (size_t)blockIdx.x * cstride;
  Fr* sn = reinterpret_cast<Fr*>(gp_sm);  // 1024 entries
  Fr* sd = sn + 1024;                     // 1024 entries
  __shared__ Fr inv_total;
  const int t = threadIdx.x;
  const int per = (T + 1023) / 1024;
  const int lo = t * per, hi = lo + per < T ? lo + per : T;
  Fr pn = Fr::one(), pd = Fr::one();
  for (int i = lo; i < hi; i++) { pn = fp_mul(pn, cn[i]); pd = fp_mul(pd, cd[i]); }
  sn[t] = pn;
  sd[t] = pd;
  __syncthreads();
  // inclusive prefix product of sn, inclusive suffix product of sd (Hillis-Steele)
  for (int o = 1; o < 1024; o <<= 1) {
    Fr a, b;
    bool ha = t >= o, hb = t + o < 1024;
    if (ha) a = sn[t - o];
    if (hb) b = sd[t + o];
    __syncthreads();
    if (ha) sn[t] = fp_mul(sn[t], a);
    if (hb) sd[t] = fp_mul(sd[t], b);
    __syncthreads();
  }
  if (t == 0) inv_total = fp_inv(sd[0]);
  __syncthreads();
  // numerators: forward from the exclusive prefix of this thread
  Fr run = t == 0 ? Fr::one() : sn[t - 1];
  for (int i = lo; i < hi; i++) {
    Fr cur = cn[i];
    cn[i] = run;
    run = fp_mul(run, cur);
  }
  // denominators: backward from inv_total * (product of everything right of this thread)
  Fr rd = t + 1 < 1024 ? fp_mul(inv_total, sd[t + 1]) : inv_total;
  for (int i = hi; i-- > lo;) {
    Fr cur = cd[i];
    cd[i] = rd;  // 1 / prod_{u<=i} d_u = inv_total * prod_{u>i} d_u
    rd = fp_mul(rd, cur);
  }
}

__global__ void gp_finish(const Fr* __restrict__ num, const Fr* __restrict__ den, size_t n, size_t L, const Fr* __restrict__ cn,
                          const Fr* __restrict__ cd, size_t cstride, Fr* z) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t T = n / L;
  if (t >= T) return;
  num += (size_t)blockIdx.y * n; den += (size_t)blockIdx.y * n; z += (size_t)blockIdx.y * n;
  cn += (size_t)blockIdx.y * cstride; cd += (size_t)blockIdx.y * cstride;
  Fr pn = cn[t];
  for (size_t j = t * L; j < (t + 1) * L; j++) {
    z[j] = pn;  // N_j = prod_{i<j} num_i
    pn = fp_mul(pn, num[j]);
  }
  Fr id = cd[t];  // 1 / D_{(t+1)L}
  for (size_t j = (t + 1) * L; j-- > t * L;) {
    id = fp_mul(id, den[j]);  // 1 / D_j
    z[j] = fp_mul(z[j], id);
  }
}

void grand_product(capgpu_ctx* ctx, const Fr* wires, size_t wstride, const Fr* sig_eval, const Fr* omega_pows, size_t n, int G,
                   const GpArgs* args, Fr* num, Fr* den, Fr* cn, Fr* cd, size_t cstride, Fr* z) {
  // chunks of 8 rows (fewer for tiny domains): T = n / L chunk products, scanned by one CTA per proof
  size_t L = n >= 16 ? 8 : 1;
  size_t T = n / L;
  ProfScope prof(ctx, PROF_GRAND_PRODUCT, (double)n * G);
  gp_terms<<<dim3(ceil_div(n, 128), G), 128, 0, ctx->stream>>>(wires, wstride, sig_eval, omega_pows, n, G, args, num, den);
  CAPGPU_LAUNCH_CHECK(ctx);
  gp_chunk_prod<<<dim3(ceil_div(T, 128), G), 128, 0, ctx->stream>>>(num, den, n, L, cn, cd, cstride);
  CAPGPU_LAUNCH_CHECK(ctx);
  CAPGPU_CUDA(cudaFuncSetAttribute(gp_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 1024 * (int)sizeof(Fr)));
  gp_scan<<<G, 1024, 2 * 1024 * sizeof(Fr), ctx->stream>>>(cn, cd, (int)T, cstride);
  CAPGPU_LAUNCH_CHECK(ctx);
  gp_finish<<<dim3(ceil_div(T, 128), G), 128, 0, ctx->stream>>>(num, den, n, L, cn, cd, cstride, z);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------
// Quotient evaluation on the quotient domain (QuotDomain, poly.cuh; jf-plonk Prover::compute_quotient_polynomial's point-wise
// closure: compute_quotient_circuit_contribution + compute_quotient_copy_constraint_contribution):
//   t(x) = (t_circ + alpha*[z(x) prod(w_i + beta k_i x + gamma) - z(wx) prod(w_i + beta sigma_i + gamma)]) / Z_H(x)
//          + alpha^2 (z(x) - 1) / (n (x - 1))
// One thread per coset point; 26 coalesced 32-byte streams in, one out; ~52 field products.
// The witness-independent factors 1/Z_H (8 values) and 1/(n(x-1)) come from pk tables.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Fr pow5(const Fr& w) {
  Fr w2 = fp_sqr(w);
  return fp_mul(fp_sqr(w2), w);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) quotient_kernel(const Fr* __restrict__ coset /*7 x m: w0..w4, pi, z*/, const Fr* __restrict__ sel /*13 x m*/,
                                                       const Fr* __restrict__ sig /*5 x m*/, const Fr* __restrict__ xs /*m*/,
                                                       const Fr* __restrict__ l1inv /*m*/, const Fr* __restrict__ zh_inv /*cosets x step*/, QuotDomain qd,
                                                       int G, const QuotArgs* __restrict__ argv, Fr* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t m = qd.m;
  if (i >= m) return;
  const int g = blockIdx.y;
  const QuotArgs& a = argv[g];
  coset += (size_t)g * m;          // row (j, g) of the [7][G] block is coset[j * cm + ..]
  const size_t cm = (size_t)G * m;
  out += (size_t)g * m;
  Fr w[5];
#pragma unroll
  for (int j = 0; j < 5; j++) w[j] = coset[(size_t)j * cm + i];
  // gate part
  // the twelve selector products are summed two at a time with one Montgomery reduction per pair
  Fr acc = fp_add(sel[11 * m + i], coset[5 * cm + i]);  // q_c + PI
  acc = fp_add(acc, fp_mul_add(sel[0 * m + i], w[0], sel[1 * m + i], w[1]));
  acc = fp_add(acc, fp_mul_add(sel[2 * m + i], w[2], sel[3 * m + i], w[3]));
  Fr w01 = fp_mul(w[0], w[1]);
  Fr w23 = fp_mul(w[2], w[3]);
  acc = fp_add(acc, fp_mul_add(sel[4 * m + i], w01, sel[5 * m + i], w23));
  acc = fp_add(acc, fp_mul_add(sel[6 * m + i], pow5(w[0]), sel[7 * m + i], pow5(w[1])));
  acc = fp_add(acc, fp_mul_add(sel[8 * m + i], pow5(w[2]), sel[9 * m + i], pow5(w[3])));
  acc = fp_add(acc, fp_mul_sub(sel[12 * m + i], fp_mul(fp_mul(w01, w23), w[4]), sel[10 * m + i], w[4]));
  // permutation part
  Fr z = coset[6 * cm + i];
  const size_t isub = i & (qd.sub - 1);
  size_t inext = i + qd.step;  // w_n x, inside the same coset
  if (isub + qd.step >= qd.sub) inext -= qd.sub;
  Fr zn = coset[6 * cm + inext];
  Fr bx = fp_mul(a.beta, xs[i]);
  Fr r1 = z, r2 = zn;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    Fr wg = fp_add(w[j], a.gamma);
    Fr id = j == 0 ? bx : fp_mul(bx, a.k[j]);
    r1 = fp_mul(r1, fp_add(wg, id));
    r2 = fp_mul(r2, fp_add(wg, fp_mul(a.beta, sig[(size_t)j * m + i])));
  }
  acc = fp_add(acc, fp_mul(a.alpha, fp_sub(r1, r2)));
  // acc / Z_H(x) + alpha^2 (z - 1) / (n (x - 1)); zh_inv: device table 1 / ((s_k w_sub^i)^n - 1), period `step` in i
  out[i] = fp_mul_add(acc, zh_inv[(i >> qd.log_sub) * qd.step + (isub & (qd.step - 1))], fp_mul(a.alpha2, fp_sub(z, Fr::one())), l1inv[i]);
}

void quotient_evals(capgpu_ctx* ctx, const Fr* coset, const Fr* sel, const Fr* sig, const Fr* xs, const Fr* l1inv, const Fr* zh_inv,
                    const QuotDomain& qd, int G, const QuotArgs* args, Fr* out) {
  const size_t m = qd.m;
  ProfScope prof(ctx, PROF_QUOTIENT, (double)m * G);
  // resident CTAs per SM (register cap): 2 = 176 registers (ncu: 2 warps per scheduler, multiplier busy 65 %; 507 proofs/s), 3 / 4 (525 / 530 proofs/s) trade
  // a few spills for occupancy; CAPGPU_QUOT_MINB selects for A/B runs
  static const int minb = [] { const char* e = getenv("CAPGPU_QUOT_MINB"); int v = e ? atoi(e) : 4; return v < 2 ? 2 : (v > 4 ? 4 : v); }();
  const dim3 grid(ceil_div(m, 128), G);
  if (minb == 2) quotient_kernel<2><<<grid, 128, 0, ctx->stream>>>(coset, sel, sig, xs, l1inv, zh_inv, qd, G, args, out);
  else if (minb == 3) quotient_kernel<3><<<grid, 128, 0, ctx->stream>>>(coset, sel, sig, xs, l1inv, zh_inv, qd, G, args, out);
  else quotient_kernel<4><<<grid, 128, 0, ctx->stream>>>(coset, sel, sig, xs, l1inv, zh_inv, qd, G, args, out);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// pk tables: xs[k * sub + i] = s_k * w_sub^i, l1inv = 1 / (n * (xs - 1))
__global__ void coset_tables_kernel(const Fr* __restrict__ omega_sub, QuotDomain qd, CosetShifts shifts, Fr n_mont, Fr* xs, Fr* l1inv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= qd.m) return;
  const uint32_t k = (uint32_t)(i >> qd.log_sub);
  const Fr s = k == 0 ? shifts.s[0] : (k == 1 ? shifts.s[1] : shifts.s[2]);
  Fr x = fp_mul(s, omega_sub[i & (qd.sub - 1)]);
  xs[i] = x;
  l1inv[i] = fp_inv(fp_mul(n_mont, fp_sub(x, Fr::one())));
}

void coset_tables(capgpu_ctx* ctx, const Fr* omega_sub, const QuotDomain& qd, const CosetShifts& shifts, const Fr& n_mont, Fr* xs, Fr* l1inv) {
  coset_tables_kernel<<<ceil_div(qd.m, 128), 128, 0, ctx->stream>>>(omega_sub, qd, shifts, n_mont, xs, l1inv);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------
// Degree check of the quotient (jf-plonk split_quotient_polynomial's WrongQuotientPolyDegree)
// and the split into 5 chunks of n+2 coefficients with the zero-knowledge maskers:
//   t_i(X) + b_i X^(n+2) - b_{i-1}
// flag[0] |= 1 if any coefficient above 5n+7 is non-zero, |= 2 if coefficient 5n+7 is zero.
// ------------------------------------------------------------------------------------------
__global__ void split_kernel(const Fr* __restrict__ t, size_t n, size_t m, int G, Fr* split, size_t stride,
                             const BlindArgs* __restrict__ argv, uint32_t* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t deg = 5 * n + 7;
  const int g = blockIdx.y;
  const BlindArgs& args = argv[g];
  t += (size_t)g * m;
  flag += g;
  split += (size_t)g * stride;  // row (r, g) at split[r * rs + ..]
  const size_t rs = (size_t)G * stride;
  if (i < m) {
    Fr c = t[i];
    if (i > deg && !c.is_zero()) atomicOr(flag, 1u);
    if (i == deg && c.is_zero()) atomicOr(flag, 2u);
    if (i <= deg) {
      size_t r = i / (n + 2);
      size_t o = i - r * (n + 2);
      if (r > 4) { r = 4; o = i - 4 * (n + 2); }
      if (o == 0 && r > 0) c = fp_sub(c, args.b[r - 1]);
      split[r * rs + o] = c;
    }
  }
  // tails: row r < 4 gets b_r at position n+2, zeros after; row 4 is zero from n on
  if (i < 5 * (stride - n)) {
    size_t r = i / (stride - n), o = n + i % (stride - n);
    if (r < 4) {
      if (o == n + 2) split[r * rs + o] = args.b[r];
      else if (o > n + 2) split[r * rs + o] = Fr::zero();
    } else {
      split[r * rs + o] = Fr::zero();
    }
  }
}

void split_quotient(capgpu_ctx* ctx, const Fr* t, size_t n, size_t m, int G, Fr* split, size_t stride, const BlindArgs* args, uint32_t* flag) {
  CAPGPU_CUDA(cudaMemsetAsync(flag, 0, G * sizeof(uint32_t), ctx->stream));
  split_kernel<<<dim3(ceil_div(m, 256), G), 256, 0, ctx->stream>>>(t, n, m, G, split, stride, args, flag);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------
// Horner evaluation (Prover::compute_evaluations): EVAL_SPLIT CTAs per (polynomial, point), each
// evaluating one segment (thread-local Horner over a short chunk, scaled by x^start, CTA tree);
// a second tiny kernel adds the segment sums.
// ------------------------------------------------------------------------------------------
constexpr int EVAL_SPLIT = 16;

__global__ void __launch_bounds__(128) eval_kernel(const EvalArgs* __restrict__ argv, Fr* partials) {
  __shared__ Fr sm[128];
  const int b = blockIdx.y;
  const EvalArgs& a = argv[blockIdx.z];
  partials += (size_t)blockIdx.z * 160;
  const Fr* p = a.poly[b];
  const size_t len = a.len[b];
  const Fr x = a.x[b];
  const size_t seg = (len + EVAL_SPLIT - 1) / EVAL_SPLIT;
  const size_t seg_start = blockIdx.x * seg;
  const size_t seg_end = seg_start + seg < len ? seg_start + seg : len;
  const size_t chunk = (seg + blockDim.x - 1) / blockDim.x;
  const size_t start = seg_start + threadIdx.x * chunk;
  Fr acc = Fr::zero();
  if (start < seg_end) {
    size_t end = start + chunk < seg_end ? start + chunk : seg_end;
    acc = p[end - 1];
    for (size_t j = end - 1; j-- > start;) acc = fp_add(fp_mul(acc, x), p[j]);
    acc = fp_mul(acc, fp_pow_u64(x, start));
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sm[threadIdx.x] = fp_add(sm[threadIdx.x], sm[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[b * EVAL_SPLIT + blockIdx.x] = sm[0];
}

__global__ void eval_sum_kernel(const Fr* __restrict__ partials, int count, Fr* out) {
  int b = threadIdx.x;
  if (b >= count) return;
  partials += (size_t)blockIdx.x * 160;
  out += (size_t)blockIdx.x * 16;
  Fr acc = partials[b * EVAL_SPLIT];
  for (int i = 1; i < EVAL_SPLIT; i++) acc = fp_add(acc, partials[b * EVAL_SPLIT + i]);
  out[b] = acc;
}

void evaluate(capgpu_ctx* ctx, const EvalArgs* args, int count, int G, Fr* out, Fr* scratch) {
  eval_kernel<<<dim3(EVAL_SPLIT, count, G), 128, 0, ctx->stream>>>(args, scratch);
  CAPGPU_LAUNCH_CHECK(ctx);
  eval_sum_kernel<<<G, 32, 0, ctx->stream>>>(scratch, count, out);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------
// Linearisation polynomial and the batched opening polynomial
// (Prover::compute_{non_,}quotient_component_for_lin_poly, compute_opening_proofs):
//   lin   = sum_s a_s q_s + cz z + cs sigma_4 + sum_i ct_i t_i
//   batch = lin + sum_{i<5} v^(i+1) w_i + sum_{i<4} v^(6+i) sigma_i
// ------------------------------------------------------------------------------------------
__global__ void lin_batch_kernel(const LinArgs* __restrict__ argv, const Fr* __restrict__ sel, const Fr* __restrict__ sig, size_t n,
                                 size_t len, Fr* lin, Fr* batch, size_t ostride) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  const LinArgs& a = argv[blockIdx.y];
  lin += (size_t)blockIdx.y * ostride;
  batch += (size_t)blockIdx.y * ostride;
  Fr acc = fp_mul(a.polys[6 * a.rstride + j], a.cz);  // z poly (row 6), n+3 coefficients
  if (j < n) {
#pragma unroll
    for (int s = 0; s < 13; s++) acc = fp_add(acc, fp_mul(sel[(size_t)s * n + j], a.cs_sel[s]));
    acc = fp_add(acc, fp_mul(sig[4 * n + j], a.csig));
  }
#pragma unroll
  for (int i = 0; i < 5; i++) acc = fp_add(acc, fp_mul(a.split[(size_t)i * a.rstride + j], a.ct[i]));
  lin[j] = acc;
  Fr bt = acc;
#pragma unroll
  for (int i = 0; i < 5; i++) bt = fp_add(bt, fp_mul(a.polys[(size_t)i * a.rstride + j], a.vp[i]));
  if (j < n) {
#pragma unroll
    for (int i = 0; i < 4; i++) bt = fp_add(bt, fp_mul(sig[(size_t)i * n + j], a.vp[5 + i]));
  }
  batch[j] = bt;
}

void lin_batch(capgpu_ctx* ctx, const LinArgs* args, const Fr* sel, const Fr* sig, size_t n, size_t len, int G, Fr* lin, Fr* batch,
               size_t ostride) {
  lin_batch_kernel<<<dim3(ceil_div(len, 128), G), 128, 0, ctx->stream>>>(args, sel, sig, n, len, lin, batch, ostride);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------
// Division by (X - x) (the reference uses DensePolynomial long division, a serial recurrence
// q_{i-1} = p_i + x q_i).  Closed form: q_i = x^-(i+1) * sum_{j>i} p_j x^j, i.e. an ADDITIVE
// suffix sum of a_j = p_j x^j.  Three grid-wide kernels: per-chunk sums of a_j, a suffix scan of
// the chunk sums (additions only), and the per-chunk replay.  x^-1 comes from the host.
// ------------------------------------------------------------------------------------------
constexpr int DIV_CHUNK = 16;

__global__ void __launch_bounds__(128) div_chunk_sums(const DivArgs* __restrict__ argv, int count, Fr* totals, size_t tmax) {
  const DivArgs& a = argv[blockIdx.y / count];
  const int b = blockIdx.y % count;
  totals += (size_t)(blockIdx.y - b) * tmax;
  const size_t len = a.len[b];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t T = (len + DIV_CHUNK - 1) / DIV_CHUNK;
  if (t >= T) return;
  const Fr* p = a.src[b];
  const Fr x = a.x[b];
  const size_t start = t * DIV_CHUNK, end = start + DIV_CHUNK < len ? start + DIV_CHUNK : len;
  Fr pw = fp_pow_u64(x, start);
  Fr sum = Fr::zero();
  for (size_t j = start; j < end; j++) {
    sum = fp_add(sum, fp_mul(p[j], pw));
    pw = fp_mul(pw, x);
  }
  totals[b * tmax + t] = sum;
}

// totals[t] <- sum_{u > t} totals[u]   (one CTA per polynomial)
__global__ void __launch_bounds__(1024) div_scan(const DivArgs* __restrict__ argv, int count, Fr* totals, size_t tmax) {
  __shared__ Fr sm[1024];
  const DivArgs& a = argv[blockIdx.x / count];
  const int b = blockIdx.x % count;
  totals += (size_t)(blockIdx.x - b) * tmax;
  const size_t T = (a.len[b] + DIV_CHUNK - 1) / DIV_CHUNK;
  Fr* v = totals + b * tmax;
  const size_t per = (T + blockDim.x - 1) / blockDim.x;
  const size_t lo = threadIdx.x * per, hi = lo + per < T ? lo + per : T;
  Fr local = Fr::zero();
  for (size_t i = lo; i < hi; i++) local = fp_add(local, v[i]);
  sm[threadIdx.x] = local;
  __syncthreads();
  // inclusive suffix scan of the per-thread sums
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {
    Fr other = threadIdx.x + o < blockDim.x ? sm[threadIdx.x + o] : Fr::zero();
    __syncthreads();
    sm[threadIdx.x] = fp_add(sm[threadIdx.x], other);
    __syncthreads();
  }
  Fr run = threadIdx.x + 1 < blockDim.x ? sm[threadIdx.x + 1] : Fr::zero();  // everything right of this thread
  for (size_t i = hi; i-- > lo;) {
    Fr cur = v[i];
    v[i] = run;
    run = fp_add(run, cur);
  }
}

__global__ void __launch_bounds__(128) div_finish(const DivArgs* __restrict__ argv, int count, const Fr* __restrict__ carries, size_t tmax) {
  const DivArgs& a = argv[blockIdx.y / count];
  const int b = blockIdx.y % count;
  carries += (size_t)(blockIdx.y - b) * tmax;
  const size_t len = a.len[b];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t T = (len + DIV_CHUNK - 1) / DIV_CHUNK;
  if (t >= T) return;
  const Fr* p = a.src[b];
  Fr* q = a.dst[b];
  const Fr x = a.x[b], xinv = a.xinv[b];
  const size_t start = t * DIV_CHUNK, end = start + DIV_CHUNK < len ? start + DIV_CHUNK : len;
  Fr run = carries[b * tmax + t];        // sum_{j >= end} p_j x^j
  Fr pw = fp_pow_u64(x, end - 1);        // x^j for j = end-1
  Fr ipw = fp_pow_u64(xinv, end);        // x^-(j+1)
  for (size_t j = end; j-- > start;) {
    q[j] = j + 1 < len ? fp_mul(run, ipw) : Fr::zero();  // the quotient has len-1 coefficients
    run = fp_add(run, fp_mul(p[j], pw));
    pw = fp_mul(pw, xinv);
    ipw = fp_mul(ipw, x);
  }
}

void divide_linear(capgpu_ctx* ctx, const DivArgs* args, int count, int G, size_t maxlen, Fr* scratch, size_t tmax) {
  const size_t T = (maxlen + DIV_CHUNK - 1) / DIV_CHUNK;
  CAPGPU_REQUIRE(T <= tmax, "division scratch too small");
  dim3 grid(ceil_div(T, 128), count * G);
  div_chunk_sums<<<grid, 128, 0, ctx->stream>>>(args, count, scratch, tmax);
  CAPGPU_LAUNCH_CHECK(ctx);
  div_scan<<<count * G, 1024, 0, ctx->stream>>>(args, count, scratch, tmax);
  CAPGPU_LAUNCH_CHECK(ctx);
  div_finish<<<grid, 128, 0, ctx->stream>>>(args, count, scratch, tmax);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// public-input evaluations: zeros except rows < num_inputs
__global__ void fill_pi_kernel(Fr* dst, size_t stride, size_t n, const Fr* __restrict__ pub, size_t pub_stride, size_t l) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  dst[(size_t)blockIdx.y * stride + j] = j < l ? pub[(size_t)blockIdx.y * pub_stride + j] : Fr::zero();
}

void fill_pi(capgpu_ctx* ctx, Fr* dst, size_t stride, size_t n, int G, const Fr* pub, size_t pub_stride, size_t l) {
  fill_pi_kernel<<<dim3(ceil_div(n, 256), G), 256, 0, ctx->stream>>>(dst, stride, n, pub, pub_stride, l);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// Evaluation-form commitment of a masked wire polynomial: scalars of the bases P_0, P_1, P_n, P_{n+1}
// appended after the n evaluations of each row: (b0 + b1 X)(X^n - 1) = -b0 - b1 X + b0 X^n + b1 X^(n+1).
__global__ void lagrange_tail_kernel(Fr* evals, size_t stride, size_t n, int G, const BlindArgs* __restrict__ argv) {
  int r = threadIdx.x >> 1, t = threadIdx.x & 1;
  const BlindArgs& args = argv[blockIdx.x];
  if (r >= args.rows_blinded) return;
  Fr b = args.b[r * 2 + t];
  Fr* row = evals + ((size_t)r * G + blockIdx.x) * stride + n;
  row[t] = fp_neg(b);
  row[2 + t] = b;
}

void lagrange_tail(capgpu_ctx* ctx, Fr* evals, size_t stride, size_t n, int G, const BlindArgs* args) {
  lagrange_tail_kernel<<<G, 32, 0, ctx->stream>>>(evals, stride, n, G, args);
  CAPGPU_LAUNCH_CHECK(ctx);
}

void blind(capgpu_ctx* ctx, Fr* polys, size_t stride, size_t n, int nrows, int G, int nb, const BlindArgs* args) {
  blind_kernel<<<dim3(nrows, G), 32, 0, ctx->stream>>>(polys, stride, n, nrows, G, nb, args);
  CAPGPU_LAUNCH_CHECK(ctx);
}

}  // namespace capgpu
