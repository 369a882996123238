// Argument blocks and launchers of the polynomial-side prover kernels (poly.cu).
//
// Every launcher works on a GROUP of G proofs proved in lockstep over one proving key: vectors
// of the same kind are stored row-major as [row kind][proof][elements] (row (r, g) at
// (r * G + g) * stride), so the NTT / MSM launches of a round see one uniform batch of rows, and
// the per-proof scalars (Fiat-Shamir challenges, blinders, evaluation points) live in device
// arrays indexed by the proof's slot g = blockIdx.y.
#pragma once
#include "common.cuh"

namespace capgpu {

struct BlindArgs {
  Fr b[10];          // blinders, row-major (rows x nb)
  int rows_blinded;  // rows beyond this only get their padding cleared
};

struct GpArgs {
  Fr beta, gamma;
  Fr k[5];
};

struct QuotArgs {
  Fr alpha, alpha2, beta, gamma;
  Fr k[5];
};

struct EvalArgs {
  const Fr* poly[10];
  size_t len[10];
  Fr x[10];
};

struct LinArgs {
  const Fr* polys;  // row (r, g) of the [7][G] polynomial block for this proof's g: polys + r * rstride
  const Fr* split;  // row (i, g) of the [5][G] split block: split + i * rstride
  size_t rstride;   // G * row stride
  Fr cs_sel[13];
  Fr cz, csig;
  Fr ct[5];
  Fr vp[9];
};

struct DivArgs {
  const Fr* src[2];
  Fr* dst[2];
  size_t len[2];
  Fr x[2];
  Fr xinv[2];
};

// polys: [nrows][G] rows of `stride` elements; args[g]
void blind(capgpu_ctx* ctx, Fr* polys, size_t stride, size_t n, int nrows, int G, int nb, const BlindArgs* args);
void lagrange_tail(capgpu_ctx* ctx, Fr* evals, size_t stride, size_t n, int G, const BlindArgs* args);
// dst: G rows (stride `stride`) of n evaluations; pub: G rows of pub_stride elements
void fill_pi(capgpu_ctx* ctx, Fr* dst, size_t stride, size_t n, int G, const Fr* pub, size_t pub_stride, size_t l);
// wires: [5][G] rows of wstride; num / den / z: G rows of n; cn / cd: G rows of cstride
void grand_product(capgpu_ctx* ctx, const Fr* wires, size_t wstride, const Fr* sig_eval, const Fr* omega_pows, size_t n, int G,
                   const GpArgs* args, Fr* num, Fr* den, Fr* cn, Fr* cd, size_t cstride, Fr* z);
// The quotient domain is `cosets` cosets s_k H of `sub` points each (m = cosets * sub): one coset of 8n points (s_0 = g), or the
// three cosets g rho^k H_2n of the 6n-point domain.  Point (k, i) = s_k w_sub^i sits at index k * sub + i; w_n x is `step` = sub / n
// places further inside the coset; Z_H(x) = s_k^n (w_sub^n)^i - 1 has period `step` in i: zh_inv[k * step + (i mod step)].
struct QuotDomain {
  size_t m, sub;
  uint32_t log_sub, step;
};
// coset: [7][G] rows of m; out: G rows of m
void quotient_evals(capgpu_ctx* ctx, const Fr* coset, const Fr* sel, const Fr* sig, const Fr* xs, const Fr* l1inv, const Fr* zh_inv,
                    const QuotDomain& qd, int G, const QuotArgs* args, Fr* out);
// xs[k * sub + i] = shift[k] * omega_sub[i], l1inv = 1 / (n (xs - 1))
struct CosetShifts { Fr s[3]; };
void coset_tables(capgpu_ctx* ctx, const Fr* omega_sub, const QuotDomain& qd, const CosetShifts& shifts, const Fr& n_mont, Fr* xs, Fr* l1inv);
// t: G rows of m; split: [5][G] rows of stride; flag[g]
void split_quotient(capgpu_ctx* ctx, const Fr* t, size_t n, size_t m, int G, Fr* split, size_t stride, const BlindArgs* args, uint32_t* flag);
// out: G rows of 16 (10 used); scratch: G x 160
void evaluate(capgpu_ctx* ctx, const EvalArgs* args, int count, int G, Fr* out, Fr* scratch);
// sel / sig: the key's coefficient polynomials (13 x n, 5 x n); lin / batch: G rows of ostride
void lin_batch(capgpu_ctx* ctx, const LinArgs* args, const Fr* sel, const Fr* sig, size_t n, size_t len, int G, Fr* lin, Fr* batch,
               size_t ostride);
// scratch: G x count x tmax
void divide_linear(capgpu_ctx* ctx, const DivArgs* args, int count, int G, size_t maxlen, Fr* scratch, size_t tmax);

}  // namespace capgpu
