// Argument blocks and launchers of the polynomial-side prover kernels (poly.cu).
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
  const Fr* zh_inv;  // 8 entries (device)
};

struct EvalArgs {
  const Fr* poly[10];
  size_t len[10];
  Fr x[10];
};

struct LinArgs {
  const Fr* polys;  // 7 rows (w0..w4, pi, z), stride pstride
  const Fr* split;  // 5 rows, stride pstride
  const Fr* sel;    // 13 x n coefficients
  const Fr* sig;    // 5 x n coefficients
  size_t pstride, n, len;
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

void blind(capgpu_ctx* ctx, Fr* polys, size_t stride, size_t n, int nrows, int nb, const BlindArgs& args);
void lagrange_tail(capgpu_ctx* ctx, Fr* evals, size_t stride, size_t n, const BlindArgs& args);
void fill_pi(capgpu_ctx* ctx, Fr* dst, size_t n, const Fr* pub, size_t l);
void grand_product(capgpu_ctx* ctx, const Fr* wires, size_t wstride, const Fr* sig_eval, const Fr* omega_pows, size_t n,
                   const GpArgs& a, Fr* num, Fr* den, Fr* cn, Fr* cd, Fr* z);
void quotient_evals(capgpu_ctx* ctx, const Fr* coset, const Fr* sel, const Fr* sig, const Fr* xs, const Fr* l1inv, size_t m,
                    const QuotArgs& a, Fr* out);
void coset_tables(capgpu_ctx* ctx, const Fr* omega_m, size_t m, const Fr& gen, const Fr& n_mont, Fr* xs, Fr* l1inv);
void split_quotient(capgpu_ctx* ctx, const Fr* t, size_t n, size_t m, Fr* split, size_t stride, const BlindArgs& args, uint32_t* flag);
void evaluate(capgpu_ctx* ctx, const EvalArgs& a, int count, Fr* out, Fr* scratch /* 16 * count */);
void lin_batch(capgpu_ctx* ctx, const LinArgs& a, Fr* lin, Fr* batch);
void divide_linear(capgpu_ctx* ctx, const DivArgs& a, int count, Fr* scratch /* count * tmax */, size_t tmax);

}  // namespace capgpu
