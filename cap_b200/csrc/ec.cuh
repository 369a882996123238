// BN254 G1 (y^2 = x^3 + 3 over Fq, a = 0) group law for the MSM kernels.
//
// Replaces ark-ec 0.3.0 `short_weierstrass_jacobian::{GroupAffine, GroupProjective}` as used
// by `VariableBaseMSM::multi_scalar_mul` under `KZG10::commit` (reached from
// /root/reference/src/proof/transfer.rs:181; G1Affine bound at src/config.rs:27-36).
// Accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2;
// ZZ == 0 is the point at infinity): mixed addition of an affine base costs 8M+2S (EFD
// madd-2008-s) against 7M+4S for arkworks' Jacobian add_assign_mixed.  Results are only
// ever compared after conversion to affine, where the representation is unique.
#pragma once
#include "fp.cuh"

namespace capgpu {

struct alignas(32) G1Affine {
  Fq x, y;  // Montgomery; the all-zero pattern encodes infinity ((0,0) is not on the curve)
  __host__ __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

struct alignas(32) G1XYZZ {
  Fq X, Y, ZZ, ZZZ;
  __host__ __device__ __forceinline__ bool is_inf() const { return ZZ.is_zero(); }
  __host__ __device__ __forceinline__ static G1XYZZ inf() {
    G1XYZZ r; r.X = Fq::zero(); r.Y = Fq::zero(); r.ZZ = Fq::zero(); r.ZZZ = Fq::zero(); return r;
  }
};

__host__ __device__ __forceinline__ G1XYZZ xyzz_from_affine(const G1Affine& p) {
  G1XYZZ r;
  if (p.is_inf()) return G1XYZZ::inf();
  r.X = p.x; r.Y = p.y; r.ZZ = Fq::one(); r.ZZZ = Fq::one();
  return r;
}

// EFD dbl-2008-s-1 (a = 0)
__host__ __device__ inline G1XYZZ xyzz_dbl(const G1XYZZ& p) {
  if (p.is_inf()) return p;
  Fq U = fp_dbl(p.Y);
  Fq V = fp_sqr(U);
  Fq W = fp_mul(U, V);
  Fq S = fp_mul(p.X, V);
  Fq XX = fp_sqr(p.X);
  Fq M = fp_add(fp_dbl(XX), XX);
  G1XYZZ r;
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_mul_sub(M, fp_sub(S, r.X), W, p.Y);  // one reduction for the two products
  r.ZZ = fp_mul(V, p.ZZ);
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}

// EFD mdbl-2008-s-1: doubling of an affine point
__host__ __device__ inline G1XYZZ xyzz_dbl_affine(const Fq& x, const Fq& y) {
  Fq U = fp_dbl(y);
  Fq V = fp_sqr(U);
  Fq W = fp_mul(U, V);
  Fq S = fp_mul(x, V);
  Fq XX = fp_sqr(x);
  Fq M = fp_add(fp_dbl(XX), XX);
  G1XYZZ r;
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_mul_sub(M, fp_sub(S, r.X), W, y);
  r.ZZ = V;
  r.ZZZ = W;
  return r;
}

// acc += (x, y) with (x, y) a finite affine point; `neg` flips the sign of the addend.
// EFD madd-2008-s, with the doubling / cancellation cases handled exactly.
__host__ __device__ inline void xyzz_add_mixed(G1XYZZ& acc, const Fq& x2, const Fq& y2_in, bool neg) {
  Fq y2 = neg ? fp_neg(y2_in) : y2_in;
  if (acc.is_inf()) {
    acc.X = x2; acc.Y = y2; acc.ZZ = Fq::one(); acc.ZZZ = Fq::one();
    return;
  }
  Fq U2 = fp_mul(x2, acc.ZZ);
  Fq S2 = fp_mul(y2, acc.ZZZ);
  Fq P = fp_sub(U2, acc.X);
  Fq Rr = fp_sub(S2, acc.Y);
  if (P.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl_affine(x2, y2);
    else acc = G1XYZZ::inf();
    return;
  }
  Fq PP = fp_sqr(P);
  Fq PPP = fp_mul(P, PP);
  Fq Qq = fp_mul(acc.X, PP);
  Fq X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  Fq Y3 = fp_mul_sub(Rr, fp_sub(Qq, X3), acc.Y, PPP);  // fused: 9 reductions per mixed addition instead of 10
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = fp_mul(acc.ZZ, PP);
  acc.ZZZ = fp_mul(acc.ZZZ, PPP);
}

// acc = (x1, y1) + (x2, y2), both affine and finite (EFD mmadd-2008-s: 6 products instead of the 10 of a mixed addition,
// the ZZ = ZZZ = 1 of the first operand folded away).  The first addition into every bucket; false (acc untouched) when x1 == x2.
__host__ __device__ inline bool xyzz_set_affine2(G1XYZZ& acc, const Fq& x1, const Fq& y1, const Fq& x2, const Fq& y2) {
  Fq P = fp_sub(x2, x1);
  Fq Rr = fp_sub(y2, y1);
  if (P.is_zero()) return false;  // equal or opposite points: left to the caller's general path
  Fq PP = fp_sqr(P);
  Fq PPP = fp_mul(P, PP);
  Fq Qq = fp_mul(x1, PP);
  Fq X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  acc.Y = fp_mul_sub(Rr, fp_sub(Qq, X3), y1, PPP);
  acc.X = X3;
  acc.ZZ = PP;
  acc.ZZZ = PPP;
  return true;
}

// acc += q (both XYZZ).  EFD add-2008-s with exact special cases.
__host__ __device__ inline void xyzz_add(G1XYZZ& acc, const G1XYZZ& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  Fq U1 = fp_mul(acc.X, q.ZZ);
  Fq U2 = fp_mul(q.X, acc.ZZ);
  Fq S1 = fp_mul(acc.Y, q.ZZZ);
  Fq S2 = fp_mul(q.Y, acc.ZZZ);
  Fq P = fp_sub(U2, U1);
  Fq Rr = fp_sub(S2, S1);
  if (P.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl(acc);
    else acc = G1XYZZ::inf();
    return;
  }
  Fq PP = fp_sqr(P);
  Fq PPP = fp_mul(P, PP);
  Fq Qq = fp_mul(U1, PP);
  Fq X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  Fq Y3 = fp_mul_sub(Rr, fp_sub(Qq, X3), S1, PPP);
  acc.X = X3;
  acc.Y = Y3;
  acc.ZZ = fp_mul(fp_mul(acc.ZZ, q.ZZ), PP);
  acc.ZZZ = fp_mul(fp_mul(acc.ZZZ, q.ZZZ), PPP);
}

// XYZZ -> affine (one field inversion); infinity maps to the all-zero pattern.
__host__ __device__ inline G1Affine xyzz_to_affine(const G1XYZZ& p) {
  G1Affine r;
  if (p.is_inf()) { r.x = Fq::zero(); r.y = Fq::zero(); return r; }
  Fq t = fp_inv(fp_mul(p.ZZ, p.ZZZ));
  Fq izz = fp_mul(t, p.ZZZ);
  Fq izzz = fp_mul(t, p.ZZ);
  r.x = fp_mul(p.X, izz);
  r.y = fp_mul(p.Y, izzz);
  return r;
}

// k * p for a small integer k (double-and-add, MSB first); used by the bucket reduction.
__host__ __device__ inline G1XYZZ xyzz_mul_small(const G1XYZZ& p, uint32_t k) {
  G1XYZZ r = G1XYZZ::inf();
  int top = 31;
  while (top >= 0 && !((k >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((k >> i) & 1) xyzz_add(r, p);
  }
  return r;
}

}  // namespace capgpu
