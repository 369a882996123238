// Lane-cooperative G1 group operations for the latency-bound stages of the MSM (bucket reduction,
// final fold): a group of 2 ("pair") or 4 ("quad") adjacent lanes holds the SAME operands and each
// lane computes one of the independent field products of a level of the EFD formulas
// (add-2008-s, dbl-2008-s-1 — the formulas of ec.cuh), the results being exchanged with warp
// shuffles.  A full XYZZ addition is 14 products in a dependent chain of 7 (pair) or 4 (quad)
// product latencies instead of 14; a doubling 9 products in 5 (pair) or 3 (quad).  Special cases
// (infinity, doubling, cancellation) are decided identically by every lane of the group because the
// operands are replicated.  Results are the same group elements as ec.cuh's (representation may
// differ by the usual projective scaling; everything is compared after conversion to affine).
//
// Part of the replacement of ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul` (reached from
// /root/reference/src/proof/transfer.rs:181); see msm.cu.
#pragma once
#include "ec.cuh"

namespace capgpu {

__device__ __forceinline__ Fq fq_sel(bool c, const Fq& a, const Fq& b) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
__device__ __forceinline__ Fq fq_xchg(const Fq& a, uint32_t pmask) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(pmask, a.v[i], 1);
  return r;
}
__device__ __forceinline__ Fq fq_from_lane(const Fq& a, uint32_t mask, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(mask, a.v[i], src);
  return r;
}

// ---- pair -----------------------------------------------------------------------------------
// r0 = a0 * b0, r1 = a1 * b1; lane `role` of the pair computes product `role`
__device__ __forceinline__ void pair_mul2(const Fq& a0, const Fq& b0, const Fq& a1, const Fq& b1, bool role, uint32_t pmask, Fq& r0,
                                          Fq& r1) {
  Fq m = fp_mul(fq_sel(role, a1, a0), fq_sel(role, b1, b0));
  Fq o = fq_xchg(m, pmask);
  r0 = fq_sel(role, o, m);
  r1 = fq_sel(role, m, o);
}

__device__ __noinline__ G1XYZZ xyzz_dbl_pair(const G1XYZZ& p, bool role, uint32_t pmask) {
  if (p.is_inf()) return p;
  Fq U = fp_dbl(p.Y);
  Fq V, XX, W, S, MM, t1, t2;
  G1XYZZ r;
  pair_mul2(U, U, p.X, p.X, role, pmask, V, XX);
  Fq M = fp_add(fp_dbl(XX), XX);
  pair_mul2(U, V, p.X, V, role, pmask, W, S);
  pair_mul2(M, M, V, p.ZZ, role, pmask, MM, r.ZZ);
  r.X = fp_sub(MM, fp_dbl(S));
  pair_mul2(M, fp_sub(S, r.X), W, p.Y, role, pmask, t1, t2);
  r.Y = fp_sub(t1, t2);
  // both lanes need W * ZZZ; the pair has no second product left to share it with
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}

__device__ __noinline__ void xyzz_add_pair(G1XYZZ& acc, const G1XYZZ& q, bool role, uint32_t pmask) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  Fq U1, U2, S1, S2;
  pair_mul2(acc.X, q.ZZ, q.X, acc.ZZ, role, pmask, U1, U2);
  pair_mul2(acc.Y, q.ZZZ, q.Y, acc.ZZZ, role, pmask, S1, S2);
  Fq P = fp_sub(U2, U1);
  Fq Rr = fp_sub(S2, S1);
  if (P.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl_pair(acc, role, pmask);
    else acc = G1XYZZ::inf();
    return;
  }
  Fq ZZ12, ZZZ12, PP, RR, PPP, Qq, t1, t2;
  pair_mul2(acc.ZZ, q.ZZ, acc.ZZZ, q.ZZZ, role, pmask, ZZ12, ZZZ12);
  pair_mul2(P, P, Rr, Rr, role, pmask, PP, RR);
  pair_mul2(P, PP, U1, PP, role, pmask, PPP, Qq);
  acc.X = fp_sub(fp_sub(RR, PPP), fp_dbl(Qq));
  pair_mul2(ZZ12, PP, ZZZ12, PPP, role, pmask, acc.ZZ, acc.ZZZ);
  pair_mul2(Rr, fp_sub(Qq, acc.X), S1, PPP, role, pmask, t1, t2);
  acc.Y = fp_sub(t1, t2);
}

__device__ inline G1XYZZ xyzz_mul_small_pair(const G1XYZZ& p, uint32_t k, bool role, uint32_t pmask) {
  G1XYZZ r = G1XYZZ::inf();
  int top = 31;
  while (top >= 0 && !((k >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    r = xyzz_dbl_pair(r, role, pmask);
    if ((k >> i) & 1) xyzz_add_pair(r, p, role, pmask);
  }
  return r;
}

// ---- quad -----------------------------------------------------------------------------------
// The four lanes 4g .. 4g+3 of a warp hold the same operands; `role` = lane & 3 picks the product a lane
// computes at each level.  The quad operations are BRANCH-FREE AT WARP LEVEL and use full-mask shuffles of
// width 4: every lane of a converged warp must call them together (quads with nothing to do pass the point
// at infinity).  Special cases are resolved with selects after the generic formula; only an exact doubling
// (equal operands) makes the whole warp run the doubling formula once more.  [Partial-mask shuffles compile
// to WARPSYNC.COLLECTIVE call sequences and cost more than the products they exchange.]
__device__ __forceinline__ uint32_t quad_role() { return threadIdx.x & 3u; }

// One shared copy of the product / squaring (arguments and result in registers): with the products inlined a
// quad addition is ~22 KB of straight-line SASS and a lone warp stalls on instruction fetch (ncu:
// no_instruction 1.6 of 4.7 cycles per issue in msm_red_planes); out of line the loop bodies stay cached.
static __device__ __noinline__ Fq fq_mul_ool(Fq a, Fq b) { return fp_mul(a, b); }
static __device__ __noinline__ Fq fq_sqr_ool(Fq a) { return fp_sqr(a); }

__device__ __forceinline__ Fq fq_sel4(uint32_t role, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
  Fq r;
  const bool hi = role & 2u, odd = role & 1u;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t lo2 = odd ? a1.v[i] : a0.v[i];
    uint32_t hi2 = odd ? a3.v[i] : a2.v[i];
    r.v[i] = hi ? hi2 : lo2;
  }
  return r;
}
__device__ __forceinline__ Fq fq_quad_get(const Fq& m, int j) {  // m of lane j of the caller's quad
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, m.v[i], j, 4);
  return r;
}
__device__ __forceinline__ G1XYZZ xyzz_sel(bool c, const G1XYZZ& a, const G1XYZZ& b) {
  G1XYZZ r;
  r.X = fq_sel(c, a.X, b.X); r.Y = fq_sel(c, a.Y, b.Y); r.ZZ = fq_sel(c, a.ZZ, b.ZZ); r.ZZZ = fq_sel(c, a.ZZZ, b.ZZZ);
  return r;
}

// EFD dbl-2008-s-1 (a = 0) in three product levels.  Infinity (ZZ = 0) maps to infinity (ZZ3 = V * ZZ = 0).
__device__ __forceinline__ G1XYZZ xyzz_dbl_quad(const G1XYZZ& p, uint32_t role) {
  Fq U = fp_dbl(p.Y);
  G1XYZZ r;
  // level 1: V = U^2, XX = X^2 (every lane runs the cheaper squaring)
  Fq m = fq_sqr_ool(fq_sel((role & 1u) != 0, p.X, U));
  Fq V = fq_quad_get(m, 0), XX = fq_quad_get(m, 1);
  Fq M = fp_add(fp_dbl(XX), XX);
  // level 2: W = U V, S = X V, MM = M^2, ZZ3 = V ZZ
  m = fq_mul_ool(fq_sel4(role, U, p.X, M, V), fq_sel4(role, V, V, M, p.ZZ));
  Fq W = fq_quad_get(m, 0), S = fq_quad_get(m, 1), MM = fq_quad_get(m, 2);
  r.ZZ = fq_quad_get(m, 3);
  r.X = fp_sub(MM, fp_dbl(S));
  // level 3: t1 = M (S - X3), t2 = W Y, ZZZ3 = W ZZZ
  m = fq_mul_ool(fq_sel4(role, M, W, W, W), fq_sel4(role, fp_sub(S, r.X), p.Y, p.ZZZ, p.ZZZ));
  r.Y = fp_sub(fq_quad_get(m, 0), fq_quad_get(m, 1));
  r.ZZZ = fq_quad_get(m, 2);
  return r;
}

// acc += b: EFD add-2008-s in four product levels, exact special cases
__device__ __forceinline__ void xyzz_add_quad(G1XYZZ& acc, const G1XYZZ& b, uint32_t role) {
  const bool a_inf = acc.is_inf(), b_inf = b.is_inf();
  // level 1: U1 = X1 ZZ2, U2 = X2 ZZ1, S1 = Y1 ZZZ2, S2 = Y2 ZZZ1
  Fq m = fq_mul_ool(fq_sel4(role, acc.X, b.X, acc.Y, b.Y), fq_sel4(role, b.ZZ, acc.ZZ, b.ZZZ, acc.ZZZ));
  Fq U1 = fq_quad_get(m, 0), S1 = fq_quad_get(m, 2);
  Fq P = fp_sub(fq_quad_get(m, 1), U1);
  Fq Rr = fp_sub(fq_quad_get(m, 3), S1);
  // level 2: PP = P^2, RR = R^2, ZZ12 = ZZ1 ZZ2, ZZZ12 = ZZZ1 ZZZ2
  m = fq_mul_ool(fq_sel4(role, P, Rr, acc.ZZ, acc.ZZZ), fq_sel4(role, P, Rr, b.ZZ, b.ZZZ));
  Fq PP = fq_quad_get(m, 0), RR = fq_quad_get(m, 1), ZZ12 = fq_quad_get(m, 2), ZZZ12 = fq_quad_get(m, 3);
  // level 3: PPP = P PP, Q = U1 PP, ZZ3 = ZZ12 PP
  m = fq_mul_ool(fq_sel4(role, P, U1, ZZ12, ZZ12), PP);
  Fq PPP = fq_quad_get(m, 0), Qq = fq_quad_get(m, 1);
  G1XYZZ r;
  r.ZZ = fq_quad_get(m, 2);
  r.X = fp_sub(fp_sub(RR, PPP), fp_dbl(Qq));
  // level 4: t1 = R (Q - X3), t2 = S1 PPP, ZZZ3 = ZZZ12 PPP
  m = fq_mul_ool(fq_sel4(role, Rr, S1, ZZZ12, ZZZ12), fq_sel4(role, fp_sub(Qq, r.X), PPP, PPP, PPP));
  r.Y = fp_sub(fq_quad_get(m, 0), fq_quad_get(m, 1));
  r.ZZZ = fq_quad_get(m, 2);
  // special cases: equal x-coordinates (P = 0) with both operands finite
  const bool same_x = !a_inf && !b_inf && P.is_zero();
  const bool need_dbl = same_x && Rr.is_zero();
  if (__any_sync(0xffffffffu, need_dbl)) {
    G1XYZZ d = xyzz_dbl_quad(acc, role);
    r = xyzz_sel(need_dbl, d, r);
  }
  if (same_x && !need_dbl) r = G1XYZZ::inf();  // P + (-P)
  r = xyzz_sel(a_inf, b, r);
  acc = xyzz_sel(b_inf, acc, r);
}


// One shared out-of-line copy of the quad addition for the reduction kernels (operands and result through
// memory: shared, global or local): every call site of a kernel then runs the same ~600 cached instructions
// instead of its own inlined copy - a lone warp walking cold straight-line code waits on instruction fetch.
// *acc += *b; lane `role` of the quad stores coordinate `role` of the sum to dst when `store` is set.
static __device__ __noinline__ void xyzz_add_quad_mem(const G1XYZZ* acc, const G1XYZZ* b, G1XYZZ* dst, bool store, uint32_t role) {
  G1XYZZ x = *acc, y = *b;
  xyzz_add_quad(x, y, role);
  if (store) reinterpret_cast<Fq*>(dst)[role] = fq_sel4(role, x.X, x.Y, x.ZZ, x.ZZZ);
}

}  // namespace capgpu
