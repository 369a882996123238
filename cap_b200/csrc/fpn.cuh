// Generic N x 32-bit-limb Montgomery fields and short-Weierstrass (a = 0) G1 arithmetic for the other two
// pairing curves the reference can be built for: BLS12-381 and BLS12-377 (cargo features `bls12_381` /
// `bls12_377`, /root/reference/src/config.rs:86-114, Cargo.toml:71-75).  Their base fields are 381 / 377 bits
// = 12 limbs (ark-ff `Fp384`: 6 x u64 little-endian, Montgomery with R = 2^384 - the same bytes as 12 x u32),
// their scalar fields fit the 8-limb scalars the MSM already recodes.  This is the portable, loop-written
// counterpart of fp.cuh / ec.cuh (which are hand-scheduled for 8 limbs): the same algorithms - CIOS Montgomery
// product, binary-GCD inversion in batched rounds, XYZZ group law (EFD madd-2008-s / add-2008-s /
// dbl-2008-s-1) - written over N with 64-bit partial products, so the one source serves both curves and the
// host emulation used by the CPU tests.  Used by msm_curve.cu (capgpu_curve_msm_g1).
#pragma once
#include "fp.cuh"

namespace capgpu {

struct Bls381FqParams {
  // q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
  static constexpr int N = 12;
  static constexpr int BITS = 381;
  static __host__ __device__ __forceinline__ constexpr uint32_t p(int i) { constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}; return t[i]; }
  static constexpr uint32_t INV = 0xfffcfffdu;  // -q^-1 mod 2^32
  static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) { constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u}; return t[i]; }  // R mod q, R = 2^384
  static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) { constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u}; return t[i]; }  // R^2 mod q
};
struct Bls377FqParams {
  // q = 0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
  static constexpr int N = 12;
  static constexpr int BITS = 377;
  static __host__ __device__ __forceinline__ constexpr uint32_t p(int i) { constexpr uint32_t t[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu, 0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u}; return t[i]; }
  static constexpr uint32_t INV = 0xffffffffu;  // -q^-1 mod 2^32
  static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) { constexpr uint32_t t[12] = {0xffffff68u, 0x02cdffffu, 0x7fffffb1u, 0x51409f83u, 0x8a7d3ff2u, 0x9f7db3a9u, 0x6e7c6305u, 0x7b4e97b7u, 0x803c84e8u, 0x4cf495bfu, 0xe2fdf49au, 0x008d6661u}; return t[i]; }  // R mod q, R = 2^384
  static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) { constexpr uint32_t t[12] = {0x9400cd22u, 0xb786686cu, 0xb00431b1u, 0x0329fcaau, 0x62d6b46du, 0x22a5f111u, 0x827dc3acu, 0xbfdf7d03u, 0x41790bf9u, 0x837e92f0u, 0x1e914b88u, 0x006dfccbu}; return t[i]; }  // R^2 mod q
};

template <class PR>
struct alignas(16) FpN {
  static constexpr int N = PR::N;
  uint32_t v[PR::N];
  __host__ __device__ __forceinline__ static FpN zero() { FpN r; for (int i = 0; i < N; i++) r.v[i] = 0; return r; }
  __host__ __device__ __forceinline__ static FpN one() { FpN r; for (int i = 0; i < N; i++) r.v[i] = PR::one(i); return r; }
  __host__ __device__ __forceinline__ static FpN r2() { FpN r; for (int i = 0; i < N; i++) r.v[i] = PR::r2(i); return r; }
  __host__ __device__ __forceinline__ bool is_zero() const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= v[i]; return o == 0; }
  __host__ __device__ __forceinline__ bool operator==(const FpN& b) const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i]; return o == 0; }
  __host__ __device__ __forceinline__ bool operator!=(const FpN& b) const { return !(*this == b); }
};

// a - p if a >= p else a, for a < 2p given with its carry bit `hi` (bit 32 N)
template <class PR>
__host__ __device__ __forceinline__ void fpn_cond_sub(FpN<PR>& a, uint32_t hi) {
  constexpr int N = PR::N;
  uint32_t t[N];
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { uint64_t d = (uint64_t)a.v[i] - PR::p(i) - br; t[i] = (uint32_t)d; br = (d >> 32) & 1u; }
  const bool ge = hi != 0 || br == 0;
#pragma unroll
  for (int i = 0; i < N; i++) a.v[i] = ge ? t[i] : a.v[i];
}

template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_add(const FpN<PR>& a, const FpN<PR>& b) {
  constexpr int N = PR::N;
  FpN<PR> r;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
  fpn_cond_sub(r, (uint32_t)c);
  return r;
}
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_sub(const FpN<PR>& a, const FpN<PR>& b) {
  constexpr int N = PR::N;
  FpN<PR> r;
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - br; r.v[i] = (uint32_t)d; br = (d >> 32) & 1u; }
  const uint32_t m = br ? 0xffffffffu : 0u;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { c += (uint64_t)r.v[i] + (PR::p(i) & m); r.v[i] = (uint32_t)c; c >>= 32; }
  return r;
}
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_neg(const FpN<PR>& a) { return a.is_zero() ? a : fp_sub(FpN<PR>::zero(), a); }
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_dbl(const FpN<PR>& a) { return fp_add(a, a); }

// CIOS Montgomery product (Koc et al.): row of a * b_i, then one reduction step, N times.  Out of line on the
// device (one copy per field instead of one per call site: the 12-limb product is ~600 instructions and the
// group law calls it 10-14 times; keeps msm_curve.cu's compile time and instruction footprint small).
#ifdef __CUDACC__
#define CAPGPU_FPN_NOINLINE __noinline__
#else
#define CAPGPU_FPN_NOINLINE
#endif
template <class PR>
__host__ __device__ CAPGPU_FPN_NOINLINE FpN<PR> fp_mul(const FpN<PR>& a, const FpN<PR>& b) {
  constexpr int N = PR::N;
  uint32_t t[N + 2];
#pragma unroll
  for (int i = 0; i < N + 2; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < N; j++) { c += (uint64_t)a.v[j] * b.v[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
    c += t[N]; t[N] = (uint32_t)c; t[N + 1] = (uint32_t)(c >> 32);
    const uint32_t m = t[0] * PR::INV;
    c = ((uint64_t)m * PR::p(0) + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < N; j++) { c += (uint64_t)m * PR::p(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
    c += t[N]; t[N - 1] = (uint32_t)c; t[N] = t[N + 1] + (uint32_t)(c >> 32);
  }
  FpN<PR> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = t[i];
  fpn_cond_sub(r, t[N]);
  return r;
}
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_sqr(const FpN<PR>& a) { return fp_mul(a, a); }
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_mul_sub(const FpN<PR>& a, const FpN<PR>& b, const FpN<PR>& c, const FpN<PR>& d) {
  return fp_sub(fp_mul(a, b), fp_mul(c, d));
}
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_from_mont(const FpN<PR>& a) { FpN<PR> o = FpN<PR>::zero(); o.v[0] = 1; return fp_mul(a, o); }
template <class PR>
__host__ __device__ __forceinline__ FpN<PR> fp_to_mont(const FpN<PR>& a) { return fp_mul(a, FpN<PR>::r2()); }

// a^(p-2) (cross-check of fp_inv)
template <class PR>
__host__ __device__ inline FpN<PR> fp_inv_fermat(const FpN<PR>& a) {
  constexpr int N = PR::N;
  uint32_t e[N];
  uint64_t br = 2;  // e = p - 2 (the low limb of the BLS12-377 modulus is 1: the borrow travels)
  for (int i = 0; i < N; i++) { uint64_t d = (uint64_t)PR::p(i) - br; e[i] = (uint32_t)d; br = (d >> 32) & 1u; }
  FpN<PR> r = FpN<PR>::one();
  for (int i = 32 * N - 1; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
  }
  return r;
}

// Binary GCD in batched rounds, as fp.cuh's fp_inv (same invariants: a R^2 = u y, b R^2 = v y mod p), over N limbs:
// ceil((2 BITS - 1) / 30) rounds of 30 steps on 62-bit approximations.  inv(0) = 0.
template <class PR>
__host__ __device__ CAPGPU_FPN_NOINLINE FpN<PR> fp_inv(const FpN<PR>& y) {
  constexpr int N = PR::N;
  if (y.is_zero()) return y;
  uint32_t a[N], b[N], u[N], v[N];
  for (int i = 0; i < N; i++) { a[i] = y.v[i]; b[i] = PR::p(i); u[i] = PR::r2(i); v[i] = 0; }
  constexpr int ROUNDS = (2 * PR::BITS - 1 + 29) / 30 + 1;
  for (int round = 0; round < ROUNDS; round++) {
    uint32_t nz = 0;
    for (int i = 0; i < N; i++) nz |= a[i];
    if (nz == 0) break;
    uint32_t top = a[1] | b[1];
    int n = 32;
    for (int i = 2; i < N; i++) { uint32_t w = a[i] | b[i]; if (w) { top = w; n = 32 * i; } }
    n += 32 - clz32(top);
    if (n < 62) n = 62;
    const int s = n - 32, q = s >> 5, r = s & 31;
    uint32_t alo = 0, ahi = 0, blo = 0, bhi = 0;
    for (int i = 0; i < N; i++) {
      if (i == q) { alo = a[i]; blo = b[i]; ahi = i + 1 < N ? a[i + 1] : 0u; bhi = i + 1 < N ? b[i + 1] : 0u; }
    }
    const uint32_t atop = r ? (alo >> r) | (ahi << (32 - r)) : alo;
    const uint32_t btop = r ? (blo >> r) | (bhi << (32 - r)) : blo;
    uint64_t xa = ((uint64_t)atop << 30) | (a[0] & 0x3fffffffu);
    uint64_t xb = ((uint64_t)btop << 30) | (b[0] & 0x3fffffffu);
    int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
    for (int j = 0; j < 30; j++) {
      const uint64_t d1 = (xa - xb) >> 1, d2 = (xb - xa) >> 1, h = xa >> 1;
      const bool odd = (xa & 1u) != 0;
      const bool sw = odd && xa < xb;
      const uint64_t nxa = odd ? (sw ? d2 : d1) : h;
      xb = sw ? xa : xb;
      xa = nxa;
      const int32_t tf0 = sw ? f1 : f0, tf1 = sw ? f0 : f1, tg0 = sw ? g1 : g0, tg1 = sw ? g0 : g1;
      f0 = tf0 - (odd ? tf1 : 0);
      g0 = tg0 - (odd ? tg1 : 0);
      f1 = tf1 * 2;
      g1 = tg1 * 2;
    }
    // (a, b) <- |(f0 a + g0 b, f1 a + g1 b)| / 2^30 exactly; signs folded into the matrix rows
    uint32_t na[N], nb[N], nu[N], nv[N];
    for (int row = 0; row < 2; row++) {
      int32_t f = row ? f1 : f0, g = row ? g1 : g0;
      const uint32_t fa = (uint32_t)(f < 0 ? -f : f), ga = (uint32_t)(g < 0 ? -g : g);
      const uint32_t mf = f < 0 ? 0xffffffffu : 0u, mg = g < 0 ? 0xffffffffu : 0u;
      uint32_t t[N + 1];
      uint64_t p1 = 0, p2 = 0, c1 = mf & 1u, c2 = mg & 1u, cs = 0;
      for (int i = 0; i <= N; i++) {
        p1 += i < N ? (uint64_t)a[i] * fa : 0; p2 += i < N ? (uint64_t)b[i] * ga : 0;
        c1 += (uint64_t)((uint32_t)p1 ^ mf); c2 += (uint64_t)((uint32_t)p2 ^ mg);
        cs += (uint64_t)(uint32_t)c1 + (uint32_t)c2;
        t[i] = (uint32_t)cs;
        p1 >>= 32; p2 >>= 32; c1 >>= 32; c2 >>= 32; cs >>= 32;
      }
      const bool neg = (t[N] >> 31) != 0;
      const uint32_t mn = neg ? 0xffffffffu : 0u;
      uint64_t cn = mn & 1u;
      for (int i = 0; i <= N; i++) { cn += (uint64_t)(t[i] ^ mn); t[i] = (uint32_t)cn; cn >>= 32; }
      uint32_t* dst = row ? nb : na;
      for (int i = 0; i < N; i++) dst[i] = (t[i] >> 30) | (t[i + 1] << 2);
      if (neg) { f = -f; g = -g; }
      // (u, v) row: (u f + v g) / 2^30 mod p with p - x standing in for -x
      uint32_t uu[N], vv[N];
      uint64_t bu = 0, bv = 0;
      for (int i = 0; i < N; i++) {
        uint64_t du = (uint64_t)PR::p(i) - u[i] - bu; bu = (du >> 32) & 1u;
        uint64_t dv = (uint64_t)PR::p(i) - v[i] - bv; bv = (dv >> 32) & 1u;
        uu[i] = f < 0 ? (uint32_t)du : u[i];
        vv[i] = g < 0 ? (uint32_t)dv : v[i];
      }
      const uint32_t fb = (uint32_t)(f < 0 ? -f : f), gb = (uint32_t)(g < 0 ? -g : g);
      uint32_t w[N + 1];
      uint64_t c = 0;
      for (int i = 0; i < N; i++) { c += (uint64_t)uu[i] * fb + (uint64_t)vv[i] * gb; w[i] = (uint32_t)c; c >>= 32; }
      w[N] = (uint32_t)c;
      const uint32_t qq = (w[0] * PR::INV) & 0x3fffffffu;
      c = 0;
      for (int i = 0; i < N; i++) { c += (uint64_t)PR::p(i) * qq + w[i]; w[i] = (uint32_t)c; c >>= 32; }
      w[N] += (uint32_t)c;
      FpN<PR> red;
      for (int i = 0; i < N; i++) red.v[i] = (w[i] >> 30) | (w[i + 1] << 2);
      fpn_cond_sub(red, 0);
      uint32_t* du2 = row ? nv : nu;
      for (int i = 0; i < N; i++) du2[i] = red.v[i];
    }
    for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
  }
  FpN<PR> out;
  for (int i = 0; i < N; i++) out.v[i] = v[i];
  return out;
}

// ---- G1: y^2 = x^3 + b, a = 0 (both curves), XYZZ accumulators as in ec.cuh ---------------------------------
template <class F>
struct alignas(16) G1AffineT {
  F x, y;  // Montgomery; all-zero = infinity ((0, 0) is on neither curve: b != 0)
  __host__ __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
template <class F>
struct alignas(16) G1XyzzT {
  F X, Y, ZZ, ZZZ;
  __host__ __device__ __forceinline__ bool is_inf() const { return ZZ.is_zero(); }
  __host__ __device__ __forceinline__ static G1XyzzT inf() { G1XyzzT r; r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero(); return r; }
};

template <class F>
__host__ __device__ inline G1XyzzT<F> xyzz_dbl(const G1XyzzT<F>& p) {
  if (p.is_inf()) return p;
  F U = fp_dbl(p.Y), V = fp_sqr(U), W = fp_mul(U, V), S = fp_mul(p.X, V), XX = fp_sqr(p.X);
  F M = fp_add(fp_dbl(XX), XX);
  G1XyzzT<F> r;
  r.X = fp_sub(fp_sqr(M), fp_dbl(S));
  r.Y = fp_mul_sub(M, fp_sub(S, r.X), W, p.Y);
  r.ZZ = fp_mul(V, p.ZZ);
  r.ZZZ = fp_mul(W, p.ZZZ);
  return r;
}
template <class F>
__host__ __device__ inline void xyzz_add_mixed(G1XyzzT<F>& acc, const F& x2, const F& y2_in, bool neg) {
  F y2 = neg ? fp_neg(y2_in) : y2_in;
  if (acc.is_inf()) { acc.X = x2; acc.Y = y2; acc.ZZ = F::one(); acc.ZZZ = F::one(); return; }
  F U2 = fp_mul(x2, acc.ZZ), S2 = fp_mul(y2, acc.ZZZ);
  F P = fp_sub(U2, acc.X), Rr = fp_sub(S2, acc.Y);
  if (P.is_zero()) {
    if (Rr.is_zero()) { G1XyzzT<F> q; q.X = x2; q.Y = y2; q.ZZ = F::one(); q.ZZZ = F::one(); acc = xyzz_dbl(q); }
    else acc = G1XyzzT<F>::inf();
    return;
  }
  F PP = fp_sqr(P), PPP = fp_mul(P, PP), Qq = fp_mul(acc.X, PP);
  F X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  F Y3 = fp_mul_sub(Rr, fp_sub(Qq, X3), acc.Y, PPP);
  acc.X = X3; acc.Y = Y3;
  acc.ZZ = fp_mul(acc.ZZ, PP);
  acc.ZZZ = fp_mul(acc.ZZZ, PPP);
}
template <class F>
__host__ __device__ inline void xyzz_add(G1XyzzT<F>& acc, const G1XyzzT<F>& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc = q; return; }
  F U1 = fp_mul(acc.X, q.ZZ), U2 = fp_mul(q.X, acc.ZZ), S1 = fp_mul(acc.Y, q.ZZZ), S2 = fp_mul(q.Y, acc.ZZZ);
  F P = fp_sub(U2, U1), Rr = fp_sub(S2, S1);
  if (P.is_zero()) {
    if (Rr.is_zero()) acc = xyzz_dbl(acc);
    else acc = G1XyzzT<F>::inf();
    return;
  }
  F PP = fp_sqr(P), PPP = fp_mul(P, PP), Qq = fp_mul(U1, PP);
  F X3 = fp_sub(fp_sub(fp_sqr(Rr), PPP), fp_dbl(Qq));
  F Y3 = fp_mul_sub(Rr, fp_sub(Qq, X3), S1, PPP);
  acc.X = X3; acc.Y = Y3;
  acc.ZZ = fp_mul(fp_mul(acc.ZZ, q.ZZ), PP);
  acc.ZZZ = fp_mul(fp_mul(acc.ZZZ, q.ZZZ), PPP);
}
template <class F>
__host__ __device__ inline G1AffineT<F> xyzz_to_affine(const G1XyzzT<F>& p) {
  G1AffineT<F> r;
  if (p.is_inf()) { r.x = F::zero(); r.y = F::zero(); return r; }
  F t = fp_inv(fp_mul(p.ZZ, p.ZZZ));
  r.x = fp_mul(p.X, fp_mul(t, p.ZZZ));
  r.y = fp_mul(p.Y, fp_mul(t, p.ZZ));
  return r;
}

typedef FpN<Bls381FqParams> Fq381;
typedef FpN<Bls377FqParams> Fq377;

}  // namespace capgpu
