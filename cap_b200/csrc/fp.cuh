// 254-bit Montgomery field arithmetic on 8 x 32-bit limbs for BN254 Fr and Fq (sm_100a).
//
// Replaces ark-ff 0.3.0 `Fp256<FrParameters>` / `Fp256<FqParameters>` (the field types the
// reference fixes at /root/reference/src/config.rs:78-83).  Same value representation:
// little-endian limbs of a*R mod p, R = 2^256, so host buffers of arkworks field elements
// are usable byte-for-byte (4 x u64 LE == 8 x u32 LE).
//
// Multiplication is the even/odd column-split CIOS: the partial products a[j]*b[i] are
// accumulated as 64-bit wide multiply-adds into two interleaved accumulators (even and odd
// limb alignment) so that every mad.lo.cc/madc.hi.cc pair lowers to one IMAD.WIDE.U32 with
// carry-in/out and no separate carry-fold instructions on the ALU pipe.
//
// Measured on B200 (profiles/README.md): IMAD.WIDE.U32 issues at 64 lanes/clk/SM only when its
// multiplicands sit in the operand-reuse cache; with the distinct register operands of a real
// 8 x 8 limb product it sustains ~2.6x less, and this CIOS (half of whose multiply-adds take the
// modulus as an immediate) tops out at ~68 G products/s with every warp slot busy.  A reduced-
// radix variant (9 x 29-bit limbs, carry-free 64-bit column sums, 163 plain IMAD.WIDE) was
// built, proven bit-identical and measured at 44 G products/s: slower, so it is not kept.
//
// All functions keep values fully reduced in [0, p): results are bit-exact canonical
// Montgomery residues, which is what the parity tests compare.
//
// Host emulation (CAPGPU_HOST_EMU) exists ONLY so tests/cpu_emu can run these exact
// algorithms against the big-int oracle without a GPU; the product never builds with it.
#pragma once
#include <stdint.h>

#if defined(CAPGPU_HOST_EMU) && !defined(__CUDACC__)
#define __device__
#define __host__
#define __forceinline__ inline
#define CAPGPU_EMU 1
#endif

#if defined(__CUDA_ARCH__)
#define CAPGPU_DEV 1
#endif

namespace capgpu {

// ------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------
#ifdef CAPGPU_DEV
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// 64-bit wide multiply(-add) on a register pair (lo, hi)
__device__ __forceinline__ void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  // one IMAD.WIDE.U32 (mul.lo + mul.hi would lower to IMAD + IMAD.HI: two fma-pipe slots)
  asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b, starts a carry chain (no carry-in), leaves carry-out
__device__ __forceinline__ void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b + carry-in, leaves carry-out
__device__ __forceinline__ void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) = a*b + (clo,chi) + carry-in, leaves carry-out
__device__ __forceinline__ void madc_wide_cc_to(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + carry-in, ends the chain (the high word cannot overflow)
__device__ __forceinline__ void madc_wide_end(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
#else
// Host emulation of the PTX condition-code semantics (tests only).
struct EmuCC { static uint32_t& cf() { static thread_local uint32_t c = 0; return c; } };
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; EmuCC::cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + EmuCC::cf(); EmuCC::cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + EmuCC::cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; EmuCC::cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - EmuCC::cf(); EmuCC::cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - EmuCC::cf(); }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
inline void emu_mad(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi, uint32_t cin, bool cout) {
  uint64_t p = (uint64_t)a * b;
  uint64_t l = (uint64_t)(uint32_t)p + clo + cin;
  uint64_t h = (p >> 32) + chi + (l >> 32);
  lo = (uint32_t)l; hi = (uint32_t)h;
  if (cout) EmuCC::cf() = (uint32_t)(h >> 32);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, lo, hi, 0, true); }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, lo, hi, EmuCC::cf(), true); }
inline void madc_wide_cc_to(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) { emu_mad(lo, hi, a, b, clo, chi, EmuCC::cf(), true); }
inline void madc_wide_end(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, 0, 0, EmuCC::cf(), false); }
#endif

// ------------------------------------------------------------------------------------------
// field parameters (BN254; SURVEY.md App. B, checked against the oracle in tests)
// ------------------------------------------------------------------------------------------
struct FrParams {
  // r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
  static __host__ __device__ __forceinline__ constexpr uint32_t p(int i) { constexpr uint32_t t[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
  static constexpr uint32_t INV = 0xefffffffu;  // -p^-1 mod 2^32
  static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) { constexpr uint32_t t[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }  // R mod p
  static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) { constexpr uint32_t t[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}; return t[i]; }  // R^2 mod p
};
struct FqParams {
  // q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
  static __host__ __device__ __forceinline__ constexpr uint32_t p(int i) { constexpr uint32_t t[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}; return t[i]; }
  static constexpr uint32_t INV = 0xe4866389u;
  static __host__ __device__ __forceinline__ constexpr uint32_t one(int i) { constexpr uint32_t t[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}; return t[i]; }
  static __host__ __device__ __forceinline__ constexpr uint32_t r2(int i) { constexpr uint32_t t[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u}; return t[i]; }
};

template <class PR>
struct alignas(32) Fp {
  uint32_t v[8];

  __host__ __device__ __forceinline__ static Fp zero() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
  __host__ __device__ __forceinline__ static Fp one() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = PR::one(i); return r; }
  __host__ __device__ __forceinline__ static Fp r2() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = PR::r2(i); return r; }
  __host__ __device__ __forceinline__ bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= v[i]; return o == 0; }
  __host__ __device__ __forceinline__ bool operator==(const Fp& b) const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i]; return o == 0; }
  __host__ __device__ __forceinline__ bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a   (a < 2p assumed, given as 8 limbs + no overflow bit)
template <class PR>
__host__ __device__ __forceinline__ void fp_final_sub(Fp<PR>& a) {
  uint32_t t[8];
  t[0] = sub_cc(a.v[0], PR::p(0));
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = subc_cc(a.v[i], PR::p(i));
  uint32_t borrow = subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_add(const Fp<PR>& a, const Fp<PR>& b) {
  Fp<PR> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  // p < 2^254 so a + b < 2^255: no carry out of limb 7
  fp_final_sub(r);
  return r;
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_sub(const Fp<PR>& a, const Fp<PR>& b) {
  Fp<PR> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0, 0);
  uint32_t t[8];
  t[0] = add_cc(r.v[0], PR::p(0));
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = addc_cc(r.v[i], PR::p(i));
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : r.v[i];
  return r;
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_neg(const Fp<PR>& a) {
  Fp<PR> r;
  r.v[0] = sub_cc(PR::p(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) r.v[i] = subc_cc(PR::p(i), a.v[i]);
  bool z = a.is_zero();
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
  return r;
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_dbl(const Fp<PR>& a) { return fp_add(a, a); }

// ---- Montgomery multiplication ------------------------------------------------------------
// X is the accumulator whose word 0 sits at the limb position being reduced in this step,
// Y the accumulator aligned one limb higher.  After each step the roles swap (the reduced
// low word of X is dropped, its word 1 is folded into Y[0], words 2.. become the new
// high accumulator).
template <class PR>
__host__ __device__ __forceinline__ void mont_reduce_step(uint32_t* X, uint32_t* Y) {
  uint32_t m = X[0] * PR::INV;
  mad_wide_cc(Y[0], Y[1], PR::p(1), m);
  madc_wide_cc(Y[2], Y[3], PR::p(3), m);
  madc_wide_cc(Y[4], Y[5], PR::p(5), m);
  madc_wide_cc(Y[6], Y[7], PR::p(7), m);  // top limb of p < 2^30: no carry out
  mad_wide_cc(X[0], X[1], PR::p(0), m);
  madc_wide_cc(X[2], X[3], PR::p(2), m);
  madc_wide_cc(X[4], X[5], PR::p(4), m);
  madc_wide_cc(X[6], X[7], PR::p(6), m);
  Y[7] = addc(Y[7], 0);
}

template <class PR>
__host__ __device__ __forceinline__ void mont_mad_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
  // X: aligned at this row's base position; Y: previous row's X (word 0 is zero, word 1 is
  // folded into X[0], words 2..7 shift down by two to become the accumulator at base+1).
  X[0] = add_cc(X[0], Y[1]);
  madc_wide_cc_to(Y[0], Y[1], a[1], bi, Y[2], Y[3]);
  madc_wide_cc_to(Y[2], Y[3], a[3], bi, Y[4], Y[5]);
  madc_wide_cc_to(Y[4], Y[5], a[5], bi, Y[6], Y[7]);
  madc_wide_end(Y[6], Y[7], a[7], bi);
  mad_wide_cc(X[0], X[1], a[0], bi);
  madc_wide_cc(X[2], X[3], a[2], bi);
  madc_wide_cc(X[4], X[5], a[4], bi);
  madc_wide_cc(X[6], X[7], a[6], bi);
  Y[7] = addc(Y[7], 0);
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_mul(const Fp<PR>& a, const Fp<PR>& b) {
  uint32_t E[8], O[8];
  // row 0
  mul_wide(E[0], E[1], a.v[0], b.v[0]);
  mul_wide(E[2], E[3], a.v[2], b.v[0]);
  mul_wide(E[4], E[5], a.v[4], b.v[0]);
  mul_wide(E[6], E[7], a.v[6], b.v[0]);
  mul_wide(O[0], O[1], a.v[1], b.v[0]);
  mul_wide(O[2], O[3], a.v[3], b.v[0]);
  mul_wide(O[4], O[5], a.v[5], b.v[0]);
  mul_wide(O[6], O[7], a.v[7], b.v[0]);
  mont_reduce_step<PR>(E, O);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    mont_mad_row<PR>(O, E, a.v, b.v[i]);
    mont_reduce_step<PR>(O, E);
    if (i + 1 < 8) {
      mont_mad_row<PR>(E, O, a.v, b.v[i + 1]);
      mont_reduce_step<PR>(E, O);
    }
  }
  // after row 7: X = O (aligned at limb 7, word 0 == 0), Y = E (aligned at limb 8)
  Fp<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fp_final_sub(r);
  return r;
}

// ---- fused a*b + c*d -------------------------------------------------------------------------
// Both products' rows go into the same accumulators before each reduction step: 8 x (16 + 8) + 8
// wide MADs instead of 2 x 136 — one Montgomery reduction for the sum.  The sliding 9-limb window
// still cannot overflow: after row i it holds less than 3 B^(i+1) p and p < 0.19 B^8.  The result
// is < 1.6 p before the final conditional subtraction.  (Y3 of the curve formulas, gate sums.)
template <class PR>
__host__ __device__ __forceinline__ void mont_mad_row_noshift(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
  // X aligned at this row's base, Y one limb higher (already shifted by the first product's row)
  mad_wide_cc(X[0], X[1], a[0], bi);
  madc_wide_cc(X[2], X[3], a[2], bi);
  madc_wide_cc(X[4], X[5], a[4], bi);
  madc_wide_cc(X[6], X[7], a[6], bi);
  Y[7] = addc(Y[7], 0);
  mad_wide_cc(Y[0], Y[1], a[1], bi);
  madc_wide_cc(Y[2], Y[3], a[3], bi);
  madc_wide_cc(Y[4], Y[5], a[5], bi);
  madc_wide_cc(Y[6], Y[7], a[7], bi);  // no carry out: the window bound above
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_mul_add(const Fp<PR>& a, const Fp<PR>& b, const Fp<PR>& c, const Fp<PR>& d) {
  uint32_t E[8], O[8];
  mul_wide(E[0], E[1], a.v[0], b.v[0]);
  mul_wide(E[2], E[3], a.v[2], b.v[0]);
  mul_wide(E[4], E[5], a.v[4], b.v[0]);
  mul_wide(E[6], E[7], a.v[6], b.v[0]);
  mul_wide(O[0], O[1], a.v[1], b.v[0]);
  mul_wide(O[2], O[3], a.v[3], b.v[0]);
  mul_wide(O[4], O[5], a.v[5], b.v[0]);
  mul_wide(O[6], O[7], a.v[7], b.v[0]);
  mont_mad_row_noshift<PR>(E, O, c.v, d.v[0]);
  mont_reduce_step<PR>(E, O);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    mont_mad_row<PR>(O, E, a.v, b.v[i]);
    mont_mad_row_noshift<PR>(O, E, c.v, d.v[i]);
    mont_reduce_step<PR>(O, E);
    if (i + 1 < 8) {
      mont_mad_row<PR>(E, O, a.v, b.v[i + 1]);
      mont_mad_row_noshift<PR>(E, O, c.v, d.v[i + 1]);
      mont_reduce_step<PR>(E, O);
    }
  }
  Fp<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fp_final_sub(r);
  return r;
}

// a*b - c*d
template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_mul_sub(const Fp<PR>& a, const Fp<PR>& b, const Fp<PR>& c, const Fp<PR>& d) {
  return fp_mul_add(a, b, c, fp_neg(d));
}

// ---- Montgomery squaring --------------------------------------------------------------------
// Same row / reduce interleaving as fp_mul, but row i multiplies a_i by the doubled tail of a
// only:  a^2 = sum_i a_i B^i (a_i B^i + 2 sum_{j>i} a_j B^j), so row i has 8 - i products instead
// of 8 (36 instead of 64 overall; 108 wide MADs per square instead of 136).  The doubled tail is
// taken from d = 2a limb-wise: row i uses a_i, (a_{i+1} << 1), d_{i+2}, .., d_7 (the bit shifted
// out of a_{i+1} sits in d_{i+2}; a_7 < 2^30, so nothing is shifted out of the top limb).
// Skipped lanes of the shifted accumulator still propagate the carry (two adds instead of a MAD).
template <class PR, int I>
__host__ __device__ __forceinline__ void mont_sqr_row(uint32_t* X, uint32_t* Y, const uint32_t* a, const uint32_t* d) {
  const uint32_t bi = a[I];
#define CAPGPU_SQR_V(j) ((j) == I ? a[I] : ((j) == I + 1 ? (a[(j) & 7] << 1) : d[(j) & 7]))
  X[0] = add_cc(X[0], Y[1]);
  if (1 >= I) madc_wide_cc_to(Y[0], Y[1], CAPGPU_SQR_V(1), bi, Y[2], Y[3]);
  else { Y[0] = addc_cc(Y[2], 0); Y[1] = addc_cc(Y[3], 0); }
  if (3 >= I) madc_wide_cc_to(Y[2], Y[3], CAPGPU_SQR_V(3), bi, Y[4], Y[5]);
  else { Y[2] = addc_cc(Y[4], 0); Y[3] = addc_cc(Y[5], 0); }
  if (5 >= I) madc_wide_cc_to(Y[4], Y[5], CAPGPU_SQR_V(5), bi, Y[6], Y[7]);
  else { Y[4] = addc_cc(Y[6], 0); Y[5] = addc_cc(Y[7], 0); }
  madc_wide_end(Y[6], Y[7], CAPGPU_SQR_V(7), bi);
  // even multiplicand limbs j >= I: the chain starts at the first of them
  if (I <= 6) {
    constexpr int J0 = (I + 1) & ~1;  // first even j >= I
    mad_wide_cc(X[J0], X[J0 + 1], CAPGPU_SQR_V(J0), bi);
    if (J0 + 2 <= 6) madc_wide_cc(X[J0 + 2], X[J0 + 3], CAPGPU_SQR_V(J0 + 2), bi);
    if (J0 + 4 <= 6) madc_wide_cc(X[J0 + 4], X[J0 + 5], CAPGPU_SQR_V(J0 + 4), bi);
    if (J0 + 6 <= 6) madc_wide_cc(X[J0 + 6], X[J0 + 7], CAPGPU_SQR_V(J0 + 6), bi);
    Y[7] = addc(Y[7], 0);
  }
#undef CAPGPU_SQR_V
}

template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_sqr(const Fp<PR>& a) {
  uint32_t E[8], O[8], d[8];
  d[0] = a.v[0] << 1;
#pragma unroll
  for (int j = 1; j < 8; j++) d[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
  // row 0: a_0 * [a_0, a_1 << 1, d_2 .. d_7]
  const uint32_t a1s = a.v[1] << 1;
  mul_wide(E[0], E[1], a.v[0], a.v[0]);
  mul_wide(E[2], E[3], d[2], a.v[0]);
  mul_wide(E[4], E[5], d[4], a.v[0]);
  mul_wide(E[6], E[7], d[6], a.v[0]);
  mul_wide(O[0], O[1], a1s, a.v[0]);
  mul_wide(O[2], O[3], d[3], a.v[0]);
  mul_wide(O[4], O[5], d[5], a.v[0]);
  mul_wide(O[6], O[7], d[7], a.v[0]);
  mont_reduce_step<PR>(E, O);
  mont_sqr_row<PR, 1>(O, E, a.v, d);
  mont_reduce_step<PR>(O, E);
  mont_sqr_row<PR, 2>(E, O, a.v, d);
  mont_reduce_step<PR>(E, O);
  mont_sqr_row<PR, 3>(O, E, a.v, d);
  mont_reduce_step<PR>(O, E);
  mont_sqr_row<PR, 4>(E, O, a.v, d);
  mont_reduce_step<PR>(E, O);
  mont_sqr_row<PR, 5>(O, E, a.v, d);
  mont_reduce_step<PR>(O, E);
  mont_sqr_row<PR, 6>(E, O, a.v, d);
  mont_reduce_step<PR>(E, O);
  mont_sqr_row<PR, 7>(O, E, a.v, d);
  mont_reduce_step<PR>(O, E);
  Fp<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fp_final_sub(r);
  return r;
}

// Montgomery -> canonical integer (multiply by 1) and back (multiply by R^2).
template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_from_mont(const Fp<PR>& a) {
  Fp<PR> o = Fp<PR>::zero();
  o.v[0] = 1;
  return fp_mul(a, o);
}
template <class PR>
__host__ __device__ __forceinline__ Fp<PR> fp_to_mont(const Fp<PR>& a) { return fp_mul(a, Fp<PR>::r2()); }

// a^e for a 256-bit exponent given as 8 LE limbs (square-and-multiply, MSB first).
template <class PR>
__host__ __device__ inline Fp<PR> fp_pow(const Fp<PR>& a, const uint32_t* e) {
  Fp<PR> r = Fp<PR>::one();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? fp_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}

template <class PR>
__host__ __device__ inline Fp<PR> fp_pow_u64(const Fp<PR>& a, uint64_t e) {
  uint32_t ee[8] = {(uint32_t)e, (uint32_t)(e >> 32), 0, 0, 0, 0, 0, 0};
  return fp_pow(a, ee);
}

// Inversion by Fermat: a^(p-2) (381 dependent products; kept as the cross-check of fp_inv).
template <class PR>
__host__ __device__ inline Fp<PR> fp_inv_fermat(const Fp<PR>& a) {
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = PR::p(i);
  e[0] -= 2;  // p is odd and p[0] >= 2 for both fields
  return fp_pow(a, e);
}

// x >>= 1 with `top` shifted into bit 255
__host__ __device__ __forceinline__ void limbs_shr1(uint32_t* x, uint32_t top) {
#pragma unroll
  for (int i = 0; i < 7; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
  x[7] = (x[7] >> 1) | (top << 31);
}

// x = (x even ? x : x + p) / 2 for x in [0, p)
template <class PR>
__host__ __device__ __forceinline__ void fp_halve(uint32_t* x) {
  uint32_t odd = x[0] & 1u;
  uint32_t t[8];
  t[0] = add_cc(x[0], odd ? PR::p(0) : 0u);
#pragma unroll
  for (int i = 1; i < 8; i++) t[i] = addc_cc(x[i], odd ? PR::p(i) : 0u);
  uint32_t carry = addc(0, 0);
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = t[i];
  limbs_shr1(x, carry);
}

// Plain binary extended Euclidean inversion (the algorithm ark-ff's Fp256::inverse uses; kept as
// the cross-check of fp_inv): ~750 shift/subtract steps on 8 limbs instead of 381 dependent Montgomery products,
// which matters where a single thread inverts on the critical path (MSM result -> affine,
// grand-product scan).  b starts at R^2 so the result is the Montgomery form of a^-1.
// inv(0) = 0 (callers treat zero explicitly).
template <class PR>
__host__ __device__ inline Fp<PR> fp_inv_euclid(const Fp<PR>& a) {
  if (a.is_zero()) return a;
  uint32_t u[8], v[8];
  Fp<PR> b = Fp<PR>::r2(), c = Fp<PR>::zero();
#pragma unroll
  for (int i = 0; i < 8; i++) { u[i] = a.v[i]; v[i] = PR::p(i); }
  for (;;) {
    uint32_t u_hi = 0, v_hi = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) { u_hi |= u[i]; v_hi |= v[i]; }
    if ((u[0] == 1u && u_hi == 0u) || (v[0] == 1u && v_hi == 0u)) break;
    while ((u[0] & 1u) == 0u) { limbs_shr1(u, 0); fp_halve<PR>(b.v); }
    while ((v[0] & 1u) == 0u) { limbs_shr1(v, 0); fp_halve<PR>(c.v); }
    // t = u - v
    uint32_t t[8];
    t[0] = sub_cc(u[0], v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(u[i], v[i]);
    uint32_t borrow = subc(0, 0);
    if (borrow == 0u) {  // u >= v
#pragma unroll
      for (int i = 0; i < 8; i++) u[i] = t[i];
      b = fp_sub(b, c);
    } else {
      v[0] = sub_cc(v[0], u[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) v[i] = subc_cc(v[i], u[i]);
      c = fp_sub(c, b);
    }
  }
  uint32_t u_hi = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) u_hi |= u[i];
  return (u[0] == 1u && u_hi == 0u) ? b : c;
}


// ---- inversion: binary GCD with batched steps ------------------------------------------------
// The value a single thread computes on the critical path of every MSM result (XYZZ -> affine) and
// of the grand-product scan.  ark-ff's Fp256::inverse is the plain binary extended Euclid
// (fp_inv_euclid above: ~750 data-dependent shift/subtract rounds on 8 limbs, measured 256 k cycles
// = 130 us for one warp on B200).  Here the same GCD runs in 17 outer rounds (T. Pornin, "Optimized
// Binary GCD for Modular Inversion", 2020): each round takes 62-bit approximations of a and b (the
// low 30 bits exactly, the top 32 bits of the longer of the two), runs 30 branch-free
// shift/subtract steps on them while recording the transition matrix (f0 g0; f1 g1), |f|,|g| <= 2^30,
// then applies the matrix once to the full-width (a, b) -- exactly divisible by 2^30 -- and to
// (u, v) modulo p with a 30-bit Montgomery-style division.  Invariants a*R^2 = u*y, b*R^2 = v*y
// (mod p) hold throughout, so when a reaches 0, b = 1 and v = R^2 / y: the Montgomery form of the
// inverse of the value y represents.  inv(0) = 0 (callers treat zero explicitly).  The result is the
// unique canonical residue, identical to ark-ff's.
__host__ __device__ __forceinline__ int clz32(uint32_t x) {
#ifdef CAPGPU_DEV
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

// out = |a*f + b*g| / 2^30 (exact), neg = sign of a*f + b*g; a, b < 2^255, |f|, |g| <= 2^30
__host__ __device__ __forceinline__ void gcd_lincomb(const uint32_t* a, const uint32_t* b, int32_t f, int32_t g, uint32_t* out, bool& neg) {
  const uint32_t fa = (uint32_t)(f < 0 ? -f : f), ga = (uint32_t)(g < 0 ? -g : g);
  uint32_t t1[9], t2[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)a[i] * fa; t1[i] = (uint32_t)c; c >>= 32; }
  t1[8] = (uint32_t)c;
  c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)b[i] * ga; t2[i] = (uint32_t)c; c >>= 32; }
  t2[8] = (uint32_t)c;
  // two's complement over 288 bits (|value| < 2^286)
  const uint32_t mf = f < 0 ? 0xffffffffu : 0u, mg = g < 0 ? 0xffffffffu : 0u;
  uint32_t t[9];
  uint64_t c1 = mf & 1u, c2 = mg & 1u, cs = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    c1 += (uint64_t)(t1[i] ^ mf); c2 += (uint64_t)(t2[i] ^ mg);
    cs += (uint64_t)(uint32_t)c1 + (uint32_t)c2;
    t[i] = (uint32_t)cs;
    c1 >>= 32; c2 >>= 32; cs >>= 32;
  }
  neg = (t[8] >> 31) != 0;
  const uint32_t mn = neg ? 0xffffffffu : 0u;
  uint64_t cn = mn & 1u;
#pragma unroll
  for (int i = 0; i < 9; i++) { cn += (uint64_t)(t[i] ^ mn); t[i] = (uint32_t)cn; cn >>= 32; }
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = (t[i] >> 30) | (t[i + 1] << 2);
}

// out = (u*f + v*g) / 2^30 mod p for u, v in [0, p), |f| + |g| <= 2^30
template <class PR>
__host__ __device__ __forceinline__ void gcd_lincomb_mod(const uint32_t* u, const uint32_t* v, int32_t f, int32_t g, uint32_t* out) {
  // negative coefficients: use p - x instead of x (p - 0 = p is harmless: only the residue matters)
  uint32_t uu[8], vv[8];
  {
    uint64_t bu = 0, bv = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t du = (uint64_t)PR::p(i) - u[i] - bu; bu = (du >> 32) & 1u;
      uint64_t dv = (uint64_t)PR::p(i) - v[i] - bv; bv = (dv >> 32) & 1u;
      uu[i] = f < 0 ? (uint32_t)du : u[i];
      vv[i] = g < 0 ? (uint32_t)dv : v[i];
    }
  }
  const uint32_t fa = (uint32_t)(f < 0 ? -f : f), ga = (uint32_t)(g < 0 ? -g : g);
  uint32_t t[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {  // each product < 2^62: the running sum fits 64 bits
    c += (uint64_t)uu[i] * fa + (uint64_t)vv[i] * ga;
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  t[8] = (uint32_t)c;  // total < 2^30 * (p + 1) < 2^285
  // add q*p with q = -t/p mod 2^30: the sum is divisible by 2^30 and < 2^31 * p
  const uint32_t q = (t[0] * PR::INV) & 0x3fffffffu;
  c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { c += (uint64_t)PR::p(i) * q + t[i]; t[i] = (uint32_t)c; c >>= 32; }
  t[8] += (uint32_t)c;
  Fp<PR> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 30) | (t[i + 1] << 2);
  fp_final_sub(r);  // < 2p before
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = r.v[i];
}

template <class PR>
__host__ __device__ inline Fp<PR> fp_inv(const Fp<PR>& y) {
  if (y.is_zero()) return y;
  uint32_t a[8], b[8], u[8], v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = y.v[i]; b[i] = PR::p(i); u[i] = PR::r2(i); v[i] = 0; }
  for (int round = 0; round < 18; round++) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= a[i];
    if (nz == 0) break;
    // n = max(len(a), len(b), 62); approximations: bits [n-32, n) above the low 30 bits
    uint32_t top = a[1] | b[1];
    int n = 32;
#pragma unroll
    for (int i = 2; i < 8; i++) { uint32_t w = a[i] | b[i]; if (w) { top = w; n = 32 * i; } }
    n += 32 - clz32(top);
    if (n < 62) n = 62;
    const int s = n - 32, q = s >> 5, r = s & 31;
    uint32_t alo = 0, ahi = 0, blo = 0, bhi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i == q) { alo = a[i]; blo = b[i]; ahi = i + 1 < 8 ? a[i + 1] : 0u; bhi = i + 1 < 8 ? b[i + 1] : 0u; }
    }
    const uint32_t atop = r ? (alo >> r) | (ahi << (32 - r)) : alo;
    const uint32_t btop = r ? (blo >> r) | (bhi << (32 - r)) : blo;
    uint64_t xa = ((uint64_t)atop << 30) | (a[0] & 0x3fffffffu);
    uint64_t xb = ((uint64_t)btop << 30) | (b[0] & 0x3fffffffu);
    int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 6
    for (int j = 0; j < 30; j++) {
      // both differences and the plain halving are formed before the parity / order of (xa, xb) is looked at:
      // the dependent chain per step is subtract -> shift -> select
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
    uint32_t na[8], nb[8], nu[8], nv[8];
    bool nega, negb;
    gcd_lincomb(a, b, f0, g0, na, nega);
    gcd_lincomb(a, b, f1, g1, nb, negb);
    if (nega) { f0 = -f0; g0 = -g0; }
    if (negb) { f1 = -f1; g1 = -g1; }
    gcd_lincomb_mod<PR>(u, v, f0, g0, nu);
    gcd_lincomb_mod<PR>(u, v, f1, g1, nv);
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
  }
  Fp<PR> out;
#pragma unroll
  for (int i = 0; i < 8; i++) out.v[i] = v[i];
  return out;
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace capgpu
