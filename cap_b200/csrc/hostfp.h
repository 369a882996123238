// Host-side 4 x 64-bit Montgomery arithmetic for the O(1)-per-proof scalar work of the prover
// (Fiat-Shamir challenges, linearisation scalars, point compression).  The heavy loops all
// run on the GPU; this mirrors what jf-plonk itself does on the CPU between rounds.
#pragma once
#include <stdint.h>
#include <string.h>

namespace capgpu {

typedef unsigned __int128 u128;

struct HFrParams {
  static constexpr uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t INV = 0xc2e1f593efffffffull;
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
};
struct HFqParams {
  static constexpr uint64_t P[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  static constexpr uint64_t INV = 0x87d20782e4866389ull;
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
  static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
};

template <class PR>
struct HFp {
  uint64_t v[4];

  static HFp zero() { HFp r; memset(r.v, 0, sizeof r.v); return r; }
  static HFp one() { HFp r; memcpy(r.v, PR::ONE, sizeof r.v); return r; }
  static HFp from_limbs(const uint64_t* l) { HFp r; memcpy(r.v, l, sizeof r.v); return r; }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const HFp& o) const { return memcmp(v, o.v, sizeof v) == 0; }

  static bool geq_p(const uint64_t* a) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > PR::P[i]) return true;
      if (a[i] < PR::P[i]) return false;
    }
    return true;
  }
  static void sub_p(uint64_t* a) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)a[i] - PR::P[i] - borrow;
      a[i] = (uint64_t)t;
      borrow = (t >> 64) & 1;
    }
  }
  HFp operator+(const HFp& o) const {
    HFp r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    if (geq_p(r.v)) sub_p(r.v);
    return r;
  }
  HFp operator-(const HFp& o) const {
    HFp r;
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
      u128 t = (u128)v[i] - o.v[i] - borrow;
      r.v[i] = (uint64_t)t;
      borrow = (t >> 64) & 1;
    }
    if (borrow) {
      u128 c = 0;
      for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + PR::P[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  HFp neg() const { return zero() - *this; }
  // Montgomery product (CIOS)
  HFp operator*(const HFp& o) const {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
      u128 c = 0;
      for (int j = 0; j < 4; j++) { c += (u128)v[j] * o.v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * PR::INV;
      c = (u128)m * PR::P[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; j++) { c += (u128)m * PR::P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
      c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    HFp r;
    memcpy(r.v, t, sizeof r.v);
    if (t[4] || geq_p(r.v)) sub_p(r.v);
    return r;
  }
  HFp sqr() const { return *this * *this; }
  HFp pow(const uint64_t* e, int limbs) const {
    HFp r = one();
    bool started = false;
    for (int i = limbs * 64 - 1; i >= 0; i--) {
      if (started) r = r.sqr();
      if ((e[i >> 6] >> (i & 63)) & 1) { r = started ? r * *this : *this; started = true; }
    }
    return r;
  }
  HFp pow_u64(uint64_t e) const { return pow(&e, 1); }
  HFp inv() const {
    uint64_t e[4];
    memcpy(e, PR::P, sizeof e);
    e[0] -= 2;
    return pow(e, 4);
  }
  // value (canonical integer limbs) -> Montgomery
  static HFp from_canonical(const uint64_t* l) { HFp a = from_limbs(l); HFp r2 = from_limbs(PR::R2); return a * r2; }
  static HFp from_u64(uint64_t x) { uint64_t l[4] = {x, 0, 0, 0}; return from_canonical(l); }
  // Montgomery -> canonical integer limbs
  void to_canonical(uint64_t* out) const {
    HFp o; memset(o.v, 0, sizeof o.v); o.v[0] = 1;
    HFp r = *this * o;
    memcpy(out, r.v, sizeof r.v);
  }
  // ark-ff `from_le_bytes_mod_order`: interpret little-endian bytes as an integer, reduce mod p.
  static HFp from_le_bytes_mod_order(const uint8_t* bytes, size_t len) {
    // Horner over bytes from the most significant end: acc = acc * 256 + byte
    HFp acc = zero();
    HFp c256 = from_u64(256);
    for (size_t i = len; i-- > 0;) acc = acc * c256 + from_u64(bytes[i]);
    return acc;
  }
};

typedef HFp<HFrParams> HFr;
typedef HFp<HFqParams> HFq;

}  // namespace capgpu
