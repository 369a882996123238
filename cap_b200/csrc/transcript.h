// Host-side Fiat-Shamir transcript: jf-plonk 0.1.2 `SolidityTranscript` (Keccak-256), the
// transcript type the reference instantiates at /root/reference/src/proof/transfer.rs:44,181
// (mint.rs:113, freeze.rs:151).  [UPSTREAM-RECALL: jellyfish @ bcd92b2c is not vendored; byte
// conventions restated from the published source and the CAP on-chain verifier; twin of
// oracle/transcript.py.  Everything that is convention rather than mathematics lives here so
// it can be corrected without touching a kernel; the round-level ABI bypasses it entirely.]
//
//   * append-only byte vector + 64-byte state (zero-initialised); labels are ignored;
//   * Fr: 32-byte little-endian canonical value (ark-serialize);
//   * G1: ark-serialize compressed form: x little-endian, bit 7 of byte 31 set iff y > -y,
//     bit 6 set (x = 0) for infinity;
//   * challenge: state = keccak256(state|transcript|0) | keccak256(state|transcript|1);
//     value = Fr::from_le_bytes_mod_order(state[0..48]); the transcript vector is kept.
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

#include "hostfp.h"

namespace capgpu {

// ---- Keccak-256 (original padding 0x01, rate 136) ------------------------------------------
inline uint64_t keccak_rol(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

inline void keccak_f1600(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
      0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
      0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
      0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int round = 0; round < 24; round++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ keccak_rol(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) st[i] ^= d[i % 5];
    // rho + pi: lane (x, y) at index x + 5y moves to (y, 2x + 3y)
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = keccak_rol(st[x + 5 * y], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) st[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    st[0] ^= RC[round];
  }
}

inline void keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
  uint64_t st[25];
  memset(st, 0, sizeof st);
  const size_t rate = 136;
  size_t off = 0;
  auto absorb = [&](const uint8_t* blk) {
    for (size_t i = 0; i < rate / 8; i++) {
      uint64_t lane;
      memcpy(&lane, blk + 8 * i, 8);  // little-endian host
      st[i] ^= lane;
    }
    keccak_f1600(st);
  };
  while (len - off >= rate) { absorb(data + off); off += rate; }
  uint8_t last[136];
  memset(last, 0, sizeof last);
  memcpy(last, data + off, len - off);
  last[len - off] ^= 0x01;
  last[rate - 1] ^= 0x80;
  absorb(last);
  memcpy(out, st, 32);
}

// ---- serialisation helpers -----------------------------------------------------------------
inline void fr_to_le_bytes(const HFr& x, uint8_t out[32]) {
  uint64_t c[4];
  x.to_canonical(c);
  memcpy(out, c, 32);
}

// xy: affine point in ABI layout (x||y Montgomery, zeros = infinity) -> 32 compressed bytes
inline void g1_compress(const uint64_t xy[8], uint8_t out[32]) {
  bool inf = true;
  for (int i = 0; i < 8; i++) inf = inf && xy[i] == 0;
  if (inf) { memset(out, 0, 32); out[31] |= 0x40; return; }
  HFq x = HFq::from_limbs(xy), y = HFq::from_limbs(xy + 4);
  uint64_t xc[4], yc[4], nyc[4];
  x.to_canonical(xc);
  y.to_canonical(yc);
  y.neg().to_canonical(nyc);
  memcpy(out, xc, 32);
  bool y_larger = false;
  for (int i = 3; i >= 0; i--) {
    if (yc[i] != nyc[i]) { y_larger = yc[i] > nyc[i]; break; }
  }
  if (y_larger) out[31] |= 0x80;
}

struct SolidityTranscript {
  std::vector<uint8_t> transcript;
  uint8_t state[64];

  SolidityTranscript() { memset(state, 0, sizeof state); }
  void append_message(const uint8_t* msg, size_t len) { transcript.insert(transcript.end(), msg, msg + len); }
  void append_u64_le(uint64_t v) { uint8_t b[8]; memcpy(b, &v, 8); append_message(b, 8); }
  void append_field(const HFr& x) { uint8_t b[32]; fr_to_le_bytes(x, b); append_message(b, 32); }
  void append_commitment(const uint64_t xy[8]) { uint8_t b[32]; g1_compress(xy, b); append_message(b, 32); }

  HFr get_and_append_challenge() {
    std::vector<uint8_t> buf(64 + transcript.size() + 1);
    memcpy(buf.data(), state, 64);
    if (!transcript.empty()) memcpy(buf.data() + 64, transcript.data(), transcript.size());
    uint8_t h0[32], h1[32];
    buf[buf.size() - 1] = 0;
    keccak256(buf.data(), buf.size(), h0);
    buf[buf.size() - 1] = 1;
    keccak256(buf.data(), buf.size(), h1);
    memcpy(state, h0, 32);
    memcpy(state + 32, h1, 32);
    return HFr::from_le_bytes_mod_order(state, 48);
  }
};

}  // namespace capgpu
