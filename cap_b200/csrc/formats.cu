// On-disk / wire formats of the reference -> device layout.
//
// The reference stores its universal SRS and its proving keys as ark-serialize 0.3 `CanonicalSerialize`
// blobs (`store_data` / `load_data`, /root/reference/src/parameters.rs:557-592) and gates the Aztec
// CRS behind a SHA-256 digest before deserialising it (/root/reference/src/proof/mod.rs:98-107).
// This file parses those containers on the host and hands the bulk data to the device in the form
// it arrives in: compressed G1 points are decompressed by a kernel (msm.cu), coefficient vectors
// are uploaded once.  Grammar (little-endian throughout)  [UPSTREAM-RECALL: ark-serialize 0.3.0,
// ark-poly-commit @ cafc05e3, jf-plonk 0.1.2 @ bcd92b2c are not vendored; the layouts below are the
// derive(CanonicalSerialize) field orders of the published sources and are restated, with a writer,
// in oracle/serialize.py; the parser refuses anything that does not consume the blob exactly]:
//
//   usize / u64        8 bytes            bool   1 byte          Option<T>   1 byte tag, then T if 1
//   Vec<T>             u64 length, items  BTreeMap<K, V>  u64 length, (K, V) pairs
//   Fr                 32 bytes, canonical (non-Montgomery) value
//   G1Affine           32 bytes compressed: x, bit 255 = "y is the larger root", bit 254 = infinity
//   G2Affine           64 bytes compressed: x.c0 || x.c1, flags in the top bits of the last byte
//   DensePolynomial    Vec<Fr> (leading zero coefficients trimmed)
//
//   UniversalSrs  = ark_poly_commit::kzg10::UniversalParams:
//                   powers_of_g: Vec<G1>, powers_of_gamma_g: BTreeMap<usize, G1>, h: G2, beta_h: G2,
//                   neg_powers_of_h: BTreeMap<usize, G2>
//   ProvingKey    = sigmas: Vec<DensePolynomial>, selectors: Vec<DensePolynomial>,
//                   commit_key: Powers { powers_of_g: Vec<G1>, powers_of_gamma_g: Vec<G1> },
//                   vk: VerifyingKey, plookup_pk: Option<..> (must be None)
//   VerifyingKey  = domain_size: usize, num_inputs: usize, sigma_comms: Vec<G1>, selector_comms: Vec<G1>,
//                   k: Vec<Fr>, open_key: { g: G1, gamma_g: G1, h: G2, beta_h: G2 }, is_merged: bool,
//                   plookup_vk: Option<..> (must be None)
//   Proof         = wires_poly_comms: Vec<G1>, prod_perm_poly_comm: G1, split_quot_poly_comms: Vec<G1>,
//                   opening_proof: G1, shifted_opening_proof: G1,
//                   poly_evals: { wires_evals: Vec<Fr>, wire_sigma_evals: Vec<Fr>, perm_next_eval: Fr },
//                   plookup_proof: Option<..> (None)
#include <string.h>

#include <vector>

#include "prover.h"
#include "transcript.h"

using namespace capgpu;

namespace {

// ---- SHA-256 (FIPS 180-4) --------------------------------------------------------------------
struct Sha256 {
  uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  uint8_t buf[64];
  size_t fill = 0;
  uint64_t total = 0;
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
        0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
        0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
        0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
        0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
        0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
      uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t len) {
    total += len;
    while (len) {
      if (fill == 0 && len >= 64) { block(p); p += 64; len -= 64; continue; }
      size_t take = 64 - fill < len ? 64 - fill : len;
      memcpy(buf + fill, p, take);
      fill += take; p += take; len -= take;
      if (fill == 64) { block(buf); fill = 0; }
    }
  }
  void finish(uint8_t out[32]) {
    uint64_t bits = total * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lenb[8];
    for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lenb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
  }
};

// ---- bounds-checked reader ---------------------------------------------------------------------
struct Reader {
  const uint8_t* p;
  size_t len, off = 0;
  const uint8_t* take(size_t k, const char* what) {
    if (k > len - off) throw ArgError{what};
    const uint8_t* r = p + off;
    off += k;
    return r;
  }
  uint64_t u64(const char* what) { uint64_t v; memcpy(&v, take(8, what), 8); return v; }
  uint8_t u8(const char* what) { return *take(1, what); }
  // Vec<fixed-size item>: returns pointer to the items, sets count
  const uint8_t* vec(size_t item, uint64_t* count, const char* what) {
    uint64_t c = u64(what);
    if (c > (len - off) / item) throw ArgError{what};
    *count = c;
    return take((size_t)c * item, what);
  }
  // BTreeMap<usize, fixed-size item>
  void skip_map(size_t item, const char* what) {
    uint64_t c = u64(what);
    if (c > (len - off) / (item + 8)) throw ArgError{what};
    take((size_t)c * (item + 8), what);
  }
};

// canonical little-endian Fr bytes -> Montgomery limbs (values >= r are rejected like ark-serialize does)
void fr_from_canonical_bytes(const uint8_t* b, uint64_t out[4], const char* what) {
  uint64_t l[4];
  memcpy(l, b, 32);
  if (HFr::geq_p(l)) throw ArgError{what};
  HFr v = HFr::from_canonical(l);
  memcpy(out, v.v, 32);
}

// ark-serialize compressed G1 -> affine Montgomery x || y on the host (for the 18 vk commitments)
void g1_decompress_host(const uint8_t* b, uint64_t xy[8], const char* what) {
  uint8_t t[32];
  memcpy(t, b, 32);
  const bool y_larger = t[31] & 0x80, inf = t[31] & 0x40;
  t[31] &= 0x3f;
  if (inf) { memset(xy, 0, 64); return; }
  uint64_t xl[4];
  memcpy(xl, t, 32);
  if (HFq::geq_p(xl)) throw ArgError{what};
  HFq x = HFq::from_canonical(xl);
  HFq rhs = x.sqr() * x + HFq::from_u64(3);
  static const uint64_t EXP[4] = {0x4f082305b61f3f52ull, 0x65e05aa45a1c72a3ull, 0x6e14116da0605617ull, 0x0c19139cb84c680aull};  // (q + 1) / 4
  HFq y = rhs.pow(EXP, 4);
  if (!(y.sqr() == rhs)) throw ArgError{what};
  uint64_t yc[4], nyc[4];
  y.to_canonical(yc);
  y.neg().to_canonical(nyc);
  bool larger = false;
  for (int i = 3; i >= 0; i--)
    if (yc[i] != nyc[i]) { larger = yc[i] > nyc[i]; break; }
  if (larger != y_larger) y = y.neg();
  memcpy(xy, x.v, 32);
  memcpy(xy + 4, y.v, 32);
}

}  // namespace

extern "C" int capgpu_sha256(const uint8_t* data, size_t len, uint8_t out[32]) {
  if ((!data && len) || !out) return CAPGPU_ERR_ARG;
  Sha256 s;
  s.update(data, len);
  s.finish(out);
  return CAPGPU_OK;
}

extern "C" int capgpu_srs_load_serialized(capgpu_ctx* ctx, const uint8_t* bytes, size_t len, const uint8_t* expect_sha256,
                                          size_t max_points, int window_bits, capgpu_srs** out) {
  if (!ctx || !bytes || !out) return CAPGPU_ERR_ARG;
  *out = nullptr;
  const uint8_t* pts = nullptr;
  uint64_t count = 0;
  int rc = guarded(ctx, [&] {
    if (expect_sha256) {  // the integrity gate of src/proof/mod.rs:98-107
      uint8_t dig[32];
      capgpu_sha256(bytes, len, dig);
      CAPGPU_REQUIRE(memcmp(dig, expect_sha256, 32) == 0, "Mismatched sha256sum digest, file might be corrupted!");
    }
    Reader r{bytes, len};
    pts = r.vec(32, &count, "UniversalSrs: truncated powers_of_g");
    r.skip_map(32, "UniversalSrs: truncated powers_of_gamma_g");
    r.take(64, "UniversalSrs: truncated h");
    r.take(64, "UniversalSrs: truncated beta_h");
    r.skip_map(64, "UniversalSrs: truncated neg_powers_of_h");
    CAPGPU_REQUIRE(r.off == len, "UniversalSrs: trailing bytes (not an ark-serialize UniversalParams blob?)");
    CAPGPU_REQUIRE(count >= 1, "UniversalSrs: empty powers_of_g");
  });
  if (rc != CAPGPU_OK) return rc;
  if (max_points && max_points < count) count = max_points;  // `trim` to the supported degree
  return capgpu_srs_upload_compressed(ctx, pts, (size_t)count, window_bits, out);
}

extern "C" int capgpu_pk_load_serialized(capgpu_ctx* ctx, const uint8_t* bytes, size_t len, size_t* consumed, capgpu_pk** out) {
  if (!ctx || !bytes || !out) return CAPGPU_ERR_ARG;
  *out = nullptr;
  std::vector<uint64_t> sel, sig;
  uint64_t k[20], sel_comms[13 * 8], sig_comms[5 * 8];
  const uint8_t* powers = nullptr;
  uint64_t n_powers = 0, domain = 0, num_inputs = 0;
  unsigned log_n = 0;
  int rc = guarded(ctx, [&] {
    Reader r{bytes, len};
    auto polys = [&](uint64_t expect, std::vector<const uint8_t*>& ptr, std::vector<uint64_t>& lens, const char* what) {
      uint64_t c = r.u64(what);
      if (c != expect) throw ArgError{what};
      for (uint64_t i = 0; i < c; i++) {
        uint64_t l;
        ptr.push_back(r.vec(32, &l, what));
        lens.push_back(l);
      }
    };
    std::vector<const uint8_t*> sig_p, sel_p;
    std::vector<uint64_t> sig_l, sel_l;
    polys(5, sig_p, sig_l, "ProvingKey: expected 5 sigma polynomials");
    polys(13, sel_p, sel_l, "ProvingKey: expected 13 selector polynomials");
    powers = r.vec(32, &n_powers, "ProvingKey: truncated commit_key.powers_of_g");
    uint64_t n_gamma;
    r.vec(32, &n_gamma, "ProvingKey: truncated commit_key.powers_of_gamma_g");
    domain = r.u64("ProvingKey: truncated vk.domain_size");
    num_inputs = r.u64("ProvingKey: truncated vk.num_inputs");
    CAPGPU_REQUIRE(domain >= 4 && domain <= ((uint64_t)1 << 17) && (domain & (domain - 1)) == 0, "ProvingKey: domain size is not 2^2 .. 2^17");
    while (((uint64_t)1 << log_n) < domain) log_n++;
    uint64_t c;
    const uint8_t* sc = r.vec(32, &c, "ProvingKey: truncated vk.sigma_comms");
    CAPGPU_REQUIRE(c == 5, "ProvingKey: expected 5 sigma commitments");
    for (int i = 0; i < 5; i++) g1_decompress_host(sc + 32 * i, sig_comms + 8 * i, "ProvingKey: sigma commitment is not on the curve");
    const uint8_t* qc = r.vec(32, &c, "ProvingKey: truncated vk.selector_comms");
    CAPGPU_REQUIRE(c == 13, "ProvingKey: expected 13 selector commitments (TurboPlonk)");
    for (int i = 0; i < 13; i++) g1_decompress_host(qc + 32 * i, sel_comms + 8 * i, "ProvingKey: selector commitment is not on the curve");
    const uint8_t* kk = r.vec(32, &c, "ProvingKey: truncated vk.k");
    CAPGPU_REQUIRE(c == 5, "ProvingKey: expected 5 coset representatives");
    for (int i = 0; i < 5; i++) fr_from_canonical_bytes(kk + 32 * i, k + 4 * i, "ProvingKey: k is not a canonical field element");
    r.take(32 + 32 + 64 + 64, "ProvingKey: truncated vk.open_key");
    r.u8("ProvingKey: truncated vk.is_merged");
    CAPGPU_REQUIRE(r.u8("ProvingKey: truncated vk.plookup_vk") == 0, "ProvingKey: plookup keys are not supported (CAP uses TurboPlonk)");
    CAPGPU_REQUIRE(r.u8("ProvingKey: truncated plookup_pk") == 0, "ProvingKey: plookup keys are not supported (CAP uses TurboPlonk)");
    CAPGPU_REQUIRE(consumed || r.off == len, "ProvingKey: trailing bytes");
    if (consumed) *consumed = r.off;
    CAPGPU_REQUIRE(n_powers >= domain + 3, "ProvingKey: commit key shorter than domain size + 3");
    sel.assign(13 * domain * 4, 0);
    sig.assign(5 * domain * 4, 0);
    for (int s = 0; s < 13; s++) {
      CAPGPU_REQUIRE(sel_l[s] <= domain, "ProvingKey: selector polynomial longer than the domain");
      for (uint64_t j = 0; j < sel_l[s]; j++)
        fr_from_canonical_bytes(sel_p[s] + 32 * j, &sel[((size_t)s * domain + j) * 4], "ProvingKey: non-canonical selector coefficient");
    }
    for (int s = 0; s < 5; s++) {
      CAPGPU_REQUIRE(sig_l[s] <= domain, "ProvingKey: sigma polynomial longer than the domain");
      for (uint64_t j = 0; j < sig_l[s]; j++)
        fr_from_canonical_bytes(sig_p[s] + 32 * j, &sig[((size_t)s * domain + j) * 4], "ProvingKey: non-canonical sigma coefficient");
    }
  });
  if (rc != CAPGPU_OK) return rc;
  capgpu_srs* srs = nullptr;
  rc = capgpu_srs_upload_compressed(ctx, powers, (size_t)n_powers, 0, &srs);
  if (rc != CAPGPU_OK) return rc;
  capgpu_pk* pk = nullptr;
  rc = pk_create_from_coefficients(ctx, srs, log_n, (size_t)num_inputs, sel.data(), sig.data(), k, sel_comms, sig_comms, &pk);
  if (rc != CAPGPU_OK) { capgpu_srs_destroy(srs); return rc; }
  pk->owned_srs = srs;
  *out = pk;
  return CAPGPU_OK;
}

extern "C" int capgpu_proof_serialize(const capgpu_proof* proof, uint8_t* out, size_t cap, size_t* len) {
  if (!proof || !len) return CAPGPU_ERR_ARG;
  const size_t need = (8 + 5 * 32) + 32 + (8 + 5 * 32) + 32 + 32 + (8 + 5 * 32) + (8 + 4 * 32) + 32 + 1;
  *len = need;
  if (!out) return CAPGPU_OK;  // size query
  if (cap < need) return CAPGPU_ERR_ARG;
  uint8_t* p = out;
  auto put_u64 = [&](uint64_t v) { memcpy(p, &v, 8); p += 8; };
  auto put_g1 = [&](const uint64_t* xy) { g1_compress(xy, p); p += 32; };
  auto put_fr = [&](const uint64_t* v) { fr_to_le_bytes(HFr::from_limbs(v), p); p += 32; };
  put_u64(5);
  for (int i = 0; i < 5; i++) put_g1(proof->wires_poly_comms[i]);
  put_g1(proof->prod_perm_poly_comm);
  put_u64(5);
  for (int i = 0; i < 5; i++) put_g1(proof->split_quot_poly_comms[i]);
  put_g1(proof->opening_proof);
  put_g1(proof->shifted_opening_proof);
  put_u64(5);
  for (int i = 0; i < 5; i++) put_fr(proof->wires_evals[i]);
  put_u64(4);
  for (int i = 0; i < 4; i++) put_fr(proof->wire_sigma_evals[i]);
  put_fr(proof->perm_next_eval);
  *p++ = 0;  // plookup_proof: None
  return CAPGPU_OK;
}

// The prover's blinding scalars from the raw words of the caller's RNG, the way ark-ff 0.3
// `Fp256::rand` consumes them: four next_u64 per attempt (limb 0 first), top two bits of the last limb
// cleared, the attempt rejected if the value is >= r; the accepted limbs ARE the Montgomery
// representation.  words: n_words u64 in draw order; returns the number of words consumed through
// *used.  (A replay harness records the words with a wrapping RNG and needs no knowledge of the
// prover's internals; pinned by the coset-representative known answers in tests/test_oracle_hash.py.)
extern "C" int capgpu_fr_rand_from_words(const uint64_t* words, size_t n_words, uint64_t* out, size_t n_out, size_t* used) {
  if (!words || !out) return CAPGPU_ERR_ARG;
  size_t w = 0;
  for (size_t i = 0; i < n_out; i++) {
    for (;;) {
      if (w + 4 > n_words) return CAPGPU_ERR_ARG;
      uint64_t l[4] = {words[w], words[w + 1], words[w + 2], words[w + 3] & (~0ull >> 2)};
      w += 4;
      if (!HFr::geq_p(l)) { memcpy(out + 4 * i, l, 32); break; }
    }
  }
  if (used) *used = w;
  return CAPGPU_OK;
}
