// Exposes the host-side C++ helpers of the library (hostfp.h, transcript.h) to the CPU tests.
#include "../../cap_b200/csrc/transcript.h"
using namespace capgpu;
extern "C" {
void emu_keccak256(const uint8_t* d, size_t n, uint8_t* out) { keccak256(d, n, out); }
void emu_hfr_mul(const uint64_t* a, const uint64_t* b, uint64_t* r) { HFr x = HFr::from_limbs(a) * HFr::from_limbs(b); memcpy(r, x.v, 32); }
void emu_hfq_mul(const uint64_t* a, const uint64_t* b, uint64_t* r) { HFq x = HFq::from_limbs(a) * HFq::from_limbs(b); memcpy(r, x.v, 32); }
void emu_hfr_add(const uint64_t* a, const uint64_t* b, uint64_t* r) { HFr x = HFr::from_limbs(a) + HFr::from_limbs(b); memcpy(r, x.v, 32); }
void emu_hfr_sub(const uint64_t* a, const uint64_t* b, uint64_t* r) { HFr x = HFr::from_limbs(a) - HFr::from_limbs(b); memcpy(r, x.v, 32); }
void emu_hfr_inv(const uint64_t* a, uint64_t* r) { HFr x = HFr::from_limbs(a).inv(); memcpy(r, x.v, 32); }
void emu_hfr_from_bytes(const uint8_t* b, size_t n, uint64_t* r) { HFr x = HFr::from_le_bytes_mod_order(b, n); memcpy(r, x.v, 32); }
void emu_g1_compress(const uint64_t* xy, uint8_t* out) { g1_compress(xy, out); }
// appends `msg`, draws two challenges, appends msg2, draws a third; out: 3 x 4 limbs (Montgomery)
void emu_transcript(const uint8_t* msg, size_t n, const uint8_t* msg2, size_t n2, uint64_t* out) {
  SolidityTranscript t;
  t.append_message(msg, n);
  HFr a = t.get_and_append_challenge(), b = t.get_and_append_challenge();
  t.append_message(msg2, n2);
  HFr c = t.get_and_append_challenge();
  memcpy(out, a.v, 32); memcpy(out + 4, b.v, 32); memcpy(out + 8, c.v, 32);
}
}
