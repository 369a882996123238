// Replay of a CAPFIX01 fixture from a COMPILED host through include/capgpu.hpp: the key,
// witness, public inputs, transcript message and RNG words recorded from a `PlonkKzgSnark::prove` call
// (/root/reference/src/proof/transfer.rs:159-188) are proved again on the GPU - once through the blocking call, once
// through the asynchronous queue - and the serialized `Proof` must equal the recorded bytes.  Reads like the
// reference's own test_transfer_validity_proof (src/proof/transfer.rs:600): prove, then check.
//   g++ -std=c++17 -I include tests/cpp/replay_fixture.cpp -L cap_b200 -lcapgpu -o replay_fixture
//   LD_LIBRARY_PATH=cap_b200 ./replay_fixture tests/fixtures/oracle_transfer_n64.capfix
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "capgpu.hpp"

using namespace capgpu_host;

static const uint64_t R_MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};

// canonical -> Montgomery (x * 2^256 mod r) by 256 modular doublings; the fixture stores `Fr::serialize` bytes
static void to_mont(const uint8_t* le32, uint64_t out[4]) {
  uint64_t x[4];
  memcpy(x, le32, 32);
  for (int i = 0; i < 256; i++) {
    uint64_t c = 0, y[4];
    for (int l = 0; l < 4; l++) { y[l] = (x[l] << 1) | c; c = x[l] >> 63; }
    unsigned __int128 br = 0;
    uint64_t z[4];
    for (int l = 0; l < 4; l++) { unsigned __int128 d = (unsigned __int128)y[l] - R_MOD[l] - (uint64_t)br; z[l] = (uint64_t)d; br = (d >> 64) & 1; }
    const bool ge = c || !br;
    for (int l = 0; l < 4; l++) x[l] = ge ? z[l] : y[l];
  }
  memcpy(out, x, 32);
}

static uint64_t rd64(const std::string& s, size_t off) { uint64_t v; memcpy(&v, s.data() + off, 8); return v; }

static std::vector<uint64_t> vec_fr_mont(const std::string& s, size_t& off) {
  const uint64_t cnt = rd64(s, off);
  off += 8;
  std::vector<uint64_t> out(cnt * 4);
  for (uint64_t i = 0; i < cnt; i++) to_mont(reinterpret_cast<const uint8_t*>(s.data()) + off + 32 * i, &out[4 * i]);
  off += 32 * cnt;
  return out;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s FIXTURE.capfix\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (data.size() < 16 || data.compare(0, 8, "CAPFIX01") != 0) { fprintf(stderr, "not a CAPFIX01 file\n"); return 2; }
  std::map<std::string, std::string> sec;
  size_t off = 16;
  for (uint64_t i = 0, n = rd64(data, 8); i < n; i++) {
    std::string tag(data.c_str() + off);  // zero-padded to 8 bytes
    tag = tag.substr(0, 8);
    const uint64_t len = rd64(data, off + 8);
    sec[tag] = data.substr(off + 16, len);
    off += 16 + len;
  }
  try {
    Context ctx(0);
    // keyed like the reference's key files: (note type, inputs, outputs, depth) from the META section
    uint64_t meta[4];
    memcpy(meta, sec["META"].data(), 32);
    ProvingKeyCache cache(ctx);
    int loads = 0;
    auto loader = [&] { loads++; return std::vector<uint8_t>(sec["PK"].begin(), sec["PK"].end()); };
    const ProvingKeyCache::Shape shape{meta[0], meta[1], meta[2], meta[3]};
    cache.get(shape, loader);
    const ProvingKey& pk = cache.get(shape, loader);  // second request: served from the device-resident cache
    if (loads != 1 || cache.size() != 1) throw PlonkError(CAPGPU_ERR_STATE, "proving-key cache reloaded a resident key");
    // WIRES: Vec<Vec<Fr>> -> 5 x n Montgomery limbs
    const std::string& w = sec["WIRES"];
    if (rd64(w, 0) != 5) throw PlonkError(CAPGPU_ERR_ARG, "fixture must hold 5 witness columns");
    size_t o = 8;
    std::vector<uint64_t> wires;
    for (int c = 0; c < 5; c++) { auto col = vec_fr_mont(w, o); wires.insert(wires.end(), col.begin(), col.end()); }
    size_t po = 0;
    std::vector<uint64_t> pub = vec_fr_mont(sec["PUBIN"], po);
    const std::string& e = sec["EXTMSG"];
    std::vector<uint8_t> ext(e.begin() + 8, e.begin() + 8 + rd64(e, 0));
    const std::string& g = sec["RNGU64"];
    std::vector<uint64_t> words(rd64(g, 0));
    memcpy(words.data(), g.data() + 8, words.size() * 8);
    std::vector<uint64_t> blinders = blinders_from_rng_words(words);
    const std::vector<uint8_t> want(sec["PROOF"].begin(), sec["PROOF"].end());

    Proof p1 = PlonkKzgSnark::prove(ctx, pk, wires, pub, blinders, ext);
    const bool ok1 = p1.serialize() == want;
    // the same note three times through the asynchronous queue (lockstep group), buffers dropped after submit
    Context ctx2(0);
    ProverQueue q({&ctx2}, pk);
    uint64_t t[3];
    for (int i = 0; i < 3; i++) { std::vector<uint64_t> copy = wires; t[i] = q.submit(copy, pub, blinders, ext); }
    bool ok2 = true;
    for (int i = 0; i < 3; i++) ok2 = ok2 && q.wait(t[i]).serialize() == want;
    // an unsatisfied witness must come back as WrongQuotientPolyDegree, like the reference's FailedSnark
    std::vector<uint64_t> bad = wires;
    bad[4 * 7] ^= 1;
    bool ok3 = false;
    try { PlonkKzgSnark::prove(ctx, pk, bad, pub, blinders, ext); } catch (const PlonkError& err) { ok3 = err.code == CAPGPU_ERR_DEGREE; }
    Proof p4 = PlonkKzgSnark::prove(ctx, pk, wires, pub, blinders, ext);  // the context stays usable after the failure
    const bool ok4 = p4.serialize() == want;
    printf("domain %zu inputs %zu proof_bytes %zu blocking=%d queue=%d bad_witness_rejected=%d after_failure=%d\n", pk.domain_size(), pk.num_inputs(),
           want.size(), ok1, ok2, ok3, ok4);
    if (ok1 && ok2 && ok3 && ok4) { printf("REPLAY_OK\n"); return 0; }
    return 1;
  } catch (const PlonkError& err) {
    fprintf(stderr, "PlonkError %d: %s\n", err.code, err.what());
    return 1;
  }
}
