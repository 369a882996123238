// C++ host-side mirror of the jf-plonk interface the CAP prover calls, over the C ABI of capgpu.h.
//
// The reference's host is Rust (`jf_plonk::proof_system::{PlonkKzgSnark, UniversalSNARK}`, imported at
// /root/reference/src/proof/transfer.rs:40-43, called at :133 (preprocess) and :181 (prove)); the Rust binding is
// rust/jf-plonk-gpu (not buildable in the image this repository is developed in).  This header is the same thin layer
// for a C++ host: RAII handles, `PlonkError`-style exceptions, and `PlonkKzgSnark::prove` / `ProverQueue` with the
// argument meaning of the reference (witness columns, public inputs, the RNG's field draws, extra_transcript_init_msg).
// Header-only, C++17, no dependency beyond libcapgpu.so.  Exercised by tests/cpp/replay_fixture.cpp.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "capgpu.h"

namespace capgpu_host {

// jf_plonk::errors::PlonkError as CAP sees it (mapped to TxnApiError::FailedSnark at src/proof/transfer.rs:187)
struct PlonkError : std::runtime_error {
  int code;
  PlonkError(int c, const std::string& detail)
      : std::runtime_error(std::string(c == CAPGPU_ERR_DEGREE ? "WrongQuotientPolyDegree" : c == CAPGPU_ERR_SRS_TOO_SMALL ? "IndexTooLarge / SRS too small" : capgpu_strerror(c)) +
                           (detail.empty() ? "" : " [" + detail + "]")),
        code(c) {}
};

class Context {
 public:
  explicit Context(int device = 0) { int rc = capgpu_ctx_create(device, &h_); if (rc) throw PlonkError(rc, "capgpu_ctx_create"); }
  ~Context() { if (h_) capgpu_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  capgpu_ctx* get() const { return h_; }
  void check(int rc) const { if (rc) throw PlonkError(rc, capgpu_last_error(h_)); }
  void set_latency_mode(bool on) { check(capgpu_ctx_set_latency_mode(h_, on ? 1 : 0)); }
  void set_group(int g) { check(capgpu_ctx_set_group(h_, g)); }
 private:
  capgpu_ctx* h_ = nullptr;
};

// UniversalSrs (src/proof/mod.rs:59-109)
class UniversalSrs {
 public:
  // PlonkKzgSnark::universal_setup with a known tau (tests / benchmarks: universal_setup_for_staging)
  static UniversalSrs universal_setup(Context& ctx, size_t max_degree, const uint64_t tau_mont[4]) {
    UniversalSrs s; ctx.check(capgpu_srs_setup(ctx.get(), tau_mont, max_degree + 1, 0, &s.h_)); return s;
  }
  // `UniversalSrs::deserialize` of a CanonicalSerialize blob, optional sha256 gate (src/proof/mod.rs:98-107)
  static UniversalSrs load(Context& ctx, const std::vector<uint8_t>& bytes, const uint8_t* expect_sha256 = nullptr, size_t max_points = 0) {
    UniversalSrs s; ctx.check(capgpu_srs_load_serialized(ctx.get(), bytes.data(), bytes.size(), expect_sha256, max_points, 0, &s.h_)); return s;
  }
  UniversalSrs(UniversalSrs&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~UniversalSrs() { if (h_) capgpu_srs_destroy(h_); }
  capgpu_srs* get() const { return h_; }
  size_t size() const { return capgpu_srs_size(h_); }
 private:
  UniversalSrs() = default;
  capgpu_srs* h_ = nullptr;
};

// jf_plonk ProvingKey (embedded in TransferProvingKey, src/proof/transfer.rs:60)
class ProvingKey {
 public:
  // `ProvingKey::deserialize` of the key's own CanonicalSerialize bytes
  static ProvingKey deserialize(Context& ctx, const std::vector<uint8_t>& bytes) {
    ProvingKey k; ctx.check(capgpu_pk_load_serialized(ctx.get(), bytes.data(), bytes.size(), nullptr, &k.h_)); k.read_info(); return k;
  }
  ProvingKey(ProvingKey&& o) noexcept : h_(std::exchange(o.h_, nullptr)), log_n_(o.log_n_), num_inputs_(o.num_inputs_) {}
  ~ProvingKey() { if (h_) capgpu_pk_destroy(h_); }
  capgpu_pk* get() const { return h_; }
  size_t domain_size() const { return size_t(1) << log_n_; }
  size_t num_inputs() const { return num_inputs_; }
 private:
  ProvingKey() = default;
  void read_info() { uint64_t k[20]; capgpu_pk_info(h_, &log_n_, &num_inputs_, k); }
  capgpu_pk* h_ = nullptr;
  unsigned log_n_ = 0;
  size_t num_inputs_ = 0;
};

// Device-resident proving keys keyed by (note type, inputs, outputs, tree depth): the GPU-side analogue of the
// reference's on-disk key cache `transfer_prover_{i}_input_{o}_output_{d}_depth.bin` (src/parameters.rs:485-503).
// A key is uploaded (deserialised, coset tables and Lagrange commit key derived on the device) once per shape.
class ProvingKeyCache {
 public:
  using Shape = std::tuple<uint64_t, uint64_t, uint64_t, uint64_t>;  // note type (0 transfer, 1 mint, 2 freeze), n_inputs, n_outputs, depth
  explicit ProvingKeyCache(Context& ctx) : ctx_(ctx) {}
  // `load` supplies the key's CanonicalSerialize bytes (from disk, or `ProvingKey::serialize` of a freshly preprocessed key)
  const ProvingKey& get(const Shape& shape, const std::function<std::vector<uint8_t>()>& load) {
    std::lock_guard<std::mutex> lock(mu_);
    auto it = keys_.find(shape);
    if (it == keys_.end()) it = keys_.emplace(shape, std::make_unique<ProvingKey>(ProvingKey::deserialize(ctx_, load()))).first;
    return *it->second;
  }
  size_t size() const { return keys_.size(); }
 private:
  Context& ctx_;
  std::mutex mu_;
  std::map<Shape, std::unique_ptr<ProvingKey>> keys_;
};

struct Proof {
  capgpu_proof raw;
  // `Proof::serialize` (CanonicalSerialize, compressed points)
  std::vector<uint8_t> serialize() const {
    std::vector<uint8_t> out(1024);
    size_t len = 0;
    int rc = capgpu_proof_serialize(&raw, out.data(), out.size(), &len);
    if (rc) throw PlonkError(rc, "capgpu_proof_serialize");
    out.resize(len);
    return out;
  }
};

// The field elements the prover draws from its RNG, re-drawn from recorded `next_u64` words the way `Fr::rand` does
inline std::vector<uint64_t> blinders_from_rng_words(const std::vector<uint64_t>& words) {
  for (size_t count : {size_t(17), size_t(13)}) {
    std::vector<uint64_t> bl(17 * 4, 0);
    size_t used = 0;
    if (capgpu_fr_rand_from_words(words.data(), words.size(), bl.data(), count, &used) == 0 && used == words.size()) return bl;
  }
  throw PlonkError(CAPGPU_ERR_ARG, "recorded RNG words match neither 17 nor 13 Fr::rand draws");
}

struct PlonkKzgSnark {
  // PlonkKzgSnark::prove::<_, _, SolidityTranscript>(rng, &circuit, &pk, Some(ext_msg)) (src/proof/transfer.rs:181):
  // wires = 5 x n witness values per wire column, pub_inputs, blinders (17 x 4 limbs), all Montgomery
  static Proof prove(Context& ctx, const ProvingKey& pk, const std::vector<uint64_t>& wires, const std::vector<uint64_t>& pub_inputs,
                     const std::vector<uint64_t>& blinders, const std::vector<uint8_t>& ext_msg) {
    if (wires.size() != 5 * pk.domain_size() * 4 || pub_inputs.size() != pk.num_inputs() * 4 || blinders.size() != 17 * 4)
      throw PlonkError(CAPGPU_ERR_ARG, "argument sizes do not match the proving key");
    Proof p;
    ctx.check(capgpu_prove(ctx.get(), pk.get(), wires.data(), pub_inputs.data(), blinders.data(), ext_msg.empty() ? nullptr : ext_msg.data(),
                           ext_msg.size(), &p.raw));
    return p;
  }
};

// Asynchronous proving (the overlap point src/proof/transfer.rs:167-181): submit returns at once, wait delivers
class ProverQueue {
 public:
  ProverQueue(const std::vector<Context*>& ctxs, const ProvingKey& pk, size_t ring_slots = 0) {
    std::vector<capgpu_ctx*> raw;
    for (Context* c : ctxs) raw.push_back(c->get());
    int rc = capgpu_queue_create(raw.data(), raw.size(), pk.get(), ring_slots, &q_);
    if (rc) throw PlonkError(rc, "capgpu_queue_create");
  }
  ~ProverQueue() { if (q_) capgpu_queue_destroy(q_); }
  ProverQueue(const ProverQueue&) = delete;
  uint64_t submit(const std::vector<uint64_t>& wires, const std::vector<uint64_t>& pub_inputs, const std::vector<uint64_t>& blinders,
                  const std::vector<uint8_t>& ext_msg) {
    uint64_t t = 0;
    int rc = capgpu_submit(q_, wires.data(), pub_inputs.data(), blinders.data(), ext_msg.empty() ? nullptr : ext_msg.data(), ext_msg.size(), &t);
    if (rc) throw PlonkError(rc, "capgpu_submit");
    return t;
  }
  bool poll(uint64_t ticket) { int done = 0; int rc = capgpu_poll(q_, ticket, &done); if (rc) throw PlonkError(rc, "capgpu_poll"); return done != 0; }
  Proof wait(uint64_t ticket) { Proof p; int rc = capgpu_wait(q_, ticket, &p.raw); if (rc) throw PlonkError(rc, "capgpu_wait"); return p; }
 private:
  capgpu_queue* q_ = nullptr;
};

}  // namespace capgpu_host
