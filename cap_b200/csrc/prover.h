// Internal declarations shared by prover.cu (rounds, groups, queue) and formats.cu (key loaders).
#pragma once
#include <functional>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "hostfp.h"

#define CAPGPU_MAX_GROUP 64

struct capgpu_pk {
  int device = 0;
  unsigned log_n = 0;
  size_t n = 0, m = 0, num_inputs = 0;
  // quotient domain (QuotDomain, poly.cuh): one coset of m = 8n points, or (q3) the three cosets g rho^k H_2n of the 6n-point domain
  bool q3 = false;
  size_t qsub = 0;
  unsigned qlog_sub = 0, qstep = 0;
  const capgpu_srs* srs = nullptr;
  capgpu_srs* owned_srs = nullptr;  // commit key embedded in a serialized ProvingKey (freed with the key)
  capgpu_srs* lag = nullptr;        // Lagrange-basis commit key [L_0..L_{n-1}, P_0, P_1, P_n, P_{n+1}] (owned)
  bool use_lag = true;
  capgpu::Fr *sel_coef = nullptr, *sig_coef = nullptr, *sig_eval = nullptr, *sel_coset = nullptr, *sig_coset = nullptr;
  capgpu::Fr *xs = nullptr, *l1inv = nullptr, *zh_inv = nullptr, *omega_n = nullptr;
  capgpu::HFr k[5];
  uint64_t sel_comms[13][8], sig_comms[5][8];
  std::vector<uint8_t> vk_bytes;
};

namespace capgpu {

// One note handed to the prover: 5 x n wire values (host or device memory), public inputs,
// the 17 blinders in draw order, the extra transcript message.
struct NoteIn {
  const uint64_t* wires;
  bool wires_on_device;
  const uint64_t* pub_inputs;
  const uint64_t* blinders;
  const uint8_t* ext_msg;
  size_t ext_msg_len;
};

void prove_group(capgpu_ctx* ctx, const capgpu_pk* pk, int G, const NoteIn* notes, capgpu_proof* const* out, int* status,
                 const std::function<void()>& inputs_consumed, int cap_hint);

int pk_create_from_coefficients(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, size_t num_inputs, const uint64_t* selectors,
                                const uint64_t* sigmas, const uint64_t* k, const uint64_t* selector_comms_xy,
                                const uint64_t* sigma_comms_xy, capgpu_pk** out);

}  // namespace capgpu
