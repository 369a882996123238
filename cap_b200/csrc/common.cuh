// Shared host-side plumbing of libcapgpu: context, error handling, device buffers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <mutex>

#include "../../include/capgpu.h"
#include "ec.cuh"

namespace capgpu {

struct CudaError {
  cudaError_t code;
  const char* file;
  int line;
};

#define CAPGPU_CUDA(expr)                                                \
  do {                                                                   \
    cudaError_t _e = (expr);                                             \
    if (_e != cudaSuccess) throw ::capgpu::CudaError{_e, __FILE__, __LINE__}; \
  } while (0)

struct ArgError { const char* what; };
#define CAPGPU_REQUIRE(cond, msg) do { if (!(cond)) throw ::capgpu::ArgError{msg}; } while (0)
struct CodeError { int code; };

// Simple growable device buffer owned by a ctx (stream-ordered frees are not needed: a ctx
// is single-threaded and buffers live as long as the ctx).
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void reserve(size_t b) {
    if (b <= bytes) return;
    if (p) CAPGPU_CUDA(cudaFree(p));
    p = nullptr; bytes = 0;
    CAPGPU_CUDA(cudaMalloc(&p, b));
    bytes = b;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct NttDomain;  // ntt.cu

}  // namespace capgpu

struct capgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  uint64_t launches = 0;
  int sm_count = 148;
  // NTT domains (twiddle tables), keyed by log_n; owned by the ctx
  std::map<unsigned, capgpu::NttDomain*> domains;
  // workspaces
  capgpu::DevBuf ntt_tmp, ntt_io;
  capgpu::DevBuf msm_scalars, msm_digits, msm_counts, msm_entries, msm_buckets, msm_partials, msm_out;
  capgpu::DevBuf msm_flat;  // chunk partials of the flat (low-latency) accumulation
  capgpu::DevBuf msm_ticket;  // per-vector arrival counters of msm_red_planes (zero between launches)
  cudaEvent_t sync_ev = nullptr;  // cudaEventBlockingSync: host threads sleep while a round runs
  void* pinned = nullptr;  // small pinned staging area
  size_t pinned_bytes = 0;
  // optional per-kernel timing (capgpu_profile_enable): CUDA events on the ctx stream around
  // the instrumented launches, accumulated per kernel class
  bool latency_mode = false;  // prover MSMs favour depth over total work (capgpu_ctx_set_latency_mode)
  int group = 8;              // notes proved in lockstep by capgpu_prove_batch / the queue (capgpu_ctx_set_group)
  bool profile = false;
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  double prof_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double prof_units[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t prof_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  struct capgpu_job* cached_job = nullptr;  // workspace reused across capgpu_prove calls
};

struct capgpu_srs {
  int device = 0;
  size_t n = 0;
  int c = 0, W = 0;
  size_t K = 0;  // 2^(c-1) buckets
  capgpu::G1Affine* table = nullptr;  // W x n window-shifted bases: table[w*n + i] = 2^(c*w) * P_i
};

namespace capgpu {

inline void set_device(const capgpu_ctx* ctx) { CAPGPU_CUDA(cudaSetDevice(ctx->device)); }

// Wraps an ABI entry point: maps exceptions to status codes, records CUDA error text.
template <class F>
int guarded(capgpu_ctx* ctx, F&& f) {
  try {
    if (ctx) set_device(ctx);
    f();
    return CAPGPU_OK;
  } catch (const CudaError& e) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s (%s) at %s:%d", cudaGetErrorName(e.code), cudaGetErrorString(e.code), e.file, e.line);
    if (ctx) ctx->last_error = buf;
    return CAPGPU_ERR_CUDA;
  } catch (const ArgError& e) {
    if (ctx) ctx->last_error = e.what;
    return CAPGPU_ERR_ARG;
  } catch (const CodeError& e) {
    return e.code;
  } catch (const std::bad_alloc&) {
    if (ctx) ctx->last_error = "host allocation failed";
    return CAPGPU_ERR_ARG;
  }
}

#define CAPGPU_LAUNCH_CHECK(ctx) do { (ctx)->launches++; CAPGPU_CUDA(cudaGetLastError()); } while (0)

// Waits for everything queued on the ctx stream.  In the default (throughput) mode the host
// thread blocks on an event created with cudaEventBlockingSync instead of spinning, so several
// prover contexts per GPU (and several GPUs per host) do not burn a core each between rounds;
// latency mode keeps the lower-latency spinning synchronise.
inline void ctx_wait(capgpu_ctx* ctx) {
  if (ctx->latency_mode || !ctx->sync_ev) {
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  CAPGPU_CUDA(cudaEventRecord(ctx->sync_ev, ctx->stream));
  CAPGPU_CUDA(cudaEventSynchronize(ctx->sync_ev));
}

inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Kernel classes timed by the optional profiler (ids of capgpu_profile_read).
enum ProfId { PROF_MSM_ACCUMULATE = 0, PROF_NTT = 1, PROF_QUOTIENT = 2, PROF_MSM_SORT = 3, PROF_MSM_REDUCE = 4, PROF_GRAND_PRODUCT = 5 };

// Brackets the launches issued in its scope with CUDA events on the ctx stream when profiling
// is on (the scope then ends with an event synchronise, so profiled runs are serialised).
struct ProfScope {
  capgpu_ctx* ctx;
  int id;
  double units;
  ProfScope(capgpu_ctx* c, int i, double u) : ctx(c), id(i), units(u) {
    if (ctx->profile) cudaEventRecord(ctx->pe0, ctx->stream);
  }
  ~ProfScope() {
    if (!ctx->profile) return;
    cudaEventRecord(ctx->pe1, ctx->stream);
    if (cudaEventSynchronize(ctx->pe1) != cudaSuccess) return;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->pe0, ctx->pe1) != cudaSuccess) return;
    ctx->prof_ms[id] += ms;
    ctx->prof_units[id] += units;
    ctx->prof_cnt[id]++;
  }
};

// ---- ntt.cu -----------------------------------------------------------------------------
NttDomain* get_domain(capgpu_ctx* ctx, unsigned log_n);
void destroy_domain(NttDomain* d);
// dst (batch x n, stride dst_stride) = NTT of src (batch vectors of src_len valid elements,
// stride src_stride); tmp: scratch of batch x n elements (stride n), may equal dst only for
// single-pass sizes.  All pointers are device pointers.
void ntt_device(capgpu_ctx* ctx, unsigned log_n, const Fr* src, size_t src_len, size_t src_stride, Fr* dst,
                size_t dst_stride, Fr* tmp, size_t batch, bool inverse, bool coset);
const Fr* domain_omega_powers(capgpu_ctx* ctx, unsigned log_n);  // omega^j, j < n (device)
// The 3 * 2^log_n-point domain g <rho> as three cosets (ntt.cu): forward writes rows (b, k) of N = 2^log_n values, dst stride 3N per
// input; inverse turns such rows into the 3N coefficients in place.
void ntt3_forward(capgpu_ctx* ctx, unsigned log_n, const Fr* src, size_t src_len, size_t src_stride, Fr* dst, Fr* tmp, size_t batch);
void ntt3_inverse(capgpu_ctx* ctx, unsigned log_n, Fr* t, Fr* tmp, size_t batch);

// ---- msm.cu -----------------------------------------------------------------------------
// scalars: device, batch vectors of n Fr (stride `stride`); out: device, batch affine points.
capgpu_srs* srs_lagrange(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, const Fr* omega_pows, const Fr& n_inv);
void msm_device(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const Fr* scalars, size_t n, size_t stride,
                size_t batch, bool scalars_mont, G1Affine* out_dev, bool latency = false, size_t part = 0, size_t parts = 1,
                G1XYZZ* out_xyzz = nullptr,          // write the XYZZ sums instead of affine points (slices of a split MSM)
                const struct PeerOut* peer = nullptr);  // ... or deliver the sum into peer-mapped memory (msm_reduce.cuh)

}  // namespace capgpu
