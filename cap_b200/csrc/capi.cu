// extern "C" surface of libcapgpu: context management, standalone NTT entry point,
// calibration kernels.  (MSM entry points live in msm.cu, prover entry points in prover.cu.)
#include "common.cuh"
#include <stdlib.h>

using namespace capgpu;

extern "C" const char* capgpu_strerror(int code) {
  switch (code) {
    case CAPGPU_OK: return "ok";
    case CAPGPU_ERR_CUDA: return "CUDA runtime error";
    case CAPGPU_ERR_ARG: return "invalid argument";
    case CAPGPU_ERR_DEGREE: return "quotient polynomial has wrong degree";
    case CAPGPU_ERR_SRS_TOO_SMALL: return "commit key too small for polynomial";
    case CAPGPU_ERR_STATE: return "round API called out of order";
    default: return "unknown error";
  }
}

extern "C" const char* capgpu_last_error(const capgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

void capgpu_job_free_internal(capgpu_job* job);  // prover.cu

extern "C" int capgpu_ctx_create(int device, capgpu_ctx** out) {
  if (!out) return CAPGPU_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return CAPGPU_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= count) return CAPGPU_ERR_ARG;
  capgpu_ctx* ctx = new capgpu_ctx();
  ctx->device = device;
  int rc = guarded(ctx, [&] {
    CAPGPU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    cudaDeviceProp prop;
    CAPGPU_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("CAPGPU_GROUP")) { int g = atoi(e); if (g >= 1 && g <= 64) ctx->group = g; }
    ctx->pinned_bytes = 1 << 16;
    CAPGPU_CUDA(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    CAPGPU_CUDA(cudaEventCreateWithFlags(&ctx->sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
  });
  if (rc != CAPGPU_OK) { delete ctx; return rc; }
  *out = ctx;
  return CAPGPU_OK;
}

extern "C" void capgpu_ctx_destroy(capgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->cached_job) capgpu_job_free_internal(ctx->cached_job);
  for (auto& kv : ctx->domains) destroy_domain(kv.second);
  ctx->ntt_tmp.release(); ctx->ntt_io.release();
  ctx->msm_scalars.release(); ctx->msm_digits.release(); ctx->msm_counts.release();
  ctx->msm_entries.release(); ctx->msm_buckets.release(); ctx->msm_partials.release(); ctx->msm_out.release();
  ctx->msm_flat.release();
  ctx->msm_ticket.release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->sync_ev) cudaEventDestroy(ctx->sync_ev);
  if (ctx->pe0) cudaEventDestroy(ctx->pe0);
  if (ctx->pe1) cudaEventDestroy(ctx->pe1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" int capgpu_ctx_sync(capgpu_ctx* ctx) {
  if (!ctx) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] { CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

extern "C" int capgpu_ctx_set_latency_mode(capgpu_ctx* ctx, int on) {
  if (!ctx) return CAPGPU_ERR_ARG;
  ctx->latency_mode = on != 0;
  return CAPGPU_OK;
}

extern "C" void* capgpu_ctx_stream(capgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" uint64_t capgpu_launch_count(const capgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int capgpu_ntt(capgpu_ctx* ctx, const uint64_t* in, size_t in_len, uint64_t* out, unsigned log_n, size_t batch,
                          int inverse, int coset) {
  if (!ctx || !in || !out) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(log_n >= 1 && log_n <= 20, "NTT size must be 2^1 .. 2^20");
    const size_t n = (size_t)1 << log_n;
    CAPGPU_REQUIRE(in_len >= 1 && in_len <= n, "NTT input length out of range");
    ctx->ntt_io.reserve(batch * n * sizeof(Fr));
    ctx->ntt_tmp.reserve(batch * n * sizeof(Fr));
    Fr* io = ctx->ntt_io.as<Fr>();
    if (in_len == n) {
      CAPGPU_CUDA(cudaMemcpyAsync(io, in, batch * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    } else {
      CAPGPU_CUDA(cudaMemcpy2DAsync(io, n * sizeof(Fr), in, in_len * sizeof(Fr), in_len * sizeof(Fr), batch,
                                    cudaMemcpyHostToDevice, ctx->stream));
    }
    ntt_device(ctx, log_n, io, in_len, n, io, n, ctx->ntt_tmp.as<Fr>(), batch, inverse != 0, coset != 0);
    CAPGPU_CUDA(cudaMemcpyAsync(out, io, batch * n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

extern "C" int capgpu_ntt_dev(capgpu_ctx* ctx, const void* d_in, size_t in_len, void* d_out, unsigned log_n, size_t batch,
                              int inverse, int coset) {
  if (!ctx || !d_in || !d_out) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(log_n >= 1 && log_n <= 20, "NTT size must be 2^1 .. 2^20");
    const size_t n = (size_t)1 << log_n;
    CAPGPU_REQUIRE(in_len >= 1 && in_len <= n, "NTT input length out of range");
    CAPGPU_REQUIRE(d_in != d_out || in_len == n, "in-place NTT needs a full-length input");
    ctx->ntt_tmp.reserve(batch * n * sizeof(Fr));
    ntt_device(ctx, log_n, (const Fr*)d_in, in_len, in_len, (Fr*)d_out, n, ctx->ntt_tmp.as<Fr>(), batch, inverse != 0, coset != 0);
  });
}

extern "C" int capgpu_ntt3_dev(capgpu_ctx* ctx, const void* d_in, size_t in_len, void* d_out, unsigned log_n, size_t batch, int inverse) {
  if (!ctx || !d_in || !d_out) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(log_n >= 1 && log_n <= 20, "NTT size must be 2^1 .. 2^20");
    const size_t n = (size_t)1 << log_n;
    ctx->ntt_tmp.reserve(3 * batch * n * sizeof(Fr));
    if (inverse) {
      CAPGPU_REQUIRE(in_len == 3 * n, "the inverse takes all 3 * 2^log_n values");
      if (d_in != d_out) CAPGPU_CUDA(cudaMemcpyAsync(d_out, d_in, batch * 3 * n * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
      ntt3_inverse(ctx, log_n, (Fr*)d_out, ctx->ntt_tmp.as<Fr>(), batch);
    } else {
      CAPGPU_REQUIRE(in_len >= 1 && in_len <= n, "NTT input length out of range");
      CAPGPU_REQUIRE(d_in != d_out, "the forward 3-coset transform is out of place");
      ntt3_forward(ctx, log_n, (const Fr*)d_in, in_len, in_len, (Fr*)d_out, ctx->ntt_tmp.as<Fr>(), batch);
    }
  });
}

// ------------------------------------------------------------------------------------------
// calibration: integer multiply-add issue rate (the roofline denominator for MSM / NTT)
// ------------------------------------------------------------------------------------------
namespace capgpu {

__global__ void calib_imad(uint32_t* out, uint32_t m, uint32_t c, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a4) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a5) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a6) : "r"(m), "r"(c));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a7) : "r"(m), "r"(c));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

__global__ void calib_imad_wide(uint64_t* out, uint32_t m, int iters) {
  uint64_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const uint32_t b0 = m ^ threadIdx.x, b1 = b0 + 11, b2 = b0 + 22, b3 = b0 + 33;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"(b0), "r"(m));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"(b1), "r"(m));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"(b2), "r"(m));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"(b3), "r"(m));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"(b0), "r"(b1));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"(b1), "r"(b2));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"(b2), "r"(b3));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"(b3), "r"(b0));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

// Montgomery products with REGISTER operands on every lane (the second operand comes from memory,
// both chains depend on threadIdx): an earlier version multiplied by a compile-time constant with one
// warp-uniform chain, which the compiler moved to the uniform datapath and folded into immediates --
// it reported 109 G products/s where real kernels can reach 68.
__global__ void calib_fmul(Fq* out, const Fq* yin, int iters) {
  Fq y = yin[threadIdx.x & 1];
  Fq x = Fq::one(), x2 = Fq::r2();
  x.v[0] += threadIdx.x;
  x2.v[1] += threadIdx.x + blockIdx.x;
  for (int i = 0; i < iters; i++) {
    x = fp_mul(x, y);
    x2 = fp_mul(x2, y);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fp_add(x, x2);
}

}  // namespace capgpu

extern "C" int capgpu_calibrate(capgpu_ctx* ctx, double* gimad_per_s, double* gimad_wide_per_s, double* gfmul_per_s) {
  if (!ctx) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    const int blocks = ctx->sm_count * 8, threads = 256;
    void* buf = nullptr;
    CAPGPU_CUDA(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(Fq) + 2 * sizeof(Fq)));
    cudaEvent_t e0, e1;
    CAPGPU_CUDA(cudaEventCreate(&e0));
    CAPGPU_CUDA(cudaEventCreate(&e1));
    auto time_it = [&](auto launch) {
      float best = 1e30f;
      for (int rep = 0; rep < 4; rep++) {
        CAPGPU_CUDA(cudaEventRecord(e0, ctx->stream));
        launch();
        CAPGPU_LAUNCH_CHECK(ctx);
        CAPGPU_CUDA(cudaEventRecord(e1, ctx->stream));
        CAPGPU_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CAPGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      return (double)best * 1e-3;
    };
    const int iters = 2048;
    double t = time_it([&] { calib_imad<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)buf, 0x9e3779b1u, 12345u, iters); });
    if (gimad_per_s) *gimad_per_s = (double)blocks * threads * iters * 64.0 / t * 1e-9;
    t = time_it([&] { calib_imad_wide<<<blocks, threads, 0, ctx->stream>>>((uint64_t*)buf, 0x9e3779b1u, iters); });
    if (gimad_wide_per_s) *gimad_wide_per_s = (double)blocks * threads * iters * 64.0 / t * 1e-9;
    const int fiters = 512;
    Fq hy[2] = {Fq::r2(), Fq::one()};
    hy[1].v[2] ^= 0x5a5a5a5u;
    Fq* yin = reinterpret_cast<Fq*>(static_cast<char*>(buf) + (size_t)blocks * threads * sizeof(Fq));
    CAPGPU_CUDA(cudaMemcpyAsync(yin, hy, sizeof hy, cudaMemcpyHostToDevice, ctx->stream));
    t = time_it([&] { calib_fmul<<<blocks, threads, 0, ctx->stream>>>((Fq*)buf, yin, fiters); });
    if (gfmul_per_s) *gfmul_per_s = (double)blocks * threads * fiters * 2.0 / t * 1e-9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
  });
}

extern "C" int capgpu_profile_enable(capgpu_ctx* ctx, int on) {
  if (!ctx) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    if (on && !ctx->pe0) {
      CAPGPU_CUDA(cudaEventCreate(&ctx->pe0));
      CAPGPU_CUDA(cudaEventCreate(&ctx->pe1));
    }
    for (int i = 0; i < 8; i++) { ctx->prof_ms[i] = 0; ctx->prof_units[i] = 0; ctx->prof_cnt[i] = 0; }
    ctx->profile = on != 0;
  });
}

extern "C" int capgpu_profile_read(const capgpu_ctx* ctx, int id, double* total_ms, uint64_t* launches, double* units) {
  if (!ctx || id < 0 || id >= 8) return CAPGPU_ERR_ARG;
  if (total_ms) *total_ms = ctx->prof_ms[id];
  if (launches) *launches = ctx->prof_cnt[id];
  if (units) *units = ctx->prof_units[id];
  return CAPGPU_OK;
}
