// G1 multi-scalar multiplication and base-field arithmetic for the other pairing curves the reference can
// be configured with: BLS12-381 and BLS12-377 (/root/reference/src/config.rs:86-114: `PairingCurve`,
// `BaseField`; cargo features `bls12_381` / `bls12_377`, Cargo.toml:71-75).  Replaces ark-ec 0.3.0
// `VariableBaseMSM::multi_scalar_mul` for `GroupAffine<ark_bls12_381::g1::Parameters>` /
// `<ark_bls12_377::g1::Parameters>`; the result (an affine point) is unique, so it is bit-identical to
// arkworks' whatever the schedule.
//
// First correct path for the 12-limb fields (fpn.cuh), not the tuned BN254 pipeline of msm.cu: classic
// Pippenger with per-window buckets over the caller's bases (no precomputed tables, bases supplied per call):
//   digits     signed c-bit digits, one histogram atomic per digit that also returns the entry's rank
//   scan       exclusive scan over the W * 2^(c-1) bucket counters
//   scatter    entries grouped by (window, bucket), atomic-free
//   accumulate one thread per bucket: XYZZ mixed additions (8M + 2S)
//   reduce     one CTA per window: running sums over segments, small scalar multiples, block tree
//   final      Horner over the windows (c doublings each) and conversion to affine
// The BN254 prover (NTT domains, transcript, proving key) is not instantiated for these curves.
#include "common.cuh"
#include "fpn.cuh"
#include <vector>

namespace capgpu {

__global__ void cv_digits(const uint32_t* __restrict__ scalars, size_t n, int c, int W, int32_t* __restrict__ digits,
                          uint32_t* __restrict__ ranks, uint32_t* counts, size_t K) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[9];
  for (int l = 0; l < 8; l++) s[l] = scalars[i * 8 + l];
  s[8] = 0;
  const uint32_t mask = (1u << c) - 1u;
  const int32_t half = 1 << (c - 1);
  int32_t carry = 0;
  for (int w = 0; w < W; w++) {
    const uint32_t off = (uint32_t)(w * c), limb = off >> 5, sh = off & 31;
    uint32_t v = 0;
    if (limb < 8) {
      v = s[limb] >> sh;
      if (sh + c > 32) v |= s[limb + 1] << (32 - sh);
      v &= mask;
    }
    int32_t d = (int32_t)v + carry;
    carry = 0;
    if (d > half) { d -= (1 << c); carry = 1; }
    const size_t slot = (size_t)w * n + i;
    digits[slot] = d;
    if (d != 0) ranks[slot] = atomicAdd(&counts[(size_t)w * K + (uint32_t)(d < 0 ? -d : d) - 1], 1u);
  }
}

// exclusive scan of `total` counters in place (one CTA; counts[total] receives the grand total)
__global__ void __launch_bounds__(1024) cv_scan(uint32_t* counts, size_t total) {
  __shared__ uint32_t warp_sums[32];
  const size_t per = (total + blockDim.x - 1) / blockDim.x;
  const size_t i0 = (size_t)threadIdx.x * per;
  const size_t i1 = i0 + per < total ? i0 + per : total;
  uint32_t sum = 0;
  for (size_t i = i0; i < i1; i++) sum += counts[i];
  uint32_t x = sum;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t ws = warp_sums[threadIdx.x], z = ws;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (threadIdx.x >= o) z += y;
    }
    warp_sums[threadIdx.x] = z - ws;
  }
  __syncthreads();
  uint32_t run = warp_sums[threadIdx.x >> 5] + x - sum;
  for (size_t i = i0; i < i1; i++) { uint32_t v = counts[i]; counts[i] = run; run += v; }
  if (threadIdx.x == blockDim.x - 1) counts[total] = run;
}

__global__ void cv_scatter(const int32_t* __restrict__ digits, const uint32_t* __restrict__ ranks, size_t n, const uint32_t* __restrict__ offsets,
                           uint32_t* entries, size_t K) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t w = blockIdx.y, slot = w * n + i;
  const int32_t d = digits[slot];
  if (d == 0) return;
  const uint32_t k = (uint32_t)(d < 0 ? -d : d) - 1;
  entries[offsets[w * K + k] + ranks[slot]] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
}

template <class F>
__global__ void __launch_bounds__(128) cv_accumulate(const G1AffineT<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                                                     const uint32_t* __restrict__ offsets, G1XyzzT<F>* buckets, size_t nbuckets) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nbuckets) return;
  G1XyzzT<F> acc = G1XyzzT<F>::inf();
  for (uint32_t e = offsets[k]; e < offsets[k + 1]; e++) {
    const uint32_t u = entries[e];
    G1AffineT<F> p = bases[u & 0x7fffffffu];
    if (!p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, (u >> 31) != 0);
  }
  buckets[k] = acc;
}

template <class F>
__device__ inline G1XyzzT<F> cv_mul_small(const G1XyzzT<F>& p, uint32_t k) {
  G1XyzzT<F> r = G1XyzzT<F>::inf();
  int top = 31;
  while (top >= 0 && !((k >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    r = xyzz_dbl(r);
    if ((k >> i) & 1) xyzz_add(r, p);
  }
  return r;
}

// window w (blockIdx.x): sum_k (k + 1) * B[w][k] -> wsum[w]
template <class F>
__global__ void __launch_bounds__(256) cv_window_reduce(const G1XyzzT<F>* __restrict__ buckets, size_t K, uint32_t L, G1XyzzT<F>* wsum) {
  extern __shared__ unsigned char cv_smem_raw[];
  G1XyzzT<F>* sm = reinterpret_cast<G1XyzzT<F>*>(cv_smem_raw);
  const G1XyzzT<F>* B = buckets + (size_t)blockIdx.x * K;
  const uint32_t t = threadIdx.x;
  G1XyzzT<F> running = G1XyzzT<F>::inf(), acc = G1XyzzT<F>::inf();
  for (size_t k = (size_t)(t + 1) * L; k-- > (size_t)t * L;) {
    G1XyzzT<F> q = B[k];
    xyzz_add(running, q);
    xyzz_add(acc, running);
  }
  G1XyzzT<F> v = cv_mul_small(running, t * L);
  xyzz_add(v, acc);
  sm[t] = v;
  __syncthreads();
  for (uint32_t w = blockDim.x / 2; w >= 1; w >>= 1) {
    if (t < w) {
      G1XyzzT<F> x = sm[t], y = sm[t + w];
      xyzz_add(x, y);
      sm[t] = x;
    }
    __syncthreads();
  }
  if (t == 0) wsum[blockIdx.x] = sm[0];
}

template <class F>
__global__ void cv_final(const G1XyzzT<F>* __restrict__ wsum, int W, int c, G1AffineT<F>* out) {
  G1XyzzT<F> acc = wsum[W - 1];
  for (int w = W - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) acc = xyzz_dbl(acc);
    G1XyzzT<F> q = wsum[w];
    xyzz_add(acc, q);
  }
  *out = xyzz_to_affine(acc);
}

template <class F>
__global__ void cv_fq_op(int op, const F* a, const F* b, F* out, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  F x = a[i], y = b ? b[i] : F::zero(), r;
  switch (op) {
    case 0: r = fp_mul(x, y); break;
    case 1: r = fp_sqr(x); break;
    case 2: r = fp_inv(x); break;
    case 3: r = fp_add(x, y); break;
    case 4: r = fp_sub(x, y); break;
    default: r = fp_inv_fermat(x); break;
  }
  out[i] = r;
}

template <class F>
static void curve_msm(capgpu_ctx* ctx, const uint64_t* points_xy, const uint64_t* scalars, size_t n, uint64_t* out_xy) {
  int c = 0;
  while (((size_t)2 << c) <= n) c++;  // floor(log2 n)
  c -= 1;
  if (c < 4) c = 4;
  if (c > 13) c = 13;
  const int W = (256 + c - 1) / c;
  const size_t K = (size_t)1 << (c - 1), NB = (size_t)W * K;
  CAPGPU_REQUIRE(n < ((size_t)1 << 31) && (size_t)W * n < ((size_t)1 << 32), "too many points");
  DevBuf d_bases, d_scal, d_digits, d_counts, d_entries, d_buckets, d_wsum, d_out;
  try {
    d_bases.reserve(n * sizeof(G1AffineT<F>));
    d_scal.reserve(n * 32);
    d_digits.reserve(2 * (size_t)W * n * sizeof(int32_t));
    d_counts.reserve((NB + 1) * sizeof(uint32_t));
    d_entries.reserve((size_t)W * n * sizeof(uint32_t));
    d_buckets.reserve(NB * sizeof(G1XyzzT<F>));
    d_wsum.reserve(W * sizeof(G1XyzzT<F>));
    d_out.reserve(sizeof(G1AffineT<F>));
    cudaStream_t st = ctx->stream;
    CAPGPU_CUDA(cudaMemcpyAsync(d_bases.p, points_xy, n * sizeof(G1AffineT<F>), cudaMemcpyHostToDevice, st));
    CAPGPU_CUDA(cudaMemcpyAsync(d_scal.p, scalars, n * 32, cudaMemcpyHostToDevice, st));
    CAPGPU_CUDA(cudaMemsetAsync(d_counts.p, 0, (NB + 1) * sizeof(uint32_t), st));
    int32_t* digits = d_digits.as<int32_t>();
    uint32_t* ranks = reinterpret_cast<uint32_t*>(digits + (size_t)W * n);
    cv_digits<<<ceil_div(n, 128), 128, 0, st>>>(d_scal.as<uint32_t>(), n, c, W, digits, ranks, d_counts.as<uint32_t>(), K);
    CAPGPU_LAUNCH_CHECK(ctx);
    cv_scan<<<1, 1024, 0, st>>>(d_counts.as<uint32_t>(), NB);
    CAPGPU_LAUNCH_CHECK(ctx);
    {
      dim3 grid(ceil_div(n, 256), (unsigned)W);
      cv_scatter<<<grid, 256, 0, st>>>(digits, ranks, n, d_counts.as<uint32_t>(), d_entries.as<uint32_t>(), K);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    cv_accumulate<F><<<ceil_div(NB, 128), 128, 0, st>>>(d_bases.as<G1AffineT<F>>(), d_entries.as<uint32_t>(), d_counts.as<uint32_t>(),
                                                       d_buckets.as<G1XyzzT<F>>(), NB);
    CAPGPU_LAUNCH_CHECK(ctx);
    const unsigned threads = K >= 256 ? 256 : (unsigned)K;
    const size_t smem = threads * sizeof(G1XyzzT<F>);
    CAPGPU_CUDA(cudaFuncSetAttribute(cv_window_reduce<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * (int)sizeof(G1XyzzT<F>)));
    cv_window_reduce<F><<<W, threads, smem, st>>>(d_buckets.as<G1XyzzT<F>>(), K, (uint32_t)(K / threads), d_wsum.as<G1XyzzT<F>>());
    CAPGPU_LAUNCH_CHECK(ctx);
    cv_final<F><<<1, 1, 0, st>>>(d_wsum.as<G1XyzzT<F>>(), W, c, d_out.as<G1AffineT<F>>());
    CAPGPU_LAUNCH_CHECK(ctx);
    CAPGPU_CUDA(cudaMemcpyAsync(out_xy, d_out.p, sizeof(G1AffineT<F>), cudaMemcpyDeviceToHost, st));
    CAPGPU_CUDA(cudaStreamSynchronize(st));
  } catch (...) {
    for (DevBuf* b : {&d_bases, &d_scal, &d_digits, &d_counts, &d_entries, &d_buckets, &d_wsum, &d_out}) b->release();
    throw;
  }
  for (DevBuf* b : {&d_bases, &d_scal, &d_digits, &d_counts, &d_entries, &d_buckets, &d_wsum, &d_out}) b->release();
}

template <class F>
static void curve_fq_op(capgpu_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count) {
  DevBuf da, db, dout;
  try {
    da.reserve(count * sizeof(F));
    dout.reserve(count * sizeof(F));
    CAPGPU_CUDA(cudaMemcpyAsync(da.p, a, count * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    if (b) {
      db.reserve(count * sizeof(F));
      CAPGPU_CUDA(cudaMemcpyAsync(db.p, b, count * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    }
    cv_fq_op<F><<<ceil_div(count, 64), 64, 0, ctx->stream>>>(op, da.as<F>(), b ? db.as<F>() : nullptr, dout.as<F>(), count);
    CAPGPU_LAUNCH_CHECK(ctx);
    CAPGPU_CUDA(cudaMemcpyAsync(out, dout.p, count * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    da.release(); db.release(); dout.release();
    throw;
  }
  da.release(); db.release(); dout.release();
}

}  // namespace capgpu

using namespace capgpu;

extern "C" int capgpu_curve_msm_g1(capgpu_ctx* ctx, int curve, const uint64_t* points_xy, const uint64_t* scalars, size_t n,
                                   uint64_t* out_xy) {
  if (!ctx || !out_xy || ((!points_xy || !scalars) && n)) return CAPGPU_ERR_ARG;
  if (curve != CAPGPU_CURVE_BLS12_381 && curve != CAPGPU_CURVE_BLS12_377) return CAPGPU_ERR_ARG;
  if (n == 0) { memset(out_xy, 0, 96); return CAPGPU_OK; }
  return guarded(ctx, [&] {
    if (curve == CAPGPU_CURVE_BLS12_381) curve_msm<Fq381>(ctx, points_xy, scalars, n, out_xy);
    else curve_msm<Fq377>(ctx, points_xy, scalars, n, out_xy);
  });
}

extern "C" int capgpu_curve_fq_op(capgpu_ctx* ctx, int curve, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count) {
  if (!ctx || !a || !out) return CAPGPU_ERR_ARG;
  if (curve != CAPGPU_CURVE_BLS12_381 && curve != CAPGPU_CURVE_BLS12_377) return CAPGPU_ERR_ARG;
  if (op < 0 || op > 5 || ((op == 0 || op == 3 || op == 4) && !b)) return CAPGPU_ERR_ARG;
  if (count == 0) return CAPGPU_OK;
  return guarded(ctx, [&] {
    if (curve == CAPGPU_CURVE_BLS12_381) curve_fq_op<Fq381>(ctx, op, a, b, out, count);
    else curve_fq_op<Fq377>(ctx, op, a, b, out, count);
  });
}
