// Pippenger G1 multi-scalar multiplication over the SRS bases for sm_100a.
//
// Replaces ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul` as called by ark-poly-commit
// 0.3.0 `KZG10::commit` for each of the 13 commitments of a proof (reached from
// /root/reference/src/proof/transfer.rs:181; SRS built at src/proof/mod.rs:59-69).  The
// result (an affine G1 point) is mathematically unique, so it is bit-identical to arkworks'
// regardless of the bucket schedule used here:
//
//   upload   : bases are fixed per SRS, so the window-shifted copies 2^(c*w) * P_i are
//              precomputed once (W tables of n affine points).  All W windows of a scalar then
//              feed ONE set of 2^(c-1) buckets, and no per-window doubling fold remains.
//   recode   : Montgomery -> canonical, signed c-bit digits d in (-2^(c-1), 2^(c-1)],
//              per-bucket histogram (global atomics on 2^(c-1) counters).
//   scan     : exclusive scan of the histogram (one CTA per scalar vector).
//   scatter  : counting-sort scatter of (table index | sign) entries by bucket.
//   accumulate: LPB lanes per bucket walk the bucket's entries with XYZZ mixed additions
//              (8M+2S), then a shuffle tree folds the lanes; buckets are scheduled fullest first
//              and CTAs batch-interleaved; buckets far above the average population (repeated
//              scalars) get a whole CTA each (msm_accumulate_heavy).
//   reduce   : sum_k k*B_k by segmented running sums + small scalar multiples, CTA tree, and a
//              final fold + conversion to affine.  The low-latency schedule (lone MSMs, latency
//              mode) uses short segments and lane-pair cooperative group operations.
// `batch` scalar vectors over the same bases (the 5 wire / 5 split-quotient commitments of a
// round) share every launch.  Also here: SRS upload paths (affine, compressed, synthetic tau),
// the Lagrange commit key (group inverse DFT) and the ad-hoc-bases entry point.
#include "common.cuh"
#include "msm_reduce.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace capgpu {

// ------------------------------------------------------------------------------------------
// SRS upload: window-shifted tables
// ------------------------------------------------------------------------------------------
__global__ void msm_precompute(G1Affine* table, size_t n, int c, int W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = table[i];
  G1XYZZ a = xyzz_from_affine(p);
  for (int w = 1; w < W; w++) {
    for (int k = 0; k < c; k++) a = xyzz_dbl(a);
    G1Affine q = xyzz_to_affine(a);
    table[(size_t)w * n + i] = q;
    a = xyzz_from_affine(q);
  }
}

// powers_of_g[i] = tau^i * g, g = (1, 2)  (KZG10::setup shape; synthetic SRS)
__global__ void msm_setup_powers(G1Affine* table, size_t n, Fr tau) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s = fp_from_mont(fp_pow_u64(tau, i));
  G1Affine g;
  g.x = Fq::one();
  g.y = fp_dbl(Fq::one());
  G1XYZZ acc = G1XYZZ::inf();
  bool started = false;
  for (int b = 253; b >= 0; b--) {
    if (started) acc = xyzz_dbl(acc);
    if ((s.v[b >> 5] >> (b & 31)) & 1) { xyzz_add_mixed(acc, g.x, g.y, false); started = true; }
  }
  table[i] = xyzz_to_affine(acc);
}

// ark-serialize 0.3 compressed G1 (32 bytes: x little-endian, bit 255 = "y is the larger of
// {y, -y}", bit 254 = infinity) -> affine Montgomery.  The form `UniversalSrs` / `ProvingKey` files
// hold their points in (/root/reference/src/parameters.rs:557-592, src/proof/mod.rs:106).
// y = (x^3 + 3)^((q+1)/4) since q = 3 mod 4; flag[0] is set if some x is not on the curve.
__global__ void msm_decompress(const uint32_t* __restrict__ bytes, G1Affine* table, size_t n, uint32_t* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t EXP[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};   // (q+1)/4
  const uint32_t HALF[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};  // (q-1)/2
  Fq xc;
  for (int l = 0; l < 8; l++) xc.v[l] = bytes[i * 8 + l];
  const uint32_t flags = xc.v[7] >> 30;
  xc.v[7] &= 0x3fffffffu;
  G1Affine out;
  if (flags & 1u) {  // infinity
    out.x = Fq::zero(); out.y = Fq::zero();
    table[i] = out;
    return;
  }
  // canonical x must be < q
  bool lt = false;
  for (int l = 7; l >= 0; l--) {
    uint32_t pl = FqParams::p(l);
    if (xc.v[l] != pl) { lt = xc.v[l] < pl; break; }
  }
  Fq x = fp_to_mont(xc);
  Fq three = fp_add(fp_dbl(Fq::one()), Fq::one());
  Fq rhs = fp_add(fp_mul(fp_sqr(x), x), three);
  Fq y = fp_pow(rhs, EXP);
  if (!lt || fp_sqr(y) != rhs) { atomicOr(flag, 1u); out.x = Fq::zero(); out.y = Fq::zero(); table[i] = out; return; }
  // "positive" = canonical y > (q-1)/2
  Fq yc = fp_from_mont(y);
  bool larger = false;
  for (int l = 7; l >= 0; l--) {
    if (yc.v[l] != HALF[l]) { larger = yc.v[l] > HALF[l]; break; }
  }
  if (larger != ((flags & 2u) != 0)) y = fp_neg(y);
  out.x = x; out.y = y;
  table[i] = out;
}

// ------------------------------------------------------------------------------------------
// recode + histogram
// ------------------------------------------------------------------------------------------
// K buckets are handled by this launch: magnitudes lo + 1 .. lo + K (a bucket-range slice of a split MSM;
// lo = 0 and K = 2^(c-1) otherwise); digits outside the slice are dropped here.  The histogram atomic also
// hands every entry its rank inside its bucket, so the scatter needs no second round of atomics:
// position = bucket offset (after the scan) + rank.  WMAX > 0: windows unrolled (W <= WMAX), all the
// atomics of a scalar in flight together before the first rank is stored.
template <int WMAX>
__global__ void msm_recode(const Fr* __restrict__ scalars, size_t n, size_t stride, int mont, int c, int W, int32_t* __restrict__ digits,
                           uint32_t* __restrict__ ranks, uint32_t* counts, size_t K, uint32_t lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t b = blockIdx.y;
  Fr s = scalars[b * stride + i];
  if (mont) s = fp_from_mont(s);
  const uint32_t mask = (1u << c) - 1u;
  const int32_t half = 1 << (c - 1);
  int32_t carry = 0;
  uint32_t* cnt = counts + b * (K + 2);
  auto digit = [&](int w) {
    uint32_t off = (uint32_t)(w * c);
    uint32_t limb = off >> 5, sh = off & 31;
    uint32_t v = 0;
    if (limb < 8) {
      // dynamic limb index without local memory
      uint32_t lo_w = 0, hi_w = 0;
#pragma unroll
      for (int l = 0; l < 8; l++) {
        if (l == (int)limb) { lo_w = s.v[l]; hi_w = l + 1 < 8 ? s.v[l + 1] : 0u; }
      }
      v = lo_w >> sh;
      if (sh + c > 32) v |= hi_w << (32 - sh);
      v &= mask;
    }
    int32_t d = (int32_t)v + carry;
    carry = 0;
    if (d > half) { d -= (1 << c); carry = 1; }
    uint32_t mag = (uint32_t)(d < 0 ? -d : d);
    if (mag <= lo || mag > lo + K) d = 0;
    return d;
  };
  if (WMAX > 0) {
    int32_t d[WMAX > 0 ? WMAX : 1];
    uint32_t r[WMAX > 0 ? WMAX : 1];
#pragma unroll
    for (int w = 0; w < WMAX; w++) {
      d[w] = 0; r[w] = 0;
      if (w < W) {
        d[w] = digit(w);
        if (d[w] != 0) r[w] = atomicAdd(&cnt[(uint32_t)(d[w] < 0 ? -d[w] : d[w]) - lo], 1u);
      }
    }
#pragma unroll
    for (int w = 0; w < WMAX; w++) {
      if (w < W) {
        const size_t slot = (b * W + w) * n + i;
        digits[slot] = d[w];
        ranks[slot] = r[w];
      }
    }
  } else {
    for (int w = 0; w < W; w++) {
      const int32_t d = digit(w);
      const size_t slot = (b * W + w) * n + i;
      digits[slot] = d;
      if (d != 0) ranks[slot] = atomicAdd(&cnt[(uint32_t)(d < 0 ? -d : d) - lo], 1u);
    }
  }
}

// counts[b][0..K+1] -> exclusive offsets in place (entry K+1 = total).  One CTA per vector; every thread
// scans a contiguous run of counters, a block scan joins the runs.  Buckets holding >= heavy_thr entries
// are listed in heavy[b][..] (count in nheavy[b]) for msm_accumulate_heavy.  With make_order the bucket
// schedule of msm_accumulate is built too: ids sorted by population, fullest first (counting sort on the
// clamped size), so that lanes of a warp walk buckets of (nearly) equal length and the tail of the
// accumulation grid is made of the emptiest buckets.
__global__ void __launch_bounds__(1024) msm_scan(uint32_t* counts, uint32_t* order, uint32_t* heavy, uint32_t* nheavy, size_t K,
                                                 uint32_t heavy_thr, int make_order) {
  extern __shared__ uint32_t sc[];  // the K + 2 counters of this vector (loaded and stored coalesced)
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t nh;
  uint32_t* cnt = counts + (size_t)blockIdx.x * (K + 2);
  uint32_t* hv = heavy + (size_t)blockIdx.x * K;
  const size_t total = K + 2;
  {  // K + 2 is even and so is every vector's first index: 8-byte accesses
    const uint2* src = reinterpret_cast<const uint2*>(cnt);
    uint2* dst = reinterpret_cast<uint2*>(sc);
#pragma unroll 4
    for (size_t i = threadIdx.x; i < total / 2; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) nh = 0;  // (entries 0 and K + 1 are zero: the histogram only touches 1 .. K)
  __syncthreads();
  const size_t per = (total + blockDim.x - 1) / blockDim.x;  // odd for K = 2^j >= 2048: conflict-free strides
  const size_t i0 = (size_t)threadIdx.x * per;
  const size_t i1 = i0 + per < total ? i0 + per : total;
  uint32_t sum = 0;
  for (size_t i = i0; i < i1; i++) sum += sc[i];
  // exclusive block scan of the per-thread sums
  uint32_t x = sum;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t ws = (threadIdx.x < (blockDim.x >> 5)) ? warp_sums[threadIdx.x] : 0u;
    uint32_t z = ws;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (threadIdx.x >= o) z += y;
    }
    warp_sums[threadIdx.x] = z - ws;  // exclusive
  }
  __syncthreads();
  uint32_t run = warp_sums[threadIdx.x >> 5] + x - sum;
  for (size_t i = i0; i < i1; i++) {
    const uint32_t v = sc[i];
    sc[i] = run;
    run += v;
    if (i >= 1 && i <= K && v >= heavy_thr) hv[atomicAdd(&nh, 1u)] = (uint32_t)(i - 1);
  }
  __syncthreads();
  {
    const uint2* src = reinterpret_cast<const uint2*>(sc);
    uint2* dst = reinterpret_cast<uint2*>(cnt);
#pragma unroll 4
    for (size_t i = threadIdx.x; i < total / 2; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) nheavy[blockIdx.x] = nh;
  if (!make_order) return;
  uint32_t* ord = order + (size_t)blockIdx.x * K;
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (size_t k = threadIdx.x; k < K; k += blockDim.x) {
    uint32_t sz = sc[k + 2] - sc[k + 1];
    atomicAdd(&hist[255u - (sz > 255u ? 255u : sz)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // exclusive scan of 256 bins by one warp (8 bins per lane)
    uint32_t loc[8], s8 = 0;
    for (int j = 0; j < 8; j++) { loc[j] = hist[threadIdx.x * 8 + j]; s8 += loc[j]; }
    uint32_t z = s8;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
      if (threadIdx.x >= o) z += y;
    }
    uint32_t r2 = z - s8;
    for (int j = 0; j < 8; j++) { hist[threadIdx.x * 8 + j] = r2; r2 += loc[j]; }
  }
  __syncthreads();
  for (size_t k = threadIdx.x; k < K; k += blockDim.x) {
    uint32_t sz = sc[k + 2] - sc[k + 1];
    uint32_t pos = atomicAdd(&hist[255u - (sz > 255u ? 255u : sz)], 1u);
    ord[pos] = (uint32_t)k;
  }
}

__global__ void msm_scatter(const int32_t* digits, const uint32_t* __restrict__ ranks, size_t n, int W, const uint32_t* __restrict__ offsets,
                            uint32_t* entries, size_t K, size_t table_n, size_t base_off, uint32_t lo) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t w = blockIdx.y, b = blockIdx.z;
  const size_t slot = (b * W + w) * n + i;
  int32_t d = digits[slot];
  if (d == 0) return;
  uint32_t k = (d < 0 ? (uint32_t)(-d) : (uint32_t)d) - lo;
  uint32_t pos = offsets[b * (K + 2) + k] + ranks[slot];
  uint32_t idx = (uint32_t)(w * table_n + base_off + i);
  entries[b * ((size_t)W * n) + pos] = idx | (d < 0 ? 0x80000000u : 0u);
}

// ------------------------------------------------------------------------------------------
// bucket accumulation
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ G1XYZZ shfl_down_xyzz(const G1XYZZ& p, int delta, int width) {
  G1XYZZ r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_down_sync(0xffffffffu, p.X.v[i], delta, width);
    r.Y.v[i] = __shfl_down_sync(0xffffffffu, p.Y.v[i], delta, width);
    r.ZZ.v[i] = __shfl_down_sync(0xffffffffu, p.ZZ.v[i], delta, width);
    r.ZZZ.v[i] = __shfl_down_sync(0xffffffffu, p.ZZZ.v[i], delta, width);
  }
  return r;
}

constexpr uint32_t MSM_HEAVY = 255;  // == the population clamp of msm_scan's schedule
__device__ inline G1XYZZ block_reduce_xyzz(G1XYZZ v, G1XYZZ* smem);

template <int LPB, int MINB>
__global__ void __launch_bounds__(128, MINB) msm_accumulate(const G1Affine* __restrict__ table, const uint32_t* __restrict__ entries,
                                                            const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ order,
                                                            G1XYZZ* buckets, size_t K, size_t entries_stride, uint32_t heavy_thr) {
  // grid = (batch, CTAs per vector): CTAs are issued in x-major order, so the fullest buckets of
  // every vector of the batch run first and the emptiest of all vectors form the tail
  const size_t gid = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
  const size_t slot = gid / LPB;
  const uint32_t lane = (uint32_t)(gid % LPB);
  const size_t b = blockIdx.x;
  G1XYZZ acc = G1XYZZ::inf();
  size_t bucket = 0;
  bool heavy = false;
  if (slot < K) {
    bucket = order[b * K + slot];
    const uint32_t* off = offsets + b * (K + 2);
    uint32_t start = off[bucket + 1], end = off[bucket + 2];
    const uint32_t* ent = entries + b * entries_stride;
    heavy = end - start >= heavy_thr;  // left to msm_accumulate_heavy
    if (heavy) end = start;
    uint32_t e = start + lane;
    // the lane's first two entries are both affine: 6 products instead of a copy and a 10-product mixed addition
    if (e < end && end - e > LPB) {
      const uint32_t u0 = ent[e], u1 = ent[e + LPB];
      const G1Affine p0 = table[u0 & 0x7fffffffu], p1 = table[u1 & 0x7fffffffu];
      if (!p0.is_inf() && !p1.is_inf()) {  // (otherwise the loop below takes them one at a time)
        if (xyzz_set_affine2(acc, p0.x, (u0 >> 31) ? fp_neg(p0.y) : p0.y, p1.x, (u1 >> 31) ? fp_neg(p1.y) : p1.y)) e += 2 * LPB;
      }
    }
    for (; e < end; e += LPB) {
      uint32_t u = ent[e];
      G1Affine p = table[u & 0x7fffffffu];
      if (!p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, (u >> 31) != 0);
    }
  }
  if (LPB > 1) {
#pragma unroll
    for (int o = LPB / 2; o > 0; o >>= 1) {
      G1XYZZ other = shfl_down_xyzz(acc, o, LPB);
      xyzz_add(acc, other);
    }
  }
  if (slot < K && lane == 0 && !heavy) buckets[b * K + bucket] = acc;
}

// ------------------------------------------------------------------------------------------
// Flat accumulation (low-latency schedule).  When the whole launch fits the GPU in one wave the
// bucket-per-lane-group kernel finishes with the fullest buckets while most SMs idle (bucket
// populations spread 40..90 around an average of 64 for a lone 2^17 MSM).  Here every thread takes
// the same number S of consecutive SORTED ENTRIES instead, whichever buckets they belong to: a
// bucket that lies inside one chunk is written directly, a chunk's leading / trailing part of a
// bucket that continues in a neighbouring chunk goes to pfirst[t] / plast[t], and
// msm_combine_flat adds the (typically 2-3) parts of every bucket that straddles chunks.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 4) msm_accumulate_flat(const G1Affine* __restrict__ table, const uint32_t* __restrict__ entries,
                                                              const uint32_t* __restrict__ offsets, G1XYZZ* buckets, G1XYZZ* pfirst,
                                                              G1XYZZ* plast, size_t K, size_t entries_stride, uint32_t S, size_t nthreads,
                                                              uint32_t heavy_thr) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t bi = blockIdx.y;
  const uint32_t* off = offsets + bi * (K + 2) + 1;  // off[b] .. off[b + 1] = entries of bucket b
  const uint32_t* ent = entries + bi * entries_stride;
  const uint32_t E = off[K];
  const uint64_t e0l = (uint64_t)t * S;
  if (t >= nthreads || e0l >= E) return;
  const uint32_t e0 = (uint32_t)e0l;
  const uint32_t e1 = e0 + S < E ? e0 + S : E;
  // bucket of the first entry: largest b with off[b] <= e0 (skips empty buckets sharing the offset)
  uint32_t lo = 0, hi = (uint32_t)K;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= e0) lo = mid; else hi = mid;
  }
  uint32_t b = lo, bs = off[b], bend = off[b + 1];
  uint32_t run_s = e0;
  bool skip = bend - bs >= heavy_thr;  // heavy buckets are summed by msm_accumulate_heavy
  G1XYZZ acc = G1XYZZ::inf();
  G1XYZZ* bk = buckets + bi * K;
  // software pipeline: the point of entry e + 1 is fetched before the addition of entry e starts (every table entry
  // of a lone MSM is used once, so the gathers come from DRAM: ncu showed 10 % of the stall samples on the first use
  // of the loaded point)
  uint32_t u = ent[e0];
  G1Affine p = table[u & 0x7fffffffu];
  for (uint32_t e = e0; e < e1; e++) {
    if (e >= bend) {
      // the run [run_s, bend) of bucket b ends inside this chunk
      if (!skip) {
        if (run_s == bs) bk[b] = acc;                       // whole bucket
        else pfirst[bi * nthreads + t] = acc;               // tail of a bucket begun in an earlier chunk (run_s == e0)
      }
      do { b++; bs = bend; bend = off[b + 1]; } while (e >= bend);
      run_s = e;
      skip = bend - bs >= heavy_thr;
      acc = G1XYZZ::inf();
    }
    uint32_t un = u;
    G1Affine pn = p;
    if (e + 1 < e1) {
      un = ent[e + 1];
      pn = table[un & 0x7fffffffu];
    }
    if (!skip && !p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, (u >> 31) != 0);
    u = un;
    p = pn;
  }
  if (!skip) {
    if (run_s == bs && e1 == bend) bk[b] = acc;             // whole bucket ends exactly at the chunk end
    else if (run_s == e0) pfirst[bi * nthreads + t] = acc;  // the chunk lies inside one bucket (or starts it)
    else plast[bi * nthreads + t] = acc;                    // head of a bucket that continues in the next chunk
  }
}

__global__ void msm_combine_flat(const uint32_t* __restrict__ offsets, G1XYZZ* buckets, const G1XYZZ* __restrict__ pfirst,
                                 const G1XYZZ* __restrict__ plast, size_t K, uint32_t S, size_t nthreads, uint32_t heavy_thr) {
  const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t bi = blockIdx.y;
  if (b >= K) return;
  const uint32_t* off = offsets + bi * (K + 2) + 1;
  const uint32_t s = off[b], e = off[b + 1];
  G1XYZZ* bk = buckets + bi * K;
  if (e == s) { bk[b] = G1XYZZ::inf(); return; }
  if (e - s >= heavy_thr) return;
  const uint32_t ts = s / S, te = (e - 1) / S;
  if (ts == te) return;  // written whole by msm_accumulate_flat
  const G1XYZZ* pf = pfirst + bi * nthreads;
  G1XYZZ acc = (s == ts * S) ? pf[ts] : plast[bi * nthreads + ts];
  for (uint32_t t = ts + 1; t <= te; t++) {
    G1XYZZ q = pf[t];
    xyzz_add(acc, q);
  }
  bk[b] = acc;
}

// Buckets holding >= heavy_thr = max(MSM_HEAVY, 8 x the average population) entries (repeated
// scalars: the 0/1-valued cells of a witness column committed in evaluation form all land in
// bucket 1 of window 0) get a whole CTA each instead of LPB lanes.  The schedule lists them first
// (msm_scan clamps populations at 255), so every CTA walks the schedule with a grid stride and
// stops at the first light bucket.  (A window size whose top window holds only a few bits —
// c = 10, 12 — piles n/4 .. n/16 entries on buckets 1..4; the default sizes, c = 15 / 16, leave
// 14 bits there.  Such buckets still take one CTA each: 2^16 points at c = 12 cost 1.5 ms.)
__global__ void __launch_bounds__(128) msm_accumulate_heavy(const G1Affine* __restrict__ table, const uint32_t* __restrict__ entries,
                                                            const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ heavy,
                                                            const uint32_t* __restrict__ nheavy, G1XYZZ* buckets, size_t K,
                                                            size_t entries_stride) {
  __shared__ G1XYZZ smem[32];
  const size_t b = blockIdx.y;
  const uint32_t* off = offsets + b * (K + 2);
  const uint32_t* ent = entries + b * entries_stride;
  const uint32_t nh = nheavy[b];
  for (uint32_t slot = blockIdx.x; slot < nh; slot += gridDim.x) {
    const size_t bucket = heavy[b * K + slot];
    const uint32_t start = off[bucket + 1], end = off[bucket + 2];
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t e = start + threadIdx.x; e < end; e += blockDim.x) {
      uint32_t u = ent[e];
      G1Affine p = table[u & 0x7fffffffu];
      if (!p.is_inf()) xyzz_add_mixed(acc, p.x, p.y, (u >> 31) != 0);
    }
    acc = block_reduce_xyzz(acc, smem);
    if (threadIdx.x == 0) buckets[b * K + bucket] = acc;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// bucket reduction: sum_{k'=0}^{K-1} (k'+1) * B[k']
// ------------------------------------------------------------------------------------------
__device__ inline G1XYZZ block_reduce_xyzz(G1XYZZ v, G1XYZZ* smem /* >= 32 entries */) {
  for (int o = 16; o > 0; o >>= 1) {
    G1XYZZ other = shfl_down_xyzz(v, o, 32);
    xyzz_add(v, other);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nwarps ? smem[lane] : G1XYZZ::inf();
    for (int o = 16; o > 0; o >>= 1) {
      G1XYZZ other = shfl_down_xyzz(v, o, 32);
      xyzz_add(v, other);
    }
  }
  return v;  // valid in thread 0
}

__global__ void __launch_bounds__(128) msm_reduce_segments(const G1XYZZ* __restrict__ buckets, size_t K, uint32_t L,
                                                           G1XYZZ* partials, uint32_t wbase) {
  __shared__ G1XYZZ smem[32];
  const size_t T = K / L;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t b = blockIdx.y;
  G1XYZZ v = G1XYZZ::inf();
  if (t < T) {
    const G1XYZZ* B = buckets + b * K;
    G1XYZZ running = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (size_t k = (t + 1) * L; k-- > t * L;) {
      G1XYZZ q = B[k];
      xyzz_add(running, q);
      xyzz_add(acc, running);
    }
    v = xyzz_mul_small(running, (uint32_t)(t * L) + wbase);  // bucket j of this slice weighs wbase + j + 1
    xyzz_add(v, acc);
  }
  v = block_reduce_xyzz(v, smem);
  if (threadIdx.x == 0) partials[b * gridDim.x + blockIdx.x] = v;
}

// ------------------------------------------------------------------------------------------
// Lane-pair cooperative group operations (low-latency schedule).
// A lone MSM spends ~40 % of its time in the reduction, which is a chain of ~45 dependent group
// operations run by one warp per scheduler; each XYZZ addition is 14 dependent-ish field
// products.  Here two adjacent lanes hold the same operands and each computes one of the two
// independent products of a level of the formula, exchanging results with one shuffle: 7 product
// latencies per addition instead of 14 (5 instead of 9 per doubling).  Same formulas, same
// special cases (decided identically by both lanes).
// ------------------------------------------------------------------------------------------
// tree over the 16 pairs of a warp, then over the warps of the CTA; result valid in thread 0 (and 1)
__device__ inline G1XYZZ block_reduce_xyzz_pair(G1XYZZ v, G1XYZZ* smem /* >= 32 entries */, bool role, uint32_t pmask) {
  for (int o = 16; o > 1; o >>= 1) {
    G1XYZZ other = shfl_down_xyzz(v, o, 32);
    xyzz_add_pair(v, other, role, pmask);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane >> 1) < nwarps ? smem[lane >> 1] : G1XYZZ::inf();
    for (int o = 16; o > 1; o >>= 1) {
      G1XYZZ other = shfl_down_xyzz(v, o, 32);
      xyzz_add_pair(v, other, role, pmask);
    }
  }
  return v;
}

// msm_reduce_segments with one lane PAIR per segment
__global__ void __launch_bounds__(256) msm_reduce_segments_pair(const G1XYZZ* __restrict__ buckets, size_t K, uint32_t L,
                                                                G1XYZZ* partials, uint32_t wbase) {
  __shared__ G1XYZZ smem[32];
  const size_t T = K / L;
  const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const bool role = threadIdx.x & 1;
  const uint32_t pmask = 3u << (threadIdx.x & 30);
  const size_t b = blockIdx.y;
  G1XYZZ v = G1XYZZ::inf();
  if (t < T) {
    const G1XYZZ* B = buckets + b * K;
    G1XYZZ running = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (size_t k = (t + 1) * L; k-- > t * L;) {
      G1XYZZ q = B[k];
      xyzz_add_pair(running, q, role, pmask);
      xyzz_add_pair(acc, running, role, pmask);
    }
    v = xyzz_mul_small_pair(running, (uint32_t)(t * L) + wbase, role, pmask);
    xyzz_add_pair(v, acc, role, pmask);
  }
  v = block_reduce_xyzz_pair(v, smem, role, pmask);
  if (threadIdx.x == 0) partials[b * gridDim.x + blockIdx.x] = v;
}

__global__ void msm_finalize_pair(const G1XYZZ* partials, uint32_t nparts, G1Affine* out) {
  const size_t b = blockIdx.x;
  const bool role = threadIdx.x & 1;
  const uint32_t pmask = 3u << (threadIdx.x & 30);
  G1XYZZ v = G1XYZZ::inf();
  for (uint32_t i = threadIdx.x >> 1; i < nparts; i += 16) {
    G1XYZZ q = partials[b * nparts + i];
    xyzz_add_pair(v, q, role, pmask);
  }
  for (int o = 16; o > 1; o >>= 1) {
    G1XYZZ other = shfl_down_xyzz(v, o, 32);
    xyzz_add_pair(v, other, role, pmask);
  }
  if (threadIdx.x == 0) out[b] = xyzz_to_affine(v);
}

__global__ void msm_finalize(const G1XYZZ* partials, uint32_t nparts, G1Affine* out) {
  const size_t b = blockIdx.x;
  G1XYZZ v = G1XYZZ::inf();
  for (uint32_t i = threadIdx.x; i < nparts; i += 32) {
    G1XYZZ q = partials[b * nparts + i];
    xyzz_add(v, q);
  }
  for (int o = 16; o > 0; o >>= 1) {
    G1XYZZ other = shfl_down_xyzz(v, o, 32);
    xyzz_add(v, other);
  }
  if (threadIdx.x == 0) out[b] = xyzz_to_affine(v);
}

// ------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------
static int ceil_log2(size_t x) { int l = 0; while (((size_t)1 << l) < x) l++; return l; }

struct MsmTuning {
  int red_seg;         // buckets per thread in the segmented reduction (0 = heuristic)
  size_t acc_threads;  // target thread count when choosing lanes per bucket
  unsigned acc_block;  // CTA size of msm_accumulate
  bool pair;           // lane-pair cooperative reduction in the low-latency schedule
  bool flat;           // flat (equal chunks of entries) accumulation for one-wave launches of that schedule
  bool tree;           // row / column / bit-plane bucket reduction (msm_reduce.cuh) for K >= 512
  size_t strip_min;    // buckets per launch from which the single-lane strip form of its first stage is used
  uint32_t flat_smin;  // fewest entries per thread of the flat accumulation
  int window_min;      // smallest automatic window for n >= 2^12
  size_t flat_min_entries;  // fewest sorted entries of a launch for the flat accumulation
};

static const MsmTuning& msm_tuning() {
  static MsmTuning t = [] {
    MsmTuning x{0, 65536, 128, true, true, true, (size_t)1 << 17, 8, 15, (size_t)1 << 16};  // measured: a 2^17-point MSM runs 7 % faster with 2 lanes per bucket than with 4
    if (const char* e = getenv("CAPGPU_RED_SEG")) x.red_seg = atoi(e);
    if (const char* e = getenv("CAPGPU_ACC_THREADS")) x.acc_threads = (size_t)atol(e);
    if (const char* e = getenv("CAPGPU_ACC_BLOCK")) x.acc_block = (unsigned)atoi(e);
    if (const char* e = getenv("CAPGPU_RED_PAIR")) x.pair = atoi(e) != 0;
    if (const char* e = getenv("CAPGPU_ACC_FLAT")) x.flat = atoi(e) != 0;
    if (const char* e = getenv("CAPGPU_RED_TREE")) x.tree = atoi(e) != 0;
    if (const char* e = getenv("CAPGPU_RED_STRIP_MIN")) x.strip_min = (size_t)atol(e);
    if (const char* e = getenv("CAPGPU_FLAT_SMIN")) x.flat_smin = (uint32_t)atoi(e);
    if (const char* e = getenv("CAPGPU_WINDOW_MIN")) x.window_min = atoi(e);
    if (const char* e = getenv("CAPGPU_FLAT_MIN_ENTRIES")) x.flat_min_entries = (size_t)atol(e);
    return x;
  }();
  return t;
}

// Window size c (2^(c-1) buckets, W = ceil(255 / c) window-shifted tables).  Bucket accumulation costs
// n W mixed additions, the bucket reduction ~2.8 full additions per bucket at half the lane efficiency.
// Measured on one lockstep group of 8 proofs at n = 2^15 (profiles/r2_launches_group8.csv): with c = 16
// the reduction took 0.53 of the accumulation time; c = floor(log2 n) balances the two (c = 15 at
// n = 2^15: 6 % more additions, half the buckets).  A lone MSM of 2^12..2^14 points keeps c = 15: smaller
// windows leave only 254 mod c bits for the top window (c = 13: 7, c = 14: 2), whose few buckets then
// collect n / 2^7 .. n / 4 entries each (measured: 2^14 points at c = 14 take 0.59 ms, 0.28 ms at c = 15).
static int choose_window(size_t n_points) {
  int c = ceil_log2(n_points + 1) - 1;  // floor(log2 n)
  if (c > 16) c = 16;
  if (c < 4) c = 4;
  if (n_points >= ((size_t)1 << 12) && c < msm_tuning().window_min) c = msm_tuning().window_min;
  if (const char* e = getenv("CAPGPU_WINDOW_BITS")) { int v = atoi(e); if (v >= 2 && v <= 16) c = v; }  // A/B runs only
  return c;
}


template <int LPB, int MINB>
static void launch_accumulate2(capgpu_ctx* ctx, const capgpu_srs* srs, size_t K, const uint32_t* entries, const uint32_t* offsets,
                               const uint32_t* order, G1XYZZ* buckets, size_t entries_stride, size_t batch, uint32_t heavy_thr) {
  size_t threads = K * LPB;
  const unsigned block = msm_tuning().acc_block;
  dim3 grid((unsigned)batch, ceil_div(threads, (size_t)block));
  msm_accumulate<LPB, MINB><<<grid, block, 0, ctx->stream>>>(srs->table, entries, offsets, order, buckets, K, entries_stride, heavy_thr);
  CAPGPU_LAUNCH_CHECK(ctx);
}

template <int LPB>
static void launch_accumulate(capgpu_ctx* ctx, const capgpu_srs* srs, size_t K, const uint32_t* entries, const uint32_t* offsets,
                              const uint32_t* order, G1XYZZ* buckets, size_t entries_stride, size_t batch, uint32_t heavy_thr) {
  // resident CTAs per SM (register cap) of the one-lane-per-bucket kernel: CAPGPU_ACC_MINB = 4 (108 registers) | 5 (102)
  static const int minb = [] { const char* e = getenv("CAPGPU_ACC_MINB"); return e ? atoi(e) : 5; }();  // measured: 0.986 -> 0.976 ms per proof
  if (LPB == 1 && minb == 5) launch_accumulate2<LPB, 5>(ctx, srs, K, entries, offsets, order, buckets, entries_stride, batch, heavy_thr);
  else launch_accumulate2<LPB, 4>(ctx, srs, K, entries, offsets, order, buckets, entries_stride, batch, heavy_thr);
}

void msm_device(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const Fr* scalars, size_t n, size_t stride,
                size_t batch, bool scalars_mont, G1Affine* out_dev, bool latency, size_t part, size_t parts, G1XYZZ* out_xyzz,
                const PeerOut* peer) {
  if (batch == 0) return;
  PeerOut po;
  po.n = 0;
  if (peer) po = *peer;
  if (base_off + n > srs->n) throw CodeError{CAPGPU_ERR_SRS_TOO_SMALL};
  CAPGPU_REQUIRE(srs->device == ctx->device, "SRS lives on another device");
  CAPGPU_REQUIRE(parts >= 1 && part < parts && srs->K % parts == 0, "bucket-range split must divide the bucket count");
  // bucket-range slice `part` of `parts` (split MSM across GPUs): this launch owns magnitudes lo+1 .. lo+K
  const size_t K = srs->K / parts;
  const uint32_t lo = (uint32_t)(part * K);
  const int W = srs->W, c = srs->c;
  CAPGPU_REQUIRE(!(out_xyzz || po.n) || (srs->K / parts >= 512 && msm_tuning().tree), "XYZZ slice results need at least 512 buckets per slice");
  CAPGPU_REQUIRE(po.n == 0 || (batch == 1 && n > 0), "peer delivery is for one non-empty scalar vector");
  if (n == 0) {
    if (out_xyzz) CAPGPU_CUDA(cudaMemsetAsync(out_xyzz, 0, batch * sizeof(G1XYZZ), ctx->stream));
    else CAPGPU_CUDA(cudaMemsetAsync(out_dev, 0, batch * sizeof(G1Affine), ctx->stream));
    return;
  }
  ctx->msm_digits.reserve(2 * batch * W * n * sizeof(int32_t));
  ctx->msm_counts.reserve((batch * (K + 2) + 2 * batch * K + batch) * sizeof(uint32_t));
  ctx->msm_entries.reserve(batch * W * n * sizeof(uint32_t));
  ctx->msm_buckets.reserve(batch * K * sizeof(G1XYZZ));
  int32_t* digits = ctx->msm_digits.as<int32_t>();
  uint32_t* ranks = reinterpret_cast<uint32_t*>(digits + batch * W * n);
  uint32_t* counts = ctx->msm_counts.as<uint32_t>();
  uint32_t* order = counts + batch * (K + 2);
  uint32_t* heavy = order + batch * K;
  uint32_t* nheavy = heavy + batch * K;
  uint32_t* entries = ctx->msm_entries.as<uint32_t>();
  G1XYZZ* buckets = ctx->msm_buckets.as<G1XYZZ>();

  // lanes per bucket: aim for ~128k accumulating threads
  size_t lpb = 1;
  // ... but never so many that a lane gets fewer than ~8 additions (the lane tree costs log2(lpb) full adds)
  const size_t avg_entries = (size_t)W * n / srs->K;
  while (lpb < 32 && batch * K * lpb * 2 <= msm_tuning().acc_threads && lpb * 2 * 8 <= avg_entries) lpb <<= 1;
  const size_t es = (size_t)W * n;
  const uint32_t heavy_thr = (uint32_t)(8 * avg_entries > MSM_HEAVY ? 8 * avg_entries : MSM_HEAVY);
  // one-wave launches in the low-latency schedule: equal chunks of sorted entries per thread
  const size_t wave_threads = (size_t)ctx->sm_count * 4 * 128;
  const size_t ee = es / parts;  // expected entries of this bucket-range slice
  // (up to two waves of bucket slots: the five-vector launches of a latency-mode proof at n = 2^15 are 81920 slots)
  const bool flat = latency && msm_tuning().flat && batch * K * lpb <= 2 * wave_threads && ee >= msm_tuning().flat_min_entries;

  {
  ProfScope prof_sort(ctx, PROF_MSM_SORT, (double)batch * W * n);
  CAPGPU_CUDA(cudaMemsetAsync(counts, 0, batch * (K + 2) * sizeof(uint32_t), ctx->stream));
  {
    dim3 grid(ceil_div(n, 128), (unsigned)batch);
    if (W <= 17) msm_recode<17><<<grid, 128, 0, ctx->stream>>>(scalars, n, stride, scalars_mont ? 1 : 0, c, W, digits, ranks, counts, K, lo);
    else msm_recode<0><<<grid, 128, 0, ctx->stream>>>(scalars, n, stride, scalars_mont ? 1 : 0, c, W, digits, ranks, counts, K, lo);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  {
    static std::once_flag scan_attr[16];
    std::call_once(scan_attr[ctx->device & 15], [] {
      cudaFuncSetAttribute(msm_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((((size_t)1 << 15) + 2) * sizeof(uint32_t)));
    });
  }
  msm_scan<<<(unsigned)batch, 1024, (K + 2) * sizeof(uint32_t), ctx->stream>>>(counts, order, heavy, nheavy, K, heavy_thr, flat ? 0 : 1);
  CAPGPU_LAUNCH_CHECK(ctx);
  {
    dim3 grid(ceil_div(n, 256), (unsigned)W, (unsigned)batch);
    msm_scatter<<<grid, 256, 0, ctx->stream>>>(digits, ranks, n, W, counts, entries, K, srs->n, base_off, lo);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  }
  // mixed additions of this call = non-zero digits kept by the recode (read back only while profiling:
  // zero witness cells and bucket-range slices make the digit-slot count batch * W * n an overestimate)
  double additions = (double)batch * W * n / parts;
  if (ctx->profile) {
    std::vector<uint32_t> totals(batch);
    CAPGPU_CUDA(cudaMemcpy2DAsync(totals.data(), sizeof(uint32_t), counts + K + 1, (K + 2) * sizeof(uint32_t), sizeof(uint32_t), batch,
                                  cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    additions = 0;
    for (uint32_t t : totals) additions += t;
  }
  // the row / column / bit-plane reduction needs whole rows of 256 buckets
  const bool tree_reduce = K >= 512 && msm_tuning().tree;
  FlatParts fl{nullptr, nullptr, nullptr, 0, heavy_thr, 0};
  if (flat) {
    ProfScope prof_acc(ctx, PROF_MSM_ACCUMULATE, additions);
    const size_t per_vec_threads = wave_threads / batch;
    uint32_t S = (uint32_t)((ee + per_vec_threads - 1) / per_vec_threads);
    if (S < msm_tuning().flat_smin) S = msm_tuning().flat_smin;
    const size_t nthreads = (es + S - 1) / S;  // covers the worst case (every digit in this slice); idle threads exit at once
    ctx->msm_flat.reserve(2 * batch * nthreads * sizeof(G1XYZZ));
    G1XYZZ* pfirst = ctx->msm_flat.as<G1XYZZ>();
    G1XYZZ* plast = pfirst + batch * nthreads;
    {
      dim3 grid(ceil_div(nthreads, 128), (unsigned)batch);
      msm_accumulate_flat<<<grid, 128, 0, ctx->stream>>>(srs->table, entries, counts, buckets, pfirst, plast, K, es, S, nthreads, heavy_thr);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    if (tree_reduce) {
      fl = FlatParts{counts, pfirst, plast, S, heavy_thr, nthreads};  // the chunk partials are added up by msm_red_tiles
    } else {
      dim3 grid(ceil_div(K, 128), (unsigned)batch);
      msm_combine_flat<<<grid, 128, 0, ctx->stream>>>(counts, buckets, pfirst, plast, K, S, nthreads, heavy_thr);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    dim3 grid((unsigned)(K < 64 ? K : 64), (unsigned)batch);
    msm_accumulate_heavy<<<grid, 128, 0, ctx->stream>>>(srs->table, entries, counts, heavy, nheavy, buckets, K, es);
    CAPGPU_LAUNCH_CHECK(ctx);
  } else {
  ProfScope prof_acc(ctx, PROF_MSM_ACCUMULATE, additions);
  switch (lpb) {
    case 1: launch_accumulate<1>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
    case 2: launch_accumulate<2>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
    case 4: launch_accumulate<4>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
    case 8: launch_accumulate<8>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
    case 16: launch_accumulate<16>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
    default: launch_accumulate<32>(ctx, srs, K, entries, counts, order, buckets, es, batch, heavy_thr); break;
  }
  {
    dim3 grid((unsigned)(K < 64 ? K : 64), (unsigned)batch);
    msm_accumulate_heavy<<<grid, 128, 0, ctx->stream>>>(srs->table, entries, counts, heavy, nheavy, buckets, K, es);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  }
  ProfScope prof_red(ctx, PROF_MSM_REDUCE, (double)batch * K);
  if (tree_reduce) {
    // rows / columns / bit planes (msm_reduce.cuh)
    const size_t R = K / RED_COLS;
    // lone MSMs / small batches: lane-quad tiles of 4 x 8 (2 x 16) buckets (shortest chain); many vectors per
    // launch: single-lane strips of 16 buckets (least multiply-pipe time)
    const bool strips = !flat && batch * K >= msm_tuning().strip_min && R >= 2;
    const int log_tr = R >= 4 ? 2 : 1;
    const int log_sl = R >= 16 ? 4 : (R >= 8 ? 3 : (R >= 4 ? 2 : 1));
    const int ncb = strips ? 16 : RED_COLS >> (5 - log_tr);
    const size_t nrb = strips ? R >> log_sl : R >> log_tr;
    const size_t NS = R + RED_COLS;
    const uint32_t row0 = lo / RED_COLS;
    int nplanes = 8;
    for (uint32_t amax = row0 + (uint32_t)R - 1; amax; amax >>= 1) nplanes++;
    if (nplanes < 9) nplanes = 9;
    ctx->msm_partials.reserve(batch * (R * ncb + RED_COLS * nrb + NS + 16) * sizeof(G1XYZZ));
    if (ctx->msm_ticket.bytes < batch * sizeof(uint32_t)) {
      ctx->msm_ticket.reserve((batch < 64 ? 64 : 2 * batch) * sizeof(uint32_t));
      CAPGPU_CUDA(cudaMemsetAsync(ctx->msm_ticket.p, 0, ctx->msm_ticket.bytes, ctx->stream));
    }
    G1XYZZ* rowpart = ctx->msm_partials.as<G1XYZZ>();
    G1XYZZ* colpart = rowpart + batch * R * ncb;
    G1XYZZ* sums = colpart + batch * RED_COLS * nrb;
    G1XYZZ* planes = sums + batch * NS;
    if (strips) {
      dim3 grid(ceil_div(K / 16 + RED_COLS * nrb, 128), (unsigned)batch);
      msm_red_strips<<<grid, 128, 0, ctx->stream>>>(buckets, K, log_sl, rowpart, colpart);
      CAPGPU_LAUNCH_CHECK(ctx);
    } else {
      dim3 grid((unsigned)(K / RED_TILE), (unsigned)batch);
      msm_red_tiles<<<grid, RED_THREADS, 0, ctx->stream>>>(buckets, K, log_tr, rowpart, colpart, fl);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    if (strips) {
      dim3 grid(ceil_div(4 * NS, 128), (unsigned)batch);
      msm_red_sums_lane<<<grid, 128, 0, ctx->stream>>>(rowpart, colpart, R, ncb, (int)nrb, sums);
      CAPGPU_LAUNCH_CHECK(ctx);
    } else {
      dim3 grid(ceil_div(NS, 2), (unsigned)batch);
      msm_red_sums<<<grid, RED_THREADS, 0, ctx->stream>>>(rowpart, colpart, R, ncb, (int)nrb, sums);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    {
      dim3 grid((unsigned)nplanes, (unsigned)batch);
      if (batch * nplanes <= (size_t)ctx->sm_count)
        msm_red_planes<128><<<grid, 512, 0, ctx->stream>>>(sums, R, row0, nplanes, planes, ctx->msm_ticket.as<uint32_t>(), out_dev, out_xyzz, po);
      else
        msm_red_planes<32><<<grid, 128, 0, ctx->stream>>>(sums, R, row0, nplanes, planes, ctx->msm_ticket.as<uint32_t>(), out_dev, out_xyzz, po);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    return;
  }
  // few buckets (small windows): segmented running sums.  L buckets per thread; depth is ~2L + 36 group
  // operations and work ~(K/L)(2L + 36): L = 32 by default; in latency mode aim at ~8192 threads per
  // launch, 4 <= L <= 32.
  uint32_t L = 32;
  if (latency) {
    L = 4;
    while (L < 32 && batch * K / (2 * L) >= 8192) L <<= 1;
  }
  if (msm_tuning().red_seg > 0) L = (uint32_t)msm_tuning().red_seg;
  while (L > 1 && K / L < 32) L >>= 1;
  size_t T = K / L;
  // 128 segments per CTA: one warp per SM sub-partition, and few enough CTAs that each gets its own SM
  unsigned block = T >= 128 ? 128 : (T < 32 ? 32 : (unsigned)T);
  unsigned nblocks = ceil_div(T, block);
  ctx->msm_partials.reserve(batch * nblocks * sizeof(G1XYZZ));
  G1XYZZ* partials = ctx->msm_partials.as<G1XYZZ>();
  const bool pair = latency && msm_tuning().pair;  // lane-pair cooperative group operations
  {
    dim3 grid(nblocks, (unsigned)batch);
    if (pair) msm_reduce_segments_pair<<<grid, 2 * block, 0, ctx->stream>>>(buckets, K, L, partials, lo);
    else msm_reduce_segments<<<grid, block, 0, ctx->stream>>>(buckets, K, L, partials, lo);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  if (pair) msm_finalize_pair<<<(unsigned)batch, 32, 0, ctx->stream>>>(partials, nblocks, out_dev);
  else msm_finalize<<<(unsigned)batch, 32, 0, ctx->stream>>>(partials, nblocks, out_dev);
  CAPGPU_LAUNCH_CHECK(ctx);
}

}  // namespace capgpu

using namespace capgpu;

static int srs_create(capgpu_ctx* ctx, const uint64_t* points_xy, const uint64_t* tau, const uint8_t* compressed, size_t n_points,
                      int window_bits, capgpu_srs** out);

extern "C" int capgpu_srs_upload(capgpu_ctx* ctx, const uint64_t* points_xy, size_t n_points, int window_bits, capgpu_srs** out) {
  if (!ctx || !out || !points_xy) return CAPGPU_ERR_ARG;
  return srs_create(ctx, points_xy, nullptr, nullptr, n_points, window_bits, out);
}

extern "C" int capgpu_srs_setup(capgpu_ctx* ctx, const uint64_t* tau, size_t n_points, int window_bits, capgpu_srs** out) {
  if (!ctx || !out || !tau) return CAPGPU_ERR_ARG;
  return srs_create(ctx, nullptr, tau, nullptr, n_points, window_bits, out);
}

extern "C" int capgpu_srs_upload_compressed(capgpu_ctx* ctx, const uint8_t* bytes, size_t n_points, int window_bits, capgpu_srs** out) {
  if (!ctx || !out || !bytes) return CAPGPU_ERR_ARG;
  return srs_create(ctx, nullptr, nullptr, bytes, n_points, window_bits, out);
}

extern "C" int capgpu_srs_export(capgpu_ctx* ctx, const capgpu_srs* srs, uint64_t* points_xy, size_t n_points) {
  if (!ctx || !srs || !points_xy) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(n_points <= srs->n, "export larger than the SRS");
    CAPGPU_CUDA(cudaMemcpyAsync(points_xy, srs->table, n_points * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

static int srs_create(capgpu_ctx* ctx, const uint64_t* points_xy, const uint64_t* tau, const uint8_t* compressed, size_t n_points,
                      int window_bits, capgpu_srs** out) {
  *out = nullptr;
  capgpu_srs* srs = new capgpu_srs();
  int rc = guarded(ctx, [&] {
    CAPGPU_REQUIRE(n_points >= 1 && n_points <= ((size_t)1 << 21), "SRS size out of range");
    int c = window_bits;
    if (c == 0) c = choose_window(n_points);
    CAPGPU_REQUIRE(c >= 2 && c <= 16, "window_bits must be in [2, 16]");
    srs->device = ctx->device;
    srs->n = n_points;
    srs->c = c;
    srs->W = (255 + c - 1) / c;
    srs->K = (size_t)1 << (c - 1);
    CAPGPU_REQUIRE((size_t)srs->W * n_points < ((size_t)1 << 31), "SRS too large for 31-bit table indices");
    CAPGPU_CUDA(cudaMalloc(&srs->table, (size_t)srs->W * n_points * sizeof(G1Affine)));
    if (points_xy) {
      CAPGPU_CUDA(cudaMemcpyAsync(srs->table, points_xy, n_points * sizeof(G1Affine), cudaMemcpyHostToDevice, ctx->stream));
    } else if (compressed) {
      // stage the 32-byte encodings in the (not yet used) upper part of the table allocation
      uint32_t* stage = reinterpret_cast<uint32_t*>(srs->table + n_points * (srs->W > 1 ? 1 : 0));
      DevBuf tmp;
      if (srs->W == 1) { tmp.reserve(n_points * 32); stage = tmp.as<uint32_t>(); }
      ctx->msm_counts.reserve(sizeof(uint32_t));
      uint32_t* flag = ctx->msm_counts.as<uint32_t>();
      CAPGPU_CUDA(cudaMemsetAsync(flag, 0, sizeof(uint32_t), ctx->stream));
      CAPGPU_CUDA(cudaMemcpyAsync(stage, compressed, n_points * 32, cudaMemcpyHostToDevice, ctx->stream));
      msm_decompress<<<ceil_div(n_points, 64), 64, 0, ctx->stream>>>(stage, srs->table, n_points, flag);
      CAPGPU_LAUNCH_CHECK(ctx);
      uint32_t bad = 0;
      CAPGPU_CUDA(cudaMemcpyAsync(&bad, flag, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
      CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
      tmp.release();
      CAPGPU_REQUIRE(bad == 0, "compressed SRS contains a point that is not on the curve");
    } else {
      Fr t;
      memcpy(t.v, tau, sizeof t.v);
      msm_setup_powers<<<ceil_div(n_points, 64), 64, 0, ctx->stream>>>(srs->table, n_points, t);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    msm_precompute<<<ceil_div(n_points, 64), 64, 0, ctx->stream>>>(srs->table, n_points, srs->c, srs->W);
    CAPGPU_LAUNCH_CHECK(ctx);
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
  });
  if (rc != CAPGPU_OK) {
    if (srs->table) cudaFree(srs->table);
    delete srs;
    return rc;
  }
  *out = srs;
  return CAPGPU_OK;
}

// ------------------------------------------------------------------------------------------
// Lagrange-basis SRS: L_j = L_j(tau)·G = (1/n) sum_i omega^(-ij) P_i, i.e. the inverse
// DFT of the first n monomial points taken in the group.  Radix-2 decimation-in-frequency over
// XYZZ points in global memory (each butterfly multiplies by a 254-bit twiddle: double-and-add),
// bit-reversal folded into the final affine conversion.  One-time cost per proving key.
// Commitments of a polynomial given by its evaluations e_j on H are then sum_j e_j L_j — the same
// group element as the coefficient-form MSM, with scalars that are zero / small wherever the
// witness column is.
// ------------------------------------------------------------------------------------------
namespace capgpu {

__device__ G1XYZZ xyzz_mul_scalar(const G1XYZZ& p, const Fr& k) {  // k canonical
  G1XYZZ r = G1XYZZ::inf();
  bool started = false;
  for (int b = 253; b >= 0; b--) {
    if (started) r = xyzz_dbl(r);
    if ((k.v[b >> 5] >> (b & 31)) & 1) { xyzz_add(r, p); started = true; }
  }
  return r;
}

__global__ void lag_load(const G1Affine* __restrict__ pts, G1XYZZ* A, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = pts[i];
  A[i] = p.is_inf() ? G1XYZZ::inf() : xyzz_from_affine(p);
}

// one DIF stage of the inverse transform: (a, b) -> (a + b, (a - b) * omega^(-i * n/len))
__global__ void __launch_bounds__(64) lag_stage(G1XYZZ* A, size_t n, size_t len, const Fr* __restrict__ omega_pows) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n / 2) return;
  const size_t half = len / 2;
  const size_t blk = k / half, i = k % half;
  const size_t i0 = blk * len + i, i1 = i0 + half;
  G1XYZZ a = A[i0], b = A[i1];
  G1XYZZ sum = a;
  xyzz_add(sum, b);
  b.Y = fp_neg(b.Y);
  xyzz_add(a, b);
  const size_t e = i * (n / len);
  if (e != 0) a = xyzz_mul_scalar(a, fp_from_mont(omega_pows[n - e]));
  A[i0] = sum;
  A[i1] = a;
}

__global__ void __launch_bounds__(64) lag_finish(const G1XYZZ* __restrict__ A, unsigned log_n, Fr n_inv, G1Affine* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >> log_n) return;
  G1XYZZ v = xyzz_mul_scalar(A[i], fp_from_mont(n_inv));
  size_t j = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
  out[j] = xyzz_to_affine(v);
}

// Bases of the evaluation-form wire commitment: [L_0 .. L_{n-1}, P_0, P_1, P_n, P_{n+1}] (the last
// four carry the blinding polynomial (b0 + b1 X)(X^n - 1)).
capgpu_srs* srs_lagrange(capgpu_ctx* ctx, const capgpu_srs* srs, unsigned log_n, const Fr* omega_pows, const Fr& n_inv) {
  const size_t n = (size_t)1 << log_n;
  CAPGPU_REQUIRE(srs->n >= n + 2, "SRS too small for the Lagrange commit key");
  const size_t np = n + 4;
  capgpu_srs* lag = new capgpu_srs();
  G1XYZZ* A = nullptr;
  try {
    int c = choose_window(np);
    lag->device = ctx->device;
    lag->n = np;
    lag->c = c;
    lag->W = (255 + c - 1) / c;
    lag->K = (size_t)1 << (c - 1);
    CAPGPU_REQUIRE((size_t)lag->W * np < ((size_t)1 << 31), "SRS too large for 31-bit table indices");
    CAPGPU_CUDA(cudaMalloc(&lag->table, (size_t)lag->W * np * sizeof(G1Affine)));
    CAPGPU_CUDA(cudaMalloc(&A, n * sizeof(G1XYZZ)));
    lag_load<<<ceil_div(n, 128), 128, 0, ctx->stream>>>(srs->table, A, n);
    CAPGPU_LAUNCH_CHECK(ctx);
    for (size_t len = n; len >= 2; len >>= 1) {
      lag_stage<<<ceil_div(n / 2, 64), 64, 0, ctx->stream>>>(A, n, len, omega_pows);
      CAPGPU_LAUNCH_CHECK(ctx);
    }
    lag_finish<<<ceil_div(n, 64), 64, 0, ctx->stream>>>(A, log_n, n_inv, lag->table);
    CAPGPU_LAUNCH_CHECK(ctx);
    CAPGPU_CUDA(cudaMemcpyAsync(lag->table + n, srs->table, 2 * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx->stream));
    CAPGPU_CUDA(cudaMemcpyAsync(lag->table + n + 2, srs->table + n, 2 * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx->stream));
    msm_precompute<<<ceil_div(np, 64), 64, 0, ctx->stream>>>(lag->table, np, lag->c, lag->W);
    CAPGPU_LAUNCH_CHECK(ctx);
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(A);
  } catch (...) {
    if (A) cudaFree(A);
    if (lag->table) cudaFree(lag->table);
    delete lag;
    throw;
  }
  return lag;
}

}  // namespace capgpu

extern "C" void capgpu_srs_destroy(capgpu_srs* srs) {
  if (!srs) return;
  cudaSetDevice(srs->device);
  if (srs->table) cudaFree(srs->table);
  delete srs;
}

extern "C" size_t capgpu_srs_size(const capgpu_srs* srs) { return srs ? srs->n : 0; }

extern "C" int capgpu_msm_g1(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const uint64_t* scalars, size_t n,
                             size_t batch, int scalars_mont, uint64_t* out_xy) {
  if (!ctx || !srs || !out_xy || (!scalars && n)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    ctx->msm_scalars.reserve((batch * n + 1) * sizeof(Fr));
    ctx->msm_out.reserve(batch * sizeof(G1Affine));
    if (n) CAPGPU_CUDA(cudaMemcpyAsync(ctx->msm_scalars.p, scalars, batch * n * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    msm_device(ctx, srs, base_off, ctx->msm_scalars.as<Fr>(), n, n, batch, scalars_mont != 0, ctx->msm_out.as<G1Affine>(), true);
    CAPGPU_CUDA(cudaMemcpyAsync(out_xy, ctx->msm_out.p, batch * sizeof(G1Affine), cudaMemcpyDeviceToHost, ctx->stream));
    CAPGPU_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

extern "C" int capgpu_msm_g1_dev(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                 size_t batch, int scalars_mont, void* d_out_xy) {
  if (!ctx || !srs || !d_out_xy || (!d_scalars && n)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    msm_device(ctx, srs, base_off, (const Fr*)d_scalars, n, n, batch, scalars_mont != 0, (G1Affine*)d_out_xy, true);
  });
}

extern "C" int capgpu_msm_g1_dev_part(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                      int scalars_mont, size_t part, size_t parts, void* d_out_xy) {
  if (!ctx || !srs || !d_out_xy || (!d_scalars && n)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    msm_device(ctx, srs, base_off, (const Fr*)d_scalars, n, n, 1, scalars_mont != 0, (G1Affine*)d_out_xy, true, part, parts);
  });
}

extern "C" int capgpu_msm_g1_dev_part_xyzz(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                           int scalars_mont, size_t part, size_t parts, void* d_out_xyzz) {
  if (!ctx || !srs || !d_out_xyzz || (!d_scalars && n)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    msm_device(ctx, srs, base_off, (const Fr*)d_scalars, n, n, 1, scalars_mont != 0, nullptr, true, part, parts, (G1XYZZ*)d_out_xyzz);
  });
}

extern "C" int capgpu_g1_sum_xyzz_dev(capgpu_ctx* ctx, const void* d_points_xyzz, size_t count, void* d_out_xy) {
  if (!ctx || !d_out_xy || (!d_points_xyzz && count)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(count <= (1u << 16), "too many points");
    g1_sum_xyzz_kernel<<<1, RED_THREADS, 0, ctx->stream>>>((const G1XYZZ*)d_points_xyzz, (uint32_t)count, (G1Affine*)d_out_xy, nullptr, 0, 0);
    CAPGPU_LAUNCH_CHECK(ctx);
  });
}

extern "C" int capgpu_msm_g1_dev_part_peer(capgpu_ctx* ctx, const capgpu_srs* srs, size_t base_off, const void* d_scalars, size_t n,
                                           int scalars_mont, size_t part, size_t parts, void* const* peer_slots, void* const* peer_flags,
                                           size_t n_peers, uint32_t epoch) {
  if (!ctx || !srs || !d_scalars || !n || !peer_slots || !peer_flags || n_peers == 0 || n_peers > (size_t)MAX_PEERS) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    PeerOut po;
    po.n = (int)n_peers;
    po.epoch = epoch;
    for (size_t i = 0; i < n_peers; i++) {
      CAPGPU_REQUIRE(peer_slots[i] && peer_flags[i], "null peer pointer");
      po.slot[i] = (G1XYZZ*)peer_slots[i];
      po.flag[i] = (uint32_t*)peer_flags[i];
    }
    msm_device(ctx, srs, base_off, (const Fr*)d_scalars, n, n, 1, scalars_mont != 0, nullptr, true, part, parts, nullptr, &po);
  });
}

extern "C" int capgpu_g1_sum_xyzz_wait_dev(capgpu_ctx* ctx, const void* d_points_xyzz, const void* d_flags, size_t flag_stride_bytes,
                                           size_t count, uint32_t epoch, void* d_out_xy) {
  if (!ctx || !d_out_xy || !d_points_xyzz || !d_flags || count == 0 || count > (size_t)MAX_PEERS || flag_stride_bytes % 4) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    g1_sum_xyzz_kernel<<<1, RED_THREADS, 0, ctx->stream>>>((const G1XYZZ*)d_points_xyzz, (uint32_t)count, (G1Affine*)d_out_xy,
                                                            (const uint32_t*)d_flags, (uint32_t)(flag_stride_bytes / 4), epoch);
    CAPGPU_LAUNCH_CHECK(ctx);
  });
}

// MSM over bases that are not a resident SRS (the verifier's aggregated commitment sums,
// /root/reference/src/lib.rs:517 txn_batch_verify -> PlonkKzgSnark::batch_verify): builds the
// window-shifted tables for this one call, runs the same pipeline, frees them.
extern "C" int capgpu_msm_g1_adhoc(capgpu_ctx* ctx, const uint64_t* points_xy, const uint64_t* scalars, size_t n, int scalars_mont,
                                   uint64_t* out_xy) {
  if (!ctx || !out_xy || ((!points_xy || !scalars) && n)) return CAPGPU_ERR_ARG;
  if (n == 0) { memset(out_xy, 0, 64); return CAPGPU_OK; }
  capgpu_srs* tmp = nullptr;
  int rc = srs_create(ctx, points_xy, nullptr, nullptr, n, 0, &tmp);
  if (rc != CAPGPU_OK) return rc;
  rc = capgpu_msm_g1(ctx, tmp, 0, scalars, n, 1, scalars_mont, out_xy);
  capgpu_srs_destroy(tmp);
  return rc;
}

// ------------------------------------------------------------------------------------------
// Sum of a few affine points (the fold of a point-range-split MSM: every GPU contributes the
// 64-byte result of its slice, gathered over NVLink; EC addition is not an NCCL reduction op,
// so it is gather-then-add).  One warp.
// ------------------------------------------------------------------------------------------
namespace capgpu {
__global__ void g1_sum_kernel(const G1Affine* __restrict__ pts, uint32_t count, G1Affine* out) {
  G1XYZZ v = G1XYZZ::inf();
  for (uint32_t i = threadIdx.x; i < count; i += 32) {
    G1Affine p = pts[i];
    if (!p.is_inf()) xyzz_add_mixed(v, p.x, p.y, false);
  }
  for (int o = 16; o > 0; o >>= 1) {
    G1XYZZ other = shfl_down_xyzz(v, o, 32);
    xyzz_add(v, other);
  }
  if (threadIdx.x == 0) *out = xyzz_to_affine(v);
}
}  // namespace capgpu

extern "C" int capgpu_g1_sum_dev(capgpu_ctx* ctx, const void* d_points_xy, size_t count, void* d_out_xy) {
  if (!ctx || !d_out_xy || (!d_points_xy && count)) return CAPGPU_ERR_ARG;
  return guarded(ctx, [&] {
    CAPGPU_REQUIRE(count <= (1u << 20), "too many points");
    g1_sum_kernel<<<1, 32, 0, ctx->stream>>>((const G1Affine*)d_points_xy, (uint32_t)count, (G1Affine*)d_out_xy);
    CAPGPU_LAUNCH_CHECK(ctx);
  });
}
