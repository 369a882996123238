// Bucket reduction of the Pippenger MSM: R = sum_{k'=0}^{K-1} (lo + k' + 1) * B[k'] for K = 256 * R_ROWS
// buckets (K >= 512), as a sum of row / column sums and bit planes instead of running sums.
//
// Write k' = a * 256 + c (row a < R_ROWS, column c < 256) and a' = a + lo / 256.  Then
//     R = 256 * sum_a a' * A_a + sum_c (c + 1) * C_c,     A_a = sum_c B[a, c],  C_c = sum_a B[a, c]
//       = sum_p 2^p * Z_p,   Z_p = sum_{c : bit p of (c + 1)} C_c  +  sum_{a : bit (p - 8) of a'} A_a.
// Every bucket enters exactly two sums (2K additions, the same work as the running-sum form) but
// every sum is a balanced tree, so the dependent chain is log2(K) additions + at most 14 doublings
// instead of ~2 * segment + 36 group operations with only K / segment threads busy:
//
//   msm_red_tiles   one CTA per tile of 32 buckets (TR rows x TC columns, one lane quad per bucket):
//                   [flat accumulation only: adds up the chunk partials of each bucket first]
//                   row sums over the tile's columns and column sums over its rows, two interleaved
//                   trees in shared memory -> rowpart[a][column band], colpart[c][row band]
//   msm_red_sums    finishes A_a and C_c from the <= 32 partials each
//   msm_red_planes  one CTA per bit plane p: tree over the <= 192 terms of Z_p, p doublings; the last
//                   CTA to finish (atomic ticket) adds the <= 16 planes and converts to affine
//
// All group operations are the lane-quad forms of eclane.cuh (4 dependent products per addition, 3
// per doubling); they are warp-collective, so every loop below is uniform per warp and idle quads pass
// the point at infinity.  Small CTAs (128 threads) let the block scheduler balance the tiles.
// Replaces the tail of ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul` (the running-sum loop over
// buckets and the `into_affine` of the result), reached from /root/reference/src/proof/transfer.rs:181;
// the result is the same group element.
#pragma once
#include "eclane.cuh"

namespace capgpu {

constexpr int RED_COLS = 256;      // columns: low 8 bits of the bucket index
constexpr int RED_TILE = 32;       // buckets (= lane quads) per CTA of msm_red_tiles / msm_red_sums
constexpr int RED_THREADS = 4 * RED_TILE;

// chunk partials of the flat accumulation (msm_accumulate_flat); S == 0: buckets[] is final
struct FlatParts {
  const uint32_t* offsets;  // per vector: K + 2 exclusive offsets (offsets[1 + k] .. offsets[2 + k] = entries of bucket k)
  const G1XYZZ* pfirst;
  const G1XYZZ* plast;
  uint32_t S;
  uint32_t heavy_thr;
  size_t nthreads;
};

// Destinations of a split-MSM slice result in the peer-mapped memory of every GPU of the group (NVLink / NVSwitch):
// the last reduction kernel stores its 128-byte XYZZ sum straight into slot[i] of peer i and then releases flag[i]
// (system scope) - the exchange is fused into the kernel that produces the data, no collective launch follows.
constexpr int MAX_PEERS = 16;
struct PeerOut {
  G1XYZZ* slot[MAX_PEERS];
  uint32_t* flag[MAX_PEERS];
  int n;
  uint32_t epoch;
};

static __device__ G1XYZZ g_xyzz_inf;  // zero-initialised: the point at infinity, operand of idle quads

__device__ __forceinline__ G1XYZZ ldcg_xyzz(const G1XYZZ* p) {
  G1XYZZ r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = __ldcg(s + i);
  return r;
}
// the four lanes of a quad hold the same point: lane `role` stores coordinate `role`
__device__ __forceinline__ void st_xyzz_quad(G1XYZZ* p, const G1XYZZ& v, uint32_t role) {
  reinterpret_cast<Fq*>(p)[role] = fq_sel4(role, v.X, v.Y, v.ZZ, v.ZZZ);
}

__global__ void __launch_bounds__(RED_THREADS, 4) msm_red_tiles(const G1XYZZ* __restrict__ buckets, size_t K, int log_tr,
                                                                G1XYZZ* rowpart, G1XYZZ* colpart, FlatParts fl) {
  __shared__ G1XYZZ V[RED_TILE], RS[RED_TILE / 2], CS[RED_TILE / 2];
  const uint32_t role = quad_role();
  const int g = threadIdx.x >> 2;
  const int g0 = (threadIdx.x >> 5) * 8;  // first quad of this warp
  const size_t b = blockIdx.y;
  const int log_tc = 5 - log_tr;
  const int TR = 1 << log_tr, TC = 1 << log_tc;
  const int ncb = RED_COLS >> log_tc;  // column bands
  const int rb = blockIdx.x / ncb, cb = blockIdx.x % ncb;
  const size_t R = K / RED_COLS;
  {
    const int r = g >> log_tc, c = g & (TC - 1);
    const size_t k = ((size_t)rb * TR + r) * RED_COLS + (size_t)cb * TC + c;
    const G1XYZZ* bk = buckets + b * K;
    G1XYZZ v = G1XYZZ::inf();
    uint32_t extra = 0;          // chunk partials to add to v (flat accumulation)
    const G1XYZZ* next = nullptr;
    if (fl.S == 0) {
      v = bk[k];
    } else {
      const uint32_t* off = fl.offsets + b * (K + 2) + 1;
      const uint32_t s = off[k], e = off[k + 1];
      if (e != s) {
        const uint32_t ts = s / fl.S, te = (e - 1) / fl.S;
        if (e - s >= fl.heavy_thr || ts == te) {
          v = bk[k];  // summed whole by msm_accumulate_heavy / by one chunk of msm_accumulate_flat
        } else {
          const G1XYZZ* pf = fl.pfirst + b * fl.nthreads;
          v = (s == ts * fl.S) ? pf[ts] : fl.plast[b * fl.nthreads + ts];
          extra = te - ts;
          next = pf + ts + 1;
        }
      }
    }
    st_xyzz_quad(&V[g], v, role);
    __syncwarp();
    for (uint32_t i = 0; __any_sync(0xffffffffu, i < extra); i++) {
      const bool on = i < extra;
      xyzz_add_quad_mem(&V[g], on ? next + i : &g_xyzz_inf, &V[g], on, role);
      __syncwarp();
    }
  }
  __syncthreads();
  const int levels = log_tc > log_tr ? log_tc : log_tr;
  for (int l = 1; l <= levels; l++) {
    const int lwc = log_tc - l, lwr = log_tr - l;  // log2 of the widths left after this level
    const int nrow = l <= log_tc ? (TR << lwc) : 0;
    const int ncol = l <= log_tr ? (TC << lwr) : 0;
    if (g0 < nrow + ncol) {  // uniform per warp
      const G1XYZZ *px = &g_xyzz_inf, *py = &g_xyzz_inf;
      G1XYZZ* pd = nullptr;
      if (g < nrow) {
        const int r = g >> lwc, j = g & ((1 << lwc) - 1), wc = 1 << lwc;
        const G1XYZZ* src = l == 1 ? &V[r * TC] : &RS[r * (TC / 2)];
        px = src + j; py = src + j + wc; pd = &RS[r * (TC / 2) + j];
      } else if (g < nrow + ncol) {
        const int h = g - nrow;
        const int c = h >> lwr, j = h & ((1 << lwr) - 1), wr = 1 << lwr;
        if (l == 1) { px = &V[j * TC + c]; py = &V[(j + wr) * TC + c]; }
        else { px = &CS[c * (TR / 2) + j]; py = px + wr; }
        pd = &CS[c * (TR / 2) + j];
      }
      xyzz_add_quad_mem(px, py, pd, pd != nullptr, role);
    }
    __syncthreads();
  }
  const size_t nrb = R >> log_tr;  // row bands
  if (g < TR) {
    G1XYZZ x = RS[g * (TC / 2)];
    st_xyzz_quad(rowpart + (b * R + (size_t)rb * TR + g) * ncb + cb, x, role);
  } else if (g < TR + TC) {
    const int c = g - TR;
    G1XYZZ x = CS[c * (TR / 2)];
    st_xyzz_quad(colpart + (b * RED_COLS + (size_t)cb * TC + c) * nrb + rb, x, role);
  }
}

// Throughput form of msm_red_tiles (many vectors per launch, nothing waits for one result): one THREAD per strip
// of 16 buckets, plain single-lane additions (ec.cuh) with every lane of every warp busy - 0.66 of the
// multiply-pipe time of the quad form per addition, at 16 sequential additions of latency.  Threads
// 0 .. K/16 - 1 sum the row strips (a, 16 j .. 16 j + 15) -> rowpart[a][j]; threads K/16 .. 2 K/16 - 1 the
// column strips (SL i .. SL i + SL - 1, c), SL = min(16, R) -> colpart[c][i].
__global__ void __launch_bounds__(128, 4) msm_red_strips(const G1XYZZ* __restrict__ buckets, size_t K, int log_sl, G1XYZZ* rowpart,
                                                         G1XYZZ* colpart) {
  const size_t b = blockIdx.y;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t R = K / RED_COLS;
  const size_t nrow = K / 16, ncol = (size_t)RED_COLS * (R >> log_sl);
  const G1XYZZ* bk = buckets + b * K;
  if (t < nrow) {
    const G1XYZZ* src = bk + t * 16;  // strip j = t % 16 of row a = t / 16
    G1XYZZ acc = src[0];
    for (int i = 1; i < 16; i++) {
      G1XYZZ q = src[i];
      xyzz_add(acc, q);
    }
    rowpart[b * nrow + t] = acc;  // [a][j], 16 strips per row
  } else if (t < nrow + ncol) {
    const size_t u = t - nrow;
    const size_t c = u % RED_COLS, i = u / RED_COLS;  // adjacent threads read adjacent buckets
    const size_t SL = (size_t)1 << log_sl;
    const G1XYZZ* src = bk + (i * SL) * RED_COLS + c;
    G1XYZZ acc = src[0];
    for (size_t r = 1; r < SL; r++) {
      G1XYZZ q = src[r * RED_COLS];
      xyzz_add(acc, q);
    }
    colpart[(b * RED_COLS + c) * (R >> log_sl) + i] = acc;
  }
}

// sums[b][s], s < R: A_s = sum of rowpart[b][s][0 .. ncb);  s >= R: C_(s - R) = sum of colpart[b][s - R][0 .. nrb)
// (ncb, nrb <= 32).  Two sums per CTA, 16 quads per sum.
__global__ void __launch_bounds__(RED_THREADS, 4) msm_red_sums(const G1XYZZ* __restrict__ rowpart, const G1XYZZ* __restrict__ colpart,
                                                               size_t R, int ncb, int nrb, G1XYZZ* sums) {
  __shared__ G1XYZZ T[RED_TILE];
  const uint32_t role = quad_role();
  const int g = threadIdx.x >> 2;
  const int j0 = ((threadIdx.x >> 5) * 8) & 15;  // first quad of this warp within its sum
  const size_t b = blockIdx.y;
  const size_t NS = R + RED_COLS;
  const size_t s = (size_t)blockIdx.x * 2 + (g >> 4);
  const int j = g & 15;
  {
    const G1XYZZ *px = &g_xyzz_inf, *py = &g_xyzz_inf;
    if (s < NS) {
      const G1XYZZ* items = s < R ? rowpart + (b * R + s) * ncb : colpart + (b * RED_COLS + (s - R)) * nrb;
      const int cnt = s < R ? ncb : nrb;
      if (j < cnt) px = items + j;
      if (j + 16 < cnt) py = items + j + 16;
    }
    xyzz_add_quad_mem(px, py, &T[g], true, role);
  }
  __syncthreads();
  for (int w = 8; w >= 1; w >>= 1) {
    if (j0 < w) {  // uniform per warp
      const bool on = j < w;
      xyzz_add_quad_mem(on ? &T[g] : &g_xyzz_inf, on ? &T[g + w] : &g_xyzz_inf, &T[g], on, role);
    }
    __syncthreads();
  }
  if (j == 0 && s < NS) {
    G1XYZZ x = T[g];
    st_xyzz_quad(sums + b * NS + s, x, role);
  }
}

// Throughput form of msm_red_sums: FOUR lanes per sum, plain single-lane additions: lane j adds the partials j,
// j + 4, .. (<= 3 additions for the 16 partials of a row after msm_red_strips), two shuffle levels join the lanes
// (one thread per sum: 15 sequential additions on 320 threads per vector - 110 us per launch of pure latency).
__device__ __forceinline__ G1XYZZ shfl_xor_xyzz(const G1XYZZ& p, int m) {
  G1XYZZ r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.X.v[i] = __shfl_xor_sync(0xffffffffu, p.X.v[i], m);
    r.Y.v[i] = __shfl_xor_sync(0xffffffffu, p.Y.v[i], m);
    r.ZZ.v[i] = __shfl_xor_sync(0xffffffffu, p.ZZ.v[i], m);
    r.ZZZ.v[i] = __shfl_xor_sync(0xffffffffu, p.ZZZ.v[i], m);
  }
  return r;
}
__global__ void __launch_bounds__(128, 4) msm_red_sums_lane(const G1XYZZ* __restrict__ rowpart, const G1XYZZ* __restrict__ colpart,
                                                            size_t R, int ncb, int nrb, G1XYZZ* sums) {
  const size_t b = blockIdx.y;
  const size_t NS = R + RED_COLS;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t s = t >> 2;
  const int j = (int)(t & 3);
  G1XYZZ acc = G1XYZZ::inf();
  if (s < NS) {
    const G1XYZZ* items = s < R ? rowpart + (b * R + s) * ncb : colpart + (b * RED_COLS + (s - R)) * nrb;
    const int cnt = s < R ? ncb : nrb;
    for (int i = j; i < cnt; i += 4) {
      G1XYZZ q = items[i];
      xyzz_add(acc, q);
    }
  }
  // all 32 lanes take part in the shuffles (sums beyond NS carry infinity)
  G1XYZZ o = shfl_xor_xyzz(acc, 1);
  xyzz_add(acc, o);
  o = shfl_xor_xyzz(acc, 2);
  xyzz_add(acc, o);
  if (s < NS && j == 0) sums[b * NS + s] = acc;
}

// Plane p (blockIdx.x) of vector b (blockIdx.y): Z_p, then 2^p * Z_p -> planes[b][p]; the last CTA of a vector
// (ticket[b], self-resetting) folds the planes and writes the affine result (or, for out_xyzz, the XYZZ sum).  QUADS lane quads per CTA: 128 for a
// lone MSM (shortest chain), 32 when many vectors share the launch (four CTAs per SM).
template <int QUADS>
__global__ void __launch_bounds__(4 * QUADS, (QUADS <= 32 ? 4 : 1)) msm_red_planes(const G1XYZZ* __restrict__ sums, size_t R, uint32_t row0,
                                                                                  int nplanes, G1XYZZ* planes, uint32_t* ticket,
                                                                                  G1Affine* out, G1XYZZ* out_xyzz, PeerOut peer) {
  __shared__ G1XYZZ T[QUADS];
  __shared__ uint32_t last_s;
  const uint32_t role = quad_role();
  const int g = threadIdx.x >> 2;
  const int g0 = (threadIdx.x >> 5) * 8;  // first quad of this warp
  const int p = blockIdx.x;
  const size_t b = blockIdx.y;
  const size_t NS = R + RED_COLS;
  const G1XYZZ* A = sums + b * NS;  // row sums
  const G1XYZZ* C = A + R;          // column sums
  {
    // terms of Z_p: the i-th c with bit p of (c + 1) set (128 of them for p < 8, c = 255 alone for p = 8), and
    // row i when bit (p - 8) of a' = row0 + i is set; quad g takes the terms i = g, g + QUADS, ..  A column and a
    // row term meet only at (p = 8, i = 0): that row term is added after the loop.
    auto term = [&](int i) -> const G1XYZZ* {
      if (p < 8) {
        const uint32_t v = (((uint32_t)i >> p) << (p + 1)) | (1u << p) | ((uint32_t)i & ((1u << p) - 1u));
        return C + (v - 1);
      }
      if (p == 8 && i == 0) return C + 255;
      if ((size_t)i < R && (((row0 + (uint32_t)i) >> (p - 8)) & 1u)) return A + i;
      return &g_xyzz_inf;
    };
    const G1XYZZ* px = term(g);
    G1XYZZ x = *px;
    st_xyzz_quad(&T[g], x, role);
    __syncwarp();
    for (int it = 1; it < 128 / QUADS; it++) {  // uniform trip count
      xyzz_add_quad_mem(&T[g], term(g + it * QUADS), &T[g], true, role);
      __syncwarp();
    }
    if (p == 8) {  // uniform per CTA
      xyzz_add_quad_mem(&T[g], (g == 0 && (row0 & 1u)) ? A : &g_xyzz_inf, &T[g], true, role);
    }
  }
  __syncthreads();
  for (int w = QUADS / 2; w >= 1; w >>= 1) {
    if (g0 < w) {  // uniform per warp
      const bool on = g < w;
      xyzz_add_quad_mem(on ? &T[g] : &g_xyzz_inf, on ? &T[g + w] : &g_xyzz_inf, &T[g], on, role);
    }
    __syncthreads();
  }
  if (g0 == 0) {  // warp 0: all of its quads double (the operation is warp-collective), quad 0 keeps the result
    G1XYZZ x = T[0];
    for (int i = 0; i < p; i++) x = xyzz_dbl_quad(x, role);
    if (g == 0) st_xyzz_quad(planes + b * 16 + p, x, role);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_s = atomicAdd(&ticket[b], 1u) == (uint32_t)nplanes - 1u;
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  if (g < 16) {
    G1XYZZ x = g < nplanes ? ldcg_xyzz(planes + b * 16 + g) : G1XYZZ::inf();
    st_xyzz_quad(&T[g], x, role);
  }
  __syncthreads();
  for (int w = 8; w >= 1; w >>= 1) {
    if (g0 < w) {
      const bool on = g < w;
      xyzz_add_quad_mem(on ? &T[g] : &g_xyzz_inf, on ? &T[g + w] : &g_xyzz_inf, &T[g], on, role);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (peer.n > 0) {
      // slice of a split MSM, exchange fused in: peer-mapped stores of the sum, then the flags (release, system scope)
      const G1XYZZ v = T[0];
      for (int i = 0; i < peer.n; i++) *peer.slot[i] = v;
      __threadfence_system();
      for (int i = 0; i < peer.n; i++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer.flag[i]), "r"(peer.epoch) : "memory");
    } else if (out_xyzz) {
      out_xyzz[b] = T[0];  // slice of a split MSM: the fold across GPUs converts once
    } else {
      out[b] = xyzz_to_affine(T[0]);
    }
    ticket[b] = 0;
  }
}


// Fold of the slice results of a split MSM (XYZZ, one per GPU, gathered over NVLink): quad tree + one inversion.
// flags != nullptr: first wait until every peer has delivered its slice for this epoch (acquire loads, system scope;
// gives up after ~2 s so that a missing peer shows up as a wrong result, not as a hung GPU), then read the slots
// around L1 (they were written by other GPUs).
__global__ void __launch_bounds__(RED_THREADS) g1_sum_xyzz_kernel(const G1XYZZ* pts, uint32_t count, G1Affine* out, const uint32_t* flags,
                                                                  uint32_t flag_stride, uint32_t epoch) {
  __shared__ G1XYZZ T[RED_TILE];
  const uint32_t role = quad_role();
  const int g = threadIdx.x >> 2;
  const int g0 = (threadIdx.x >> 5) * 8;
  if (flags) {
    if (threadIdx.x < count) {
      const uint32_t* f = flags + (size_t)threadIdx.x * flag_stride;
      const long long t0 = clock64();
      uint32_t v;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      } while (v != epoch && clock64() - t0 < 4000000000LL);
    }
    __syncthreads();
    __threadfence_system();
  }
  {
    G1XYZZ x = (uint32_t)g < count ? ldcg_xyzz(pts + g) : G1XYZZ::inf();
    st_xyzz_quad(&T[g], x, role);
    __syncwarp();
    for (uint32_t i = g + RED_TILE; __any_sync(0xffffffffu, i < count); i += RED_TILE) {
      xyzz_add_quad_mem(&T[g], i < count ? pts + i : &g_xyzz_inf, &T[g], true, role);
      __syncwarp();
    }
  }
  __syncthreads();
  for (int w = RED_TILE / 2; w >= 1; w >>= 1) {
    if (g0 < w) {
      const bool on = g < w;
      xyzz_add_quad_mem(on ? &T[g] : &g_xyzz_inf, on ? &T[g + w] : &g_xyzz_inf, &T[g], on, role);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = xyzz_to_affine(T[0]);
}

}  // namespace capgpu
