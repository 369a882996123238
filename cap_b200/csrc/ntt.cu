// Radix-2 NTT / INTT / coset NTT over BN254 Fr for sm_100a.
//
// Replaces ark-poly 0.3.0 `Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}` (natural
// order in and out, generator omega_n = 5^((r-1)/n), coset shift 5), the transforms behind
// jf-relation's compute_wire_polynomials / compute_prod_permutation_polynomial and jf-plonk's
// compute_quotient_polynomial, all reached from /root/reference/src/proof/transfer.rs:181.
//
// Decomposition: n = R*C (R = 2^ceil(log n / 2) <= 1024, C = n/R).  With j = j1*C + j2 and
// i = i1 + R*i2,
//     X[i1 + R*i2] = sum_j2 w_C^(j2 i2) * [ w_n^(j2 i1) * sum_j1 x[j1*C + j2] w_R^(j1 i1) ].
// Pass 1 runs the R-point column transforms (G adjacent columns per CTA, 1024 elements in
// shared memory, limb-planar), multiplies by the `mid` table (w_n^(j2 i1), with the coset
// power g^j2 and the 1/n of an inverse transform folded in) and stores in place; pass 2 runs
// the C-point row transforms and writes the natural-order result.  The coset shift of the
// input is the R-entry `pre` table ((g^C)^j1); the g^-i / n scaling of coset_ifft is the
// `post` table.  n <= 1024 is a single pass.  Inputs shorter than n are zero-extended on
// load (the quotient-domain transforms read only n+3 coefficients).  The prover's quotient
// domain is NOT ark-poly's 8n-point coset but the 6n points g<rho> taken as three 2n-point
// cosets (ntt3_forward / ntt3_inverse at the end of this file): same quotient polynomial.
//
// Roofline: a size-n transform costs (n/2) log2 n butterflies = one Montgomery product each
// (136 IMAD.WIDE) plus <= 2 table products per element, against 2 * 32 B * n of HBM traffic
// per pass; at the prover's sizes it is bound by the integer pipe, not HBM.
#include "common.cuh"
#include <stdlib.h>

namespace capgpu {

static __device__ __constant__ uint32_t kRoot28[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
static __device__ __constant__ uint32_t kRoot28Inv[8] = {0x584bb683u, 0x89bcc016u, 0x0164a50cu, 0xe8d9887fu, 0x795eda3du, 0x755e95cbu, 0x1323b130u, 0x0f572b87u};
static __device__ __constant__ uint32_t kGen[8] = {0x9fffffe6u, 0x1b0d0ef9u, 0xa32a913fu, 0xeaba68a3u, 0xd8dd0689u, 0x47d8eb76u, 0x20f5bbc3u, 0x15d00855u};
static __device__ __constant__ uint32_t kGenInv[8] = {0x09999999u, 0xd7453974u, 0x83c3efa8u, 0xb4ada7d4u, 0xe57f3161u, 0xc49ca2f8u, 0xac156cb3u, 0x162a3754u};
// rho = 5^((r-1) / (3 * 2^28)): rho^3 = kRoot28; rho^(2^(28-L)) has order 3 * 2^L (the 3-coset quotient domain below)
static __device__ __constant__ uint32_t kRho28[8] = {0x938bb649u, 0x70f4a36du, 0xf4178187u, 0xb81a7492u, 0x69f2125eu, 0x22e5d044u, 0x6fdced23u, 0x0a0b8416u};
static __device__ __constant__ uint32_t kRho28Inv[8] = {0x2be48eafu, 0x3b5f7128u, 0xd733e268u, 0x74742b2bu, 0xe4d48d21u, 0x5633d495u, 0x2da8cb77u, 0x23355141u};
static __device__ __constant__ uint32_t kInv3[8] = {0x15555554u, 0xfad2b890u, 0x5db369e8u, 0x75101f9fu, 0x53538a2eu, 0xb4ea4db7u, 0xd3bdd51du, 0x14cf9766u};
static __device__ __constant__ uint32_t kInv2[8] = {0x1ffffffeu, 0x783c14d8u, 0x0c8d1eddu, 0xaf982f6fu, 0xfcfd4f45u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};

struct NttVariant {
  Fr* pre = nullptr;
  Fr* mid = nullptr;
  Fr* post = nullptr;
  Fr* rc = nullptr;  // 3-coset inverse only: zeta^-1, g^-N, g^-2N of the radix-3 step
  bool built = false;
};

struct NttDomain {
  unsigned log_n = 0, log_r = 0, log_c = 0;
  size_t n = 0;
  Fr* tile_tw[2] = {nullptr, nullptr};  // w_1024^k and w_1024^-k, k < 512
  NttVariant var[2][2];                 // [inverse][coset]
  NttVariant var3[2];                   // [inverse]: the three cosets g rho^k H of the 3 * 2^log_n domain, tables [3][..]
  Fr* omega_pows = nullptr;             // w_n^j, j < n (built on demand)
};

__device__ __forceinline__ Fr load_const(const uint32_t* c) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c[i];
  return r;
}

// base^(2^k)
__device__ inline Fr sqr_times(Fr x, unsigned k) {
  for (unsigned i = 0; i < k; i++) x = fp_sqr(x);
  return x;
}

// Table element kinds.  Every entry is one or two small powers, computed independently per
// thread (tables are built once per domain size and cached in the ctx).
enum TableKind {
  TBL_TILE_FWD = 0,   // w_1024^k
  TBL_TILE_INV = 1,   // w_1024^-k
  TBL_OMEGA = 2,      // w_n^j
  TBL_MID = 3,        // [g^j2] * w_n^(+-j2*i1) [* 1/n]   index = i1*C + j2
  TBL_PRE = 4,        // (g^C)^j1
  TBL_POST = 5,       // [g^-i] / n
};

// coset shift of a table: none (coset 0), g (coset 1), or g * rho^k with rho of order 3 * 2^log_n (coset 2); `inv` gives its inverse
__device__ inline Fr ntt_shift(int coset, int k, unsigned log_n, bool inv) {
  Fr s = load_const(inv ? kGenInv : kGen);
  if (coset == 2 && k > 0) {
    Fr rho = sqr_times(load_const(inv ? kRho28Inv : kRho28), 28 - log_n);
    s = fp_mul(s, k == 1 ? rho : fp_sqr(rho));
  }
  return s;
}

__global__ void ntt_build_table(Fr* out, size_t count, int kind, unsigned log_n, unsigned log_c, int inverse, int coset, int k) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  Fr root = load_const(inverse ? kRoot28Inv : kRoot28);
  Fr v;
  switch (kind) {
    case TBL_TILE_FWD:
    case TBL_TILE_INV: {
      Fr w = sqr_times(load_const(kind == TBL_TILE_INV ? kRoot28Inv : kRoot28), 28 - 10);
      v = fp_pow_u64(w, idx);
      break;
    }
    case TBL_OMEGA: {
      Fr w = sqr_times(load_const(kRoot28), 28 - log_n);
      v = fp_pow_u64(w, idx);
      break;
    }
    case TBL_MID: {
      uint64_t c_mask = ((uint64_t)1 << log_c) - 1;
      uint64_t j2 = idx & c_mask, i1 = idx >> log_c;
      Fr w = sqr_times(root, 28 - log_n);
      v = fp_pow_u64(w, j2 * i1);
      if (coset && !inverse) v = fp_mul(v, fp_pow_u64(ntt_shift(coset, k, log_n, false), j2));
      if (inverse && !coset) v = fp_mul(v, fp_pow_u64(load_const(kInv2), log_n));
      break;
    }
    case TBL_PRE: {
      Fr gc = sqr_times(ntt_shift(coset, k, log_n, false), log_c);
      v = fp_pow_u64(gc, idx);
      break;
    }
    case TBL_POST: {
      v = fp_pow_u64(load_const(kInv2), log_n);
      if (coset) v = fp_mul(v, fp_pow_u64(ntt_shift(coset, k, log_n, true), idx));
      if (coset == 2) v = fp_mul(v, load_const(kInv3));  // the 1/3 of the radix-3 recombination
      break;
    }
    default:
      v = Fr::zero();
  }
  out[idx] = v;
}

static Fr* build_table(capgpu_ctx* ctx, size_t count, int kind, unsigned log_n, unsigned log_c, int inverse, int coset) {
  const int copies = coset == 2 ? 3 : 1;  // [3][count]: one table per coset g rho^k H
  Fr* p = nullptr;
  CAPGPU_CUDA(cudaMalloc(&p, copies * count * sizeof(Fr)));
  for (int k = 0; k < copies; k++) {
    ntt_build_table<<<ceil_div(count, 128), 128, 0, ctx->stream>>>(p + (size_t)k * count, count, kind, log_n, log_c, inverse, coset, k);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  return p;
}

NttDomain* get_domain(capgpu_ctx* ctx, unsigned log_n) {
  auto it = ctx->domains.find(log_n);
  if (it != ctx->domains.end()) return it->second;
  CAPGPU_REQUIRE(log_n >= 1 && log_n <= 20, "NTT size must be 2^1 .. 2^20");
  NttDomain* d = new NttDomain();
  d->log_n = log_n;
  d->n = (size_t)1 << log_n;
  if (log_n <= 10) { d->log_r = log_n; d->log_c = 0; }
  else { d->log_r = (log_n + 1) / 2; d->log_c = log_n - d->log_r; }
  d->tile_tw[0] = build_table(ctx, 512, TBL_TILE_FWD, log_n, 0, 0, 0);
  d->tile_tw[1] = build_table(ctx, 512, TBL_TILE_INV, log_n, 0, 1, 0);
  ctx->domains[log_n] = d;
  return d;
}

// constants of ntt3_recombine_kernel: zeta = rho^N is the same primitive cube root of unity for every N = 2^log_n
__global__ void ntt3_consts_kernel(Fr* out, unsigned log_n) {
  const Fr gi1 = sqr_times(load_const(kGenInv), log_n);
  out[0] = sqr_times(load_const(kRho28Inv), 28);
  out[1] = gi1;
  out[2] = fp_sqr(gi1);
}

static void build_variant(capgpu_ctx* ctx, NttDomain* d, bool inverse, int coset) {
  NttVariant& v = coset == 2 ? d->var3[inverse] : d->var[inverse][coset];
  if (v.built) return;
  bool two_pass = d->log_c > 0;
  // the inverse transform's mid table carries no shift: one copy serves the three cosets
  if (two_pass) v.mid = build_table(ctx, d->n, TBL_MID, d->log_n, d->log_c, inverse, inverse && coset == 2 ? 1 : coset);
  if (coset && !inverse) v.pre = build_table(ctx, (size_t)1 << d->log_r, TBL_PRE, d->log_n, d->log_c, 0, coset);
  if (inverse && (coset || !two_pass)) v.post = build_table(ctx, d->n, TBL_POST, d->log_n, d->log_c, 1, coset);
  if (inverse && coset == 2) {
    CAPGPU_CUDA(cudaMalloc(&v.rc, 3 * sizeof(Fr)));
    ntt3_consts_kernel<<<1, 1, 0, ctx->stream>>>(v.rc, d->log_n);
    CAPGPU_LAUNCH_CHECK(ctx);
  }
  v.built = true;
}

const Fr* domain_omega_powers(capgpu_ctx* ctx, unsigned log_n) {
  NttDomain* d = get_domain(ctx, log_n);
  if (!d->omega_pows) d->omega_pows = build_table(ctx, d->n, TBL_OMEGA, log_n, 0, 0, 0);
  return d->omega_pows;
}

void destroy_domain(NttDomain* d) {
  if (!d) return;
  for (int i = 0; i < 2; i++) {
    if (d->tile_tw[i]) cudaFree(d->tile_tw[i]);
    for (int c = 0; c < 3; c++) {
      NttVariant& v = c == 2 ? d->var3[i] : d->var[i][c];
      if (v.pre) cudaFree(v.pre);
      if (v.mid) cudaFree(v.mid);
      if (v.post) cudaFree(v.post);
      if (v.rc) cudaFree(v.rc);
    }
  }
  if (d->omega_pows) cudaFree(d->omega_pows);
  delete d;
}

// ------------------------------------------------------------------------------------------
// tile kernel
// ------------------------------------------------------------------------------------------
struct NttPass {
  const Fr* src;
  Fr* dst;
  size_t src_stride, dst_stride;
  uint32_t src_len;
  uint32_t log_t, log_g;
  uint32_t in_p_stride, in_g_stride, out_q_stride, out_g_stride;
  const Fr* pre;
  const Fr* post;
  const Fr* tw;
  // 3-coset transforms: row y of the batch reads input y / src_div and uses table copy y % tbl_mod
  uint32_t src_div = 1, tbl_mod = 1;
  uint32_t pre_kstride = 0, post_kstride = 0;
};

// input row and table copy of batch row y
__device__ __forceinline__ void ntt_row_setup(const NttPass& P, uint32_t y, const Fr*& src, const Fr*& pre, const Fr*& post) {
  src = P.src + (size_t)(P.src_div > 1 ? y / P.src_div : y) * P.src_stride;
  const uint32_t k = P.tbl_mod > 1 ? y % P.tbl_mod : 0;
  pre = P.pre ? P.pre + (size_t)k * P.pre_kstride : nullptr;
  post = P.post ? P.post + (size_t)k * P.post_kstride : nullptr;
}

__device__ __forceinline__ Fr smem_load(const uint32_t* s, uint32_t plane, uint32_t e) {
  Fr r;
#pragma unroll
  for (int l = 0; l < 8; l++) r.v[l] = s[l * plane + e];
  return r;
}
__device__ __forceinline__ void smem_store(uint32_t* s, uint32_t plane, uint32_t e, const Fr& x) {
#pragma unroll
  for (int l = 0; l < 8; l++) s[l * plane + e] = x.v[l];
}

__global__ void __launch_bounds__(512) ntt_tile_kernel(NttPass P) {
  extern __shared__ uint32_t smem[];
  const uint32_t T = 1u << P.log_t, G = 1u << P.log_g, E = T * G;
  const uint32_t PAD = G > 1 ? (32u / G ? 32u / G : 1u) : 0u;
  const uint32_t TS = T + PAD;
  const uint32_t plane = G * TS;
  const uint32_t gbase = blockIdx.x * G;
  const Fr *src, *pre, *post;
  ntt_row_setup(P, blockIdx.y, src, pre, post);
  Fr* dst = P.dst + (size_t)blockIdx.y * P.dst_stride;

  for (uint32_t ld = threadIdx.x; ld < E; ld += blockDim.x) {
    uint32_t p, g;
    if (P.in_p_stride == 1) { p = ld & (T - 1); g = ld >> P.log_t; }
    else { g = ld & (G - 1); p = ld >> P.log_g; }
    uint32_t idx = p * P.in_p_stride + (gbase + g) * P.in_g_stride;
    Fr x;
    if (idx < P.src_len) {
      x = src[idx];
      if (pre) x = fp_mul(x, pre[p]);
    } else {
      x = Fr::zero();
    }
    smem_store(smem, plane, g * TS + p, x);
  }
  __syncthreads();

  for (uint32_t s = 0; s < P.log_t; s++) {
    const uint32_t log_m = P.log_t - 1 - s;
    const uint32_t m = 1u << log_m;
    for (uint32_t t = threadIdx.x; t < E / 2; t += blockDim.x) {
      uint32_t g = t >> (P.log_t - 1);
      uint32_t tt = t & (T / 2 - 1);
      uint32_t k = tt & (m - 1);
      uint32_t p0 = ((tt >> log_m) << (log_m + 1)) + k;
      uint32_t e0 = g * TS + p0, e1 = e0 + m;
      Fr u = smem_load(smem, plane, e0);
      Fr v = smem_load(smem, plane, e1);
      smem_store(smem, plane, e0, fp_add(u, v));
      Fr d = fp_sub(u, v);
      if (log_m > 0) d = fp_mul(d, P.tw[k << (9 - log_m)]);
      smem_store(smem, plane, e1, d);
    }
    __syncthreads();
  }

  for (uint32_t st = threadIdx.x; st < E; st += blockDim.x) {
    uint32_t g = st & (G - 1), q = st >> P.log_g;
    uint32_t p = P.log_t ? (__brev(q) >> (32 - P.log_t)) : 0;
    Fr x = smem_load(smem, plane, g * TS + p);
    uint32_t oidx = q * P.out_q_stride + (gbase + g) * P.out_g_stride;
    if (post) x = fp_mul(x, post[oidx]);
    dst[oidx] = x;
  }
}

// ------------------------------------------------------------------------------------------
// register-radix tile kernel (tiles of 2^7 .. 2^10 elements)
//
// Each thread keeps 8 elements (64 registers) and runs up to three radix-2 stages on them
// without touching memory (a radix-8 / radix-4 / radix-2 group per phase); phases exchange
// elements through shared memory once (limb-planar, index padded by pos/8 so every phase's
// access pattern is bank-conflict free).  Versus one shared-memory round trip and one barrier
// per stage this cuts LDS/STS and barriers ~3x; the butterflies themselves (one Montgomery
// product each) are unchanged.  Same index conventions as ntt_tile_kernel.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int ntt_num_phases(int log_t) { return log_t == 10 ? 4 : 3; }
__host__ __device__ constexpr int ntt_phase_stages(int log_t, int ph) {
  return log_t == 10 ? (ph < 2 ? 3 : 2)
       : log_t == 9 ? 3
       : log_t == 8 ? (ph < 2 ? 3 : 2)
       : /* 7 */ (ph < 1 ? 3 : 2);
}
__host__ __device__ constexpr int ntt_phase_start(int log_t, int ph) {
  int s = 0;
  for (int i = 0; i < ph; i++) s += ntt_phase_stages(log_t, i);
  return s;
}

__device__ __forceinline__ uint32_t ntt_pad(uint32_t pos) { return pos + (pos >> 3); }

// Out-of-line Montgomery product for the register-radix kernel: with ~40 products per thread
// fully inlined the kernel body is ~160 KB of SASS and small launches stall on instruction
// fetch (ncu: no_instruction 5.8 cycles per issue); one shared copy keeps the body in the
// instruction cache.  Arguments and result travel in registers (by value).
__device__ __noinline__ Fr ntt_mul(Fr a, Fr b) { return fp_mul(a, b); }

// position of element `slot` of thread t in the phase starting at stage S with RR stages
template <int LOG_T, int S, int RR>
__device__ __forceinline__ uint32_t ntt_slot_pos(uint32_t t, int slot) {
  constexpr int LOG_STRIDE = LOG_T - S - RR;
  constexpr int NGPT = 8 >> RR;
  const uint32_t u = (uint32_t)slot >> RR, i = (uint32_t)slot & ((1u << RR) - 1u);
  const uint32_t gid = t * NGPT + u;
  const uint32_t hi = gid >> LOG_STRIDE, lo = gid & ((1u << LOG_STRIDE) - 1u);
  return (hi << (LOG_STRIDE + RR)) + lo + (i << LOG_STRIDE);
}

template <int LOG_T, int S, int RR>
__device__ __forceinline__ void ntt_phase_butterflies(Fr (&x)[8], uint32_t t, const Fr* __restrict__ tw) {
  constexpr int LOG_STRIDE = LOG_T - S - RR;
  constexpr int NGPT = 8 >> RR;
#pragma unroll
  for (int u = 0; u < NGPT; u++) {
    const uint32_t gid = t * NGPT + u;
    const uint32_t lo = gid & ((1u << LOG_STRIDE) - 1u);
#pragma unroll
    for (int l = 0; l < RR; l++) {
      const int half = 1 << (RR - 1 - l);
      const int log_m = LOG_STRIDE + RR - 1 - l;
#pragma unroll
      for (int i = 0; i < (1 << RR); i++) {
        if (i & half) continue;
        const int a = (u << RR) + i, b = a + half;
        Fr va = x[a], vb = x[b];
        x[a] = fp_add(va, vb);
        Fr d = fp_sub(va, vb);
        // twiddle w^0 = 1: the whole last stage, and in a tile's last phase (LOG_STRIDE == 0, lo == 0) every butterfly whose
        // slot has k = 0 — known at compile time, 3 of the 8 products of a three-stage last phase
        if (log_m > 0 && !(LOG_STRIDE == 0 && (i & (half - 1)) == 0)) {
          const uint32_t k = lo + ((uint32_t)(i & (half - 1)) << LOG_STRIDE);
          d = ntt_mul(d, tw[k << (9 - log_m)]);
        }
        x[b] = d;
      }
    }
  }
}

// First phase of a transform whose input is zero beyond the first stride (the 8n-point quotient-domain transforms read
// n + 3 coefficients: in the column pass only slot 0 of a thread is loaded).  Butterflies with a zero partner are
// copies, (u, 0) -> (u, u * tw): 1 + 2 + 4 products for the three stages instead of 12, no additions.  Same twiddle
// indices as ntt_phase_butterflies<LOG_T, 0, 3>, so the result is identical.
template <int LOG_T>
__device__ __forceinline__ void ntt_phase0_slot0_only(Fr (&x)[8], uint32_t t, const Fr* __restrict__ tw) {
  constexpr int LOG_STRIDE = LOG_T - 3;
  const uint32_t lo = t & ((1u << LOG_STRIDE) - 1u);
  // stage l (half = 4, 2, 1; log_m = LOG_STRIDE + 2 - l): every butterfly here has i & (half - 1) == 0, so k = lo
  x[4] = ntt_mul(x[0], tw[lo << (9 - (LOG_STRIDE + 2))]);
  {
    const Fr w = tw[lo << (9 - (LOG_STRIDE + 1))];
    x[2] = ntt_mul(x[0], w);
    x[6] = ntt_mul(x[4], w);
  }
  {
    const Fr w = tw[lo << (9 - LOG_STRIDE)];
    x[1] = ntt_mul(x[0], w);
    x[3] = ntt_mul(x[2], w);
    x[5] = ntt_mul(x[4], w);
    x[7] = ntt_mul(x[6], w);
  }
}

template <int LOG_T, int PH>
struct NttPhaseRunner {
  static __device__ __forceinline__ void run(Fr (&x)[8], uint32_t t, const Fr* __restrict__ tw, uint32_t* tile, uint32_t plane,
                                             bool slot0_only = false) {
    constexpr int S = ntt_phase_start(LOG_T, PH);
    constexpr int RR = ntt_phase_stages(LOG_T, PH);
    if constexpr (PH == 0 && RR == 3) {
      if (slot0_only) ntt_phase0_slot0_only<LOG_T>(x, t, tw);
      else ntt_phase_butterflies<LOG_T, S, RR>(x, t, tw);
    } else {
      ntt_phase_butterflies<LOG_T, S, RR>(x, t, tw);
    }
    if constexpr (PH + 1 < ntt_num_phases(LOG_T)) {
      constexpr int S2 = ntt_phase_start(LOG_T, PH + 1);
      constexpr int RR2 = ntt_phase_stages(LOG_T, PH + 1);
#pragma unroll
      for (int e = 0; e < 8; e++) smem_store(tile, plane, ntt_pad(ntt_slot_pos<LOG_T, S, RR>(t, e)), x[e]);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; e++) x[e] = smem_load(tile, plane, ntt_pad(ntt_slot_pos<LOG_T, S2, RR2>(t, e)));
      __syncthreads();
      NttPhaseRunner<LOG_T, PH + 1>::run(x, t, tw, tile, plane);
    }
  }
};

template <int LOG_T>
__global__ void __launch_bounds__(256, 2) ntt_reg_kernel(NttPass P) {
  extern __shared__ uint32_t smem[];
  constexpr uint32_t T = 1u << LOG_T, TPT = T / 8, TP = T + T / 8 + 8;
  const uint32_t G = 1u << P.log_g;
  const uint32_t plane = G * TP;
  uint32_t g, t;
  if (P.in_p_stride == 1) { g = threadIdx.x / TPT; t = threadIdx.x % TPT; }  // rows: lanes walk positions
  else { g = threadIdx.x & (G - 1); t = threadIdx.x >> P.log_g; }           // columns: lanes walk adjacent columns
  const uint32_t gcol = blockIdx.x * G + g;
  const Fr *src, *pre, *post;
  ntt_row_setup(P, blockIdx.y, src, pre, post);
  Fr* dst = P.dst + (size_t)blockIdx.y * P.dst_stride;
  uint32_t* tile = smem + g * TP;

  Fr x[8];
  constexpr int R0 = ntt_phase_stages(LOG_T, 0);
  bool upper_loaded = false;  // any of slots 1..7 inside the input
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const uint32_t pos = ntt_slot_pos<LOG_T, 0, R0>(t, e);
    const uint32_t idx = pos * P.in_p_stride + gcol * P.in_g_stride;
    if (idx < P.src_len) {
      x[e] = src[idx];
      if (pre) x[e] = ntt_mul(x[e], pre[pos]);
      if (e > 0) upper_loaded = true;
    } else {
      x[e] = Fr::zero();
    }
  }
  // zero-extended inputs (src_len <= 1/8 of the columns' length): the first three stages only copy and scale slot 0
  NttPhaseRunner<LOG_T, 0>::run(x, t, P.tw, tile, plane, R0 == 3 && !upper_loaded);
  constexpr int LAST = ntt_num_phases(LOG_T) - 1;
  constexpr int SL = ntt_phase_start(LOG_T, LAST), RL = ntt_phase_stages(LOG_T, LAST);
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const uint32_t pos = ntt_slot_pos<LOG_T, SL, RL>(t, e);
    const uint32_t q = __brev(pos) >> (32 - LOG_T);
    const uint32_t oidx = q * P.out_q_stride + gcol * P.out_g_stride;
    Fr v = x[e];
    if (post) v = ntt_mul(v, post[oidx]);
    dst[oidx] = v;
  }
}

template <int LOG_T>
static void launch_reg_pass(capgpu_ctx* ctx, NttPass p, size_t n, size_t batch) {
  constexpr uint32_t T = 1u << LOG_T, TP = T + T / 8 + 8;
  static const uint32_t log_cta = [] { const char* e = getenv("CAPGPU_NTT_LOG_CTA"); return e ? (uint32_t)atoi(e) : 10u; }();
  uint32_t log_g = (log_cta > (uint32_t)LOG_T ? log_cta : (uint32_t)LOG_T) - LOG_T;  // 2^log_cta elements, 2^log_cta / 8 threads per CTA
  // (1024-element CTAs: 7 x 2^18 is 6.05 CTAs of 2048 per SM — the 7th costs 9 %; measured 0.42 -> 0.38 ms)
  while (((size_t)T << log_g) > n) log_g--;
  p.log_g = log_g;
  const uint32_t G = 1u << log_g;
  size_t smem = (size_t)8 * G * TP * sizeof(uint32_t);
  // per device, idempotent and cheap: set on every launch (contexts of several GPUs may share the process)
  CAPGPU_CUDA(cudaFuncSetAttribute(ntt_reg_kernel<LOG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2048 / T * TP * 4));
  dim3 grid((unsigned)(n / ((size_t)T * G)), (unsigned)batch);
  ProfScope prof(ctx, PROF_NTT, (double)batch * (double)(n / 2) * LOG_T);
  ntt_reg_kernel<LOG_T><<<grid, G * T / 8, smem, ctx->stream>>>(p);
  CAPGPU_LAUNCH_CHECK(ctx);
}

static bool ntt_force_v1() {
  static bool v = getenv("CAPGPU_NTT_V1") != nullptr;
  return v;
}

static void launch_pass(capgpu_ctx* ctx, const NttPass& p, size_t n, size_t batch) {
  if (!ntt_force_v1()) {
    switch (p.log_t) {
      case 7: launch_reg_pass<7>(ctx, p, n, batch); return;
      case 8: launch_reg_pass<8>(ctx, p, n, batch); return;
      case 9: launch_reg_pass<9>(ctx, p, n, batch); return;
      case 10: launch_reg_pass<10>(ctx, p, n, batch); return;
      default: break;
    }
  }
  uint32_t T = 1u << p.log_t, G = 1u << p.log_g, E = T * G;
  uint32_t PAD = G > 1 ? (32u / G ? 32u / G : 1u) : 0u;
  size_t smem = (size_t)8 * G * (T + PAD) * sizeof(uint32_t);
  unsigned threads = E / 2 >= 512 ? 512 : (E / 2 < 32 ? 32 : E / 2);
  dim3 grid((unsigned)(n / E), (unsigned)batch);
  ProfScope prof(ctx, PROF_NTT, (double)batch * (double)(n / 2) * p.log_t);  // units: butterflies
  ntt_tile_kernel<<<grid, threads, smem, ctx->stream>>>(p);
  CAPGPU_LAUNCH_CHECK(ctx);
}

// `cosets` = 1: `batch` transforms (plain or on the coset g H).  `cosets` = 3: every input is transformed on the three cosets
// g rho^k H (k = 0, 1, 2) of the 3 * 2^log_n-point domain; the batch rows are (input, k), output row stride dst_stride.
static void ntt_run(capgpu_ctx* ctx, unsigned log_n, const Fr* src, size_t src_len, size_t src_stride, Fr* dst, size_t dst_stride, Fr* tmp,
                    size_t batch, bool inverse, int coset) {
  NttDomain* d = get_domain(ctx, log_n);
  build_variant(ctx, d, inverse, coset);
  const NttVariant& v = coset == 2 ? d->var3[inverse] : d->var[inverse][coset];
  CAPGPU_REQUIRE(src_len <= d->n, "NTT input longer than the domain");
  if (batch == 0) return;
  const uint32_t R = 1u << d->log_r, C = 1u << d->log_c;
  const uint32_t tbl_mod = coset == 2 ? 3 : 1;
  const uint32_t src_div = coset == 2 && !inverse ? 3 : 1;  // forward: the three cosets share one input
  if (coset == 2) batch *= 3;
  CAPGPU_REQUIRE(batch <= 65535, "too many transforms in one call (the batch is the grid's y dimension)");
  NttPass p;
  p.tw = d->tile_tw[inverse ? 1 : 0];
  p.src_len = (uint32_t)src_len;
  p.tbl_mod = tbl_mod;
  if (d->log_c == 0) {
    p.src = src; p.dst = dst; p.src_stride = src_stride; p.dst_stride = dst_stride;
    p.src_div = src_div;
    p.log_t = log_n; p.log_g = 0;
    p.in_p_stride = 1; p.in_g_stride = 0; p.out_q_stride = 1; p.out_g_stride = 0;
    p.pre = v.pre; p.post = v.post;
    p.pre_kstride = R; p.post_kstride = (uint32_t)d->n;
    launch_pass(ctx, p, d->n, batch);
    return;
  }
  // pass 1: columns
  p.src = src; p.dst = tmp; p.src_stride = src_stride; p.dst_stride = d->n;
  p.src_div = src_div;
  p.log_t = d->log_r; p.log_g = 10 - d->log_r;
  if ((1u << p.log_g) > C) p.log_g = d->log_c;
  p.in_p_stride = C; p.in_g_stride = 1; p.out_q_stride = C; p.out_g_stride = 1;
  p.pre = v.pre; p.post = v.mid;
  p.pre_kstride = R; p.post_kstride = inverse ? 0u : (uint32_t)d->n;
  launch_pass(ctx, p, d->n, batch);
  // pass 2: rows
  p.src = tmp; p.dst = dst; p.src_stride = d->n; p.dst_stride = dst_stride;
  p.src_div = 1;
  p.src_len = (uint32_t)d->n;
  p.log_t = d->log_c; p.log_g = 10 - d->log_c;
  if ((1u << p.log_g) > R) p.log_g = d->log_r;
  p.in_p_stride = 1; p.in_g_stride = C; p.out_q_stride = R; p.out_g_stride = 1;
  p.pre = nullptr; p.post = v.post;
  p.pre_kstride = 0; p.post_kstride = (uint32_t)d->n;
  launch_pass(ctx, p, d->n, batch);
}

void ntt_device(capgpu_ctx* ctx, unsigned log_n, const Fr* src, size_t src_len, size_t src_stride, Fr* dst,
                size_t dst_stride, Fr* tmp, size_t batch, bool inverse, bool coset) {
  ntt_run(ctx, log_n, src, src_len, src_stride, dst, dst_stride, tmp, batch, inverse, coset ? 1 : 0);
}

// ------------------------------------------------------------------------------------------
// The 3 * 2^L-point quotient domain D = g <rho>, rho of order 3N (N = 2^L), as the three cosets s_k H_N, s_k = g rho^k:
// point (k, i) = s_k w_N^i sits at index k * N + i.  A polynomial of degree < 3N is evaluated on D by three N-point coset
// transforms of its (zero-extended) coefficients, and recovered from its values by three inverse coset transforms
//     u_k(X) = t(X) mod (X^N - c_k),  c_k = s_k^N = g^N zeta^k,  zeta = rho^N (a primitive cube root of unity),
// followed by the radix-3 step  t_{j + aN} = (1/3) g^(-aN) sum_k zeta^(-ak) u_k[j].  The PLONK quotient has degree 5n + 7,
// so N = 2n gives 6n points instead of the 8n of the next power of two: 25 % fewer point-wise evaluations and 2n-point
// transforms (16 butterfly stages at n = 2^15 instead of 18).  The result is the same polynomial, coefficient for coefficient.
// ------------------------------------------------------------------------------------------
void ntt3_forward(capgpu_ctx* ctx, unsigned log_n, const Fr* src, size_t src_len, size_t src_stride, Fr* dst, Fr* tmp, size_t batch) {
  ntt_run(ctx, log_n, src, src_len, src_stride, dst, (size_t)1 << log_n, tmp, batch, false, 2);
}

// u_k in rows (b, k) of `t` -> coefficients t_0 .. t_{3N-1} in place; rc = {zeta^-1, g^-N, g^-2N} (the 1/3 is in the inverse
// transforms' post table).  With zeta^-2 = -1 - zeta^-1 the step costs two products plus the two scalings per j.
__global__ void ntt3_recombine_kernel(Fr* t, uint32_t N, const Fr* __restrict__ rc) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const Fr z1 = rc[0], gi1 = rc[1], gi2 = rc[2];
  Fr* row = t + (size_t)blockIdx.y * 3 * N;
  const Fr u0 = row[j], u1 = row[N + j], u2 = row[2 * (size_t)N + j];
  const Fr p1 = fp_mul(z1, u1), q1 = fp_mul(z1, u2);
  // a = 1: u0 + z1 u1 + z2 u2 = u0 + p1 - u2 - q1;   a = 2: u0 + z2 u1 + z1 u2 = u0 - u1 - p1 + q1
  const Fr a1 = fp_sub(fp_add(u0, p1), fp_add(u2, q1));
  const Fr a2 = fp_sub(fp_add(u0, q1), fp_add(u1, p1));
  row[j] = fp_add(fp_add(u0, u1), u2);
  row[N + j] = fp_mul(a1, gi1);
  row[2 * (size_t)N + j] = fp_mul(a2, gi2);
}

void ntt3_inverse(capgpu_ctx* ctx, unsigned log_n, Fr* t, Fr* tmp, size_t batch) {
  const size_t N = (size_t)1 << log_n;
  ntt_run(ctx, log_n, t, N, N, t, N, tmp, batch, true, 2);
  if (batch == 0) return;
  ProfScope prof(ctx, PROF_NTT, (double)batch * 2.0 * (double)N);
  ntt3_recombine_kernel<<<dim3((unsigned)ceil_div(N, 128), (unsigned)batch), 128, 0, ctx->stream>>>(t, (uint32_t)N, get_domain(ctx, log_n)->var3[1].rc);
  CAPGPU_LAUNCH_CHECK(ctx);
}

}  // namespace capgpu
