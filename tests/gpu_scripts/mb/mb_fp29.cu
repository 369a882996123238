// Prototype: carry-free Montgomery product on 9 x 29-bit limbs (R' = 2^261) for the BN254 base field, against the
// library's 8 x 32-bit CIOS (fp_mul).  Every partial product is accumulated with a plain IMAD.WIDE.U32 into a
// 64-bit column accumulator (<= 18 products of < 2^58 per column: no overflow, no carry flags); mb_pipes.cu
// measured 2.1 cycles per plain IMAD.WIDE against 4.2-4.9 for the carry-chained IMAD.WIDE.U32.X of the CIOS rows.
// Checks the product against the library's (via conversion) and times dependent chains at 1..8 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../../cap_b200/csrc/fp.cuh"
using namespace capgpu;

constexpr uint32_t M29 = (1u << 29) - 1;
struct F29 { uint32_t v[9]; };

// q in 29-bit limbs and -q^-1 mod 2^29 (computed on the host at start-up)
__constant__ uint32_t cP29[9];
__constant__ uint32_t cPinv29;

__device__ __forceinline__ void madw(uint64_t& acc, uint32_t a, uint32_t b) { asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b)); }

template <bool CONSTP>
__device__ __forceinline__ F29 mul29(const F29& a, const F29& b, const uint32_t* p, uint32_t pinv) {
  uint32_t m[9];
  F29 r;
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
#pragma unroll
    for (int i = 0; i <= k; i++) madw(acc, a.v[i], b.v[k - i]);
#pragma unroll
    for (int i = 0; i < k; i++) madw(acc, m[i], p[k - i]);
    m[k] = ((uint32_t)acc * pinv) & M29;
    madw(acc, m[k], p[0]);
    acc >>= 29;
  }
#pragma unroll
  for (int k = 9; k < 17; k++) {
#pragma unroll
    for (int i = k - 8; i <= 8; i++) madw(acc, a.v[i], b.v[k - i]);
#pragma unroll
    for (int i = k - 8; i <= 8; i++) madw(acc, m[i], p[k - i]);
    r.v[k - 9] = (uint32_t)acc & M29;
    acc >>= 29;
  }
  r.v[8] = (uint32_t)acc;
  return r;  // < 1.04 q, limbs < 2^29 (top limb small)
}

template <int OP>
__global__ void chain(const uint32_t* in, uint32_t* out, int n, long long* cycles) {
  const int t = threadIdx.x;
  uint32_t p[9];
  for (int i = 0; i < 9; i++) p[i] = cP29[i];
  F29 a, b;
  for (int i = 0; i < 9; i++) { a.v[i] = in[(t & 31) * 9 + i]; b.v[i] = in[(32 + (t & 31)) * 9 + i]; }
  Fq x, y;
  for (int i = 0; i < 8; i++) { x.v[i] = in[1024 + (t & 31) * 8 + i]; y.v[i] = in[2048 + (t & 31) * 8 + i]; }
  __syncthreads();
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) a = mul29<false>(a, b, p, cPinv29);
    if (OP == 1) x = fp_mul(x, y);
  }
  long long c1 = clock64();
  for (int i = 0; i < 9; i++) out[t * 9 + i] = OP == 0 ? a.v[i] : (i < 8 ? x.v[i] : 0);
  if (t == 0) *cycles = c1 - c0;
}

// one product per thread for the correctness check
__global__ void one(const uint32_t* in, uint32_t* out) {
  uint32_t p[9];
  for (int i = 0; i < 9; i++) p[i] = cP29[i];
  F29 a, b;
  for (int i = 0; i < 9; i++) { a.v[i] = in[threadIdx.x * 18 + i]; b.v[i] = in[threadIdx.x * 18 + 9 + i]; }
  F29 r = mul29<false>(a, b, p, cPinv29);
  for (int i = 0; i < 9; i++) out[threadIdx.x * 9 + i] = r.v[i];
}

// host big-number helpers (unsigned __int128 limbs are enough for a 254-bit check through repeated reduction)
typedef unsigned __int128 u128;
struct Big { uint64_t w[10]; };  // 640 bits
static Big big_from29(const uint32_t* v) { Big r{}; for (int i = 0; i < 9; i++) { int bit = 29 * i; r.w[bit / 64] |= (uint64_t)v[i] << (bit % 64); if (bit % 64 > 35) r.w[bit / 64 + 1] |= (uint64_t)v[i] >> (64 - bit % 64); } return r; }
static Big big_mul(const Big& a, const Big& b) { Big r{}; for (int i = 0; i < 5; i++) { u128 c = 0; for (int j = 0; j < 5; j++) { c += (u128)a.w[i] * b.w[j] + r.w[i + j]; r.w[i + j] = (uint64_t)c; c >>= 64; } r.w[i + 5] = (uint64_t)c; } return r; }
static int big_cmp(const Big& a, const Big& b) { for (int i = 9; i >= 0; i--) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1; return 0; }
static Big big_sub(const Big& a, const Big& b) { Big r; u128 br = 0; for (int i = 0; i < 10; i++) { u128 d = (u128)a.w[i] - b.w[i] - br; r.w[i] = (uint64_t)d; br = (d >> 64) & 1; } return r; }
static Big big_shl(const Big& a, int s) { Big r{}; for (int i = 0; i < 10; i++) { int t = i + s / 64; if (t < 10) { r.w[t] |= a.w[i] << (s % 64); if (s % 64 && t + 1 < 10) r.w[t + 1] |= a.w[i] >> (64 - s % 64); } } return r; }
static Big big_mod(Big a, const Big& p) { for (int s = 380; s >= 0; s--) { Big ps = big_shl(p, s); if (big_cmp(a, ps) >= 0) a = big_sub(a, ps); } return a; }

int main() {
  // q
  Big q{}; const uint32_t ql[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  for (int i = 0; i < 8; i++) q.w[i / 2] |= (uint64_t)ql[i] << (32 * (i % 2));
  uint32_t p29[9];
  for (int i = 0; i < 9; i++) { int bit = 29 * i; uint64_t v = q.w[bit / 64] >> (bit % 64); if (bit % 64 > 35) v |= q.w[bit / 64 + 1] << (64 - bit % 64); p29[i] = (uint32_t)v & M29; }
  uint32_t inv = 1; for (int i = 0; i < 6; i++) inv *= 2 - p29[0] * inv;  // p^-1 mod 2^32
  uint32_t pinv = (0u - inv) & M29;
  cudaMemcpyToSymbol(cP29, p29, sizeof p29); cudaMemcpyToSymbol(cPinv29, &pinv, 4);
  // inputs
  static uint32_t h[4096]; uint32_t s = 99;
  for (auto& w : h) { s = s * 1664525u + 1013904223u; w = s; }
  for (int i = 0; i < 64 * 18; i++) h[i] &= M29;
  for (int e = 0; e < 128; e++) h[e * 9 + 8] &= (1u << 21) - 1;            // values < 2^253 < q
  for (int e = 0; e < 64; e++) { h[1024 + e * 8 + 7] &= 0x1fffffffu; h[2048 + e * 8 + 7] &= 0x1fffffffu; }
  uint32_t *din, *dout; long long* dc;
  cudaMalloc(&din, sizeof h); cudaMalloc(&dout, 1 << 20); cudaMalloc(&dc, 8);
  cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
  // correctness: r * 2^261 = a * b (mod q), r < 2q
  one<<<1, 64>>>(din, dout);
  static uint32_t ho[64 * 9];
  cudaMemcpy(ho, dout, sizeof ho, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int t = 0; t < 64; t++) {
    Big a = big_from29(h + t * 18), b = big_from29(h + t * 18 + 9), r = big_from29(ho + t * 9);
    Big lhs = big_mod(big_shl(r, 261), q), rhs = big_mod(big_mul(a, b), q);
    Big q2 = big_shl(q, 1);
    bool limbs_ok = true; for (int i = 0; i < 8; i++) limbs_ok &= ho[t * 9 + i] <= M29;
    if (big_cmp(lhs, rhs) != 0 || big_cmp(r, q2) >= 0 || !limbs_ok) bad++;
  }
  printf("mul29 correctness: %d / 64 wrong (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
  for (int op = 0; op < 2; op++) {
    printf("%-28s", op == 0 ? "mul29 (9 x 29-bit, carry-free)" : "fp_mul (8 x 32-bit CIOS)");
    for (int warps : {4, 8, 16, 32}) {
      if (op == 0) { chain<0><<<1, 32 * warps>>>(din, dout, 256, dc); chain<0><<<1, 32 * warps>>>(din, dout, 256, dc); }
      else { chain<1><<<1, 32 * warps>>>(din, dout, 256, dc); chain<1><<<1, 32 * warps>>>(din, dout, 256, dc); }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("  w/SMSP=%d: %6.0f", warps / 4, (double)c / 256 / (warps / 4.0));
    }
    printf("   cycles per product per scheduler (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
