// Prototype of a dual-pipe Montgomery product: the multiplication rows a * b_i issue their 8 partial products as
// plain IMAD.WIDE.U32 (no addend: ~2 cycles on the fma pipe) and accumulate them with IADD3.X carry chains on the ALU
// pipe; the reduction rows keep the carry-chained IMAD.WIDE.U32.X form (immediate multiplicands, ~4.2 cycles on the
// fma pipe).  With several warps per scheduler the two pipes overlap (mb_pipes.cu: 8 IMAD + 8 IADD3.X issue at 1.06
// cycles per instruction).  Compared bit for bit and timed against the library's fp_mul.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../../cap_b200/csrc/fp.cuh"
using namespace capgpu;

// X (aligned at this row's base) += even products; Y (aligned one limb higher, shifted down by two) += odd products
template <class PR>
__device__ __forceinline__ void mad_row_alu(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
  uint32_t l[8], h[8];
#pragma unroll
  for (int j = 0; j < 8; j++) mul_wide(l[j], h[j], a[j], bi);
  // odd limbs into the shifted accumulator: newY[2k], newY[2k+1] = (Y[2k+2], Y[2k+3]) + a[2k+1]*bi
  X[0] = add_cc(X[0], Y[1]);
  Y[0] = addc_cc(Y[2], l[1]);
  Y[1] = addc_cc(Y[3], h[1]);
  Y[2] = addc_cc(Y[4], l[3]);
  Y[3] = addc_cc(Y[5], h[3]);
  Y[4] = addc_cc(Y[6], l[5]);
  Y[5] = addc_cc(Y[7], h[5]);
  Y[6] = addc_cc(l[7], 0);
  Y[7] = addc(h[7], 0);
  // even limbs
  X[0] = add_cc(X[0], l[0]);
  X[1] = addc_cc(X[1], h[0]);
  X[2] = addc_cc(X[2], l[2]);
  X[3] = addc_cc(X[3], h[2]);
  X[4] = addc_cc(X[4], l[4]);
  X[5] = addc_cc(X[5], h[4]);
  X[6] = addc_cc(X[6], l[6]);
  X[7] = addc_cc(X[7], h[6]);
  Y[7] = addc(Y[7], 0);
}

template <class PR>
__device__ __forceinline__ Fp<PR> fp_mul2(const Fp<PR>& a, const Fp<PR>& b) {
  uint32_t E[8], O[8];
  mul_wide(E[0], E[1], a.v[0], b.v[0]);
  mul_wide(E[2], E[3], a.v[2], b.v[0]);
  mul_wide(E[4], E[5], a.v[4], b.v[0]);
  mul_wide(E[6], E[7], a.v[6], b.v[0]);
  mul_wide(O[0], O[1], a.v[1], b.v[0]);
  mul_wide(O[2], O[3], a.v[3], b.v[0]);
  mul_wide(O[4], O[5], a.v[5], b.v[0]);
  mul_wide(O[6], O[7], a.v[7], b.v[0]);
  mont_reduce_step<PR>(E, O);
#pragma unroll
  for (int i = 1; i < 8; i += 2) {
    mad_row_alu<PR>(O, E, a.v, b.v[i]);
    mont_reduce_step<PR>(O, E);
    if (i + 1 < 8) {
      mad_row_alu<PR>(E, O, a.v, b.v[i + 1]);
      mont_reduce_step<PR>(E, O);
    }
  }
  Fp<PR> r;
  r.v[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int i = 1; i < 7; i++) r.v[i] = addc_cc(E[i], O[i + 1]);
  r.v[7] = addc(E[7], 0);
  fp_final_sub(r);
  return r;
}

template <int OP>
__global__ void chain(const Fq* in, Fq* out, int n, long long* cycles) {
  const int t = threadIdx.x;
  Fq x = in[t & 31], y = in[32 + (t & 31)];
  __syncthreads();
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) x = fp_mul(x, y);
    if (OP == 1) x = fp_mul2(x, y);
  }
  long long c1 = clock64();
  out[blockIdx.x * blockDim.x + t] = x;
  if (t == 0) *cycles = c1 - c0;
}

__global__ void both(const Fq* in, Fq* o1, Fq* o2) {
  Fq x = in[threadIdx.x], y = in[1024 + threadIdx.x];
  o1[threadIdx.x] = fp_mul(x, y);
  o2[threadIdx.x] = fp_mul2(x, y);
}

int main() {
  static Fq h[2048];
  uint32_t s = 4242;
  for (auto& f : h) { for (int i = 0; i < 8; i++) { s = s * 1664525u + 1013904223u; f.v[i] = s; } f.v[7] &= 0x1fffffffu; }
  // edge operands: 0, 1, q - 1, all-ones low limbs
  const uint32_t ql[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  for (int i = 0; i < 8; i++) { h[0].v[i] = 0; h[1].v[i] = i == 0; h[2].v[i] = ql[i] - (i == 0); h[3].v[i] = i < 7 ? 0xffffffffu : 0x2fffffffu; }
  for (int k = 0; k < 4; k++) { h[1024 + k] = h[2]; h[1028 + k] = h[3]; h[4 + k] = h[k]; }
  Fq *din, *d1, *d2; long long* dc;
  cudaMalloc(&din, sizeof h); cudaMalloc(&d1, 1 << 20); cudaMalloc(&d2, 1 << 20); cudaMalloc(&dc, 8);
  cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
  both<<<1, 1024>>>(din, d1, d2);
  static Fq r1[1024], r2[1024];
  cudaMemcpy(r1, d1, sizeof r1, cudaMemcpyDeviceToHost); cudaMemcpy(r2, d2, sizeof r2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 1024; i++) for (int l = 0; l < 8; l++) if (r1[i].v[l] != r2[i].v[l]) { bad++; break; }
  printf("fp_mul2 vs fp_mul: %d / 1024 differ (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
  for (int op = 0; op < 2; op++) {
    printf("%-36s", op == 0 ? "fp_mul  (all rows IMAD.WIDE.X)" : "fp_mul2 (mul rows IMAD.WIDE + IADD3.X)");
    for (int warps : {4, 8, 16, 32}) {
      if (op == 0) { chain<0><<<1, 32 * warps>>>(din, d1, 512, dc); chain<0><<<1, 32 * warps>>>(din, d1, 512, dc); }
      else { chain<1><<<1, 32 * warps>>>(din, d1, 512, dc); chain<1><<<1, 32 * warps>>>(din, d1, 512, dc); }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("  w/SMSP=%d: %6.0f", warps / 4, (double)c / 512 / (warps / 4.0));
    }
    printf("   cycles per product per scheduler (%s)\n", cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
