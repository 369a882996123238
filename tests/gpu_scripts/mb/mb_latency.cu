// Latency / throughput micro-benchmark of the field and group operations the latency-bound MSM stages
// are built from (one CTA, 1..32 warps, dependent chains; cycles per operation per thread).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../../cap_b200/csrc/eclane.cuh"
using namespace capgpu;

template <int OP>
__global__ void chain(Fq* io, int n, long long* cycles) {
  const int t = threadIdx.x;
  Fq a = io[t & 31], b = io[32 + (t & 31)];
  G1XYZZ A, B;
  A.X = a; A.Y = b; A.ZZ = io[64 + (t & 3)]; A.ZZZ = io[68 + (t & 3)];
  B.X = b; B.Y = a; B.ZZ = io[72 + (t & 3)]; B.ZZZ = io[76 + (t & 3)];
  if (OP >= 5) {  // replicated operands inside lane groups
    int g = (OP == 5) ? (t & 30) : (t & 28);
    A.X = io[g & 31]; A.Y = io[32 + (g & 31)]; B.X = A.Y; B.Y = A.X;
    A.ZZ = io[64]; A.ZZZ = io[68]; B.ZZ = io[72]; B.ZZZ = io[76];
  }
  const bool role = t & 1;
  const uint32_t pmask = 3u << (t & 30);
  const uint32_t q = quad_role();
  __syncthreads();
  long long c0 = clock64();
  for (int i = 0; i < n; i++) {
    if (OP == 0) a = fp_mul(a, b);
    if (OP == 1) a = fp_sqr(a);
    if (OP == 2) a = fp_inv(fp_add(a, b));
    if (OP == 3) xyzz_add(A, B);
    if (OP == 4) A = xyzz_dbl(A);
    if (OP == 5) xyzz_add_pair(A, B, role, pmask);
    if (OP == 6) xyzz_add_quad(A, B, q);
    if (OP == 7) A = xyzz_dbl_quad(A, q);
    if (OP == 8) a = fp_sub(a, b);
    if (OP == 9) { a = fp_mul(a, b); b = fp_mul(b, A.ZZ); }  // two independent chains
  }
  long long c1 = clock64();
  __syncthreads();
  if (OP >= 3 && OP <= 7) a = fp_add(fp_add(A.X, A.Y), fp_add(A.ZZ, A.ZZZ));
  if (OP == 9) a = fp_add(a, b);
  io[128 + t] = a;
  if (t == 0) *cycles = c1 - c0;
}

template <int OP>
void run(const char* name, int n, Fq* d, long long* dc) {
  printf("%-16s", name);
  for (int warps : {1, 4, 8, 16, 32}) {
    chain<OP><<<1, 32 * warps>>>(d, n, dc);
    chain<OP><<<1, 32 * warps>>>(d, n, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("  w=%2d: %8.0f", warps, (double)c / n);
  }
  cudaError_t e = cudaGetLastError();
  printf("   cycles/op %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  Fq h[128 + 1024];
  uint32_t s = 12345;
  for (auto& f : h) { for (int i = 0; i < 8; i++) { s = s * 1664525u + 1013904223u; f.v[i] = s; } f.v[7] &= 0x1fffffffu; }
  Fq* d; long long* dc;
  cudaMalloc(&d, sizeof h); cudaMalloc(&dc, 8);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  run<0>("fp_mul", 256, d, dc);
  run<9>("fp_mul x2 ilp", 256, d, dc);
  run<1>("fp_sqr", 256, d, dc);
  run<8>("fp_sub", 256, d, dc);
  run<2>("fp_inv", 8, d, dc);
  run<3>("xyzz_add", 32, d, dc);
  run<4>("xyzz_dbl", 32, d, dc);
  run<5>("xyzz_add_pair", 32, d, dc);
  run<6>("xyzz_add_quad", 32, d, dc);
  run<7>("xyzz_dbl_quad", 32, d, dc);
  return 0;
}
