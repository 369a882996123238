// Issue rate of the integer multiply forms a Montgomery product can be built from (one SM sub-partition view):
// cycles per warp instruction with 8 independent chains per thread and 1..16 warps per CTA on one SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int OP>
__global__ void k(uint32_t* out, int iters, long long* cycles) {
  uint32_t a[8], b[8];
  uint64_t c[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 2654435761u + i * 40503u + 1; b[i] = threadIdx.x * 2246822519u + i * 3266489917u + 7; c[i] = a[i] ^ b[i]; }
  uint32_t lo[8], hi[8];
  for (int i = 0; i < 8; i++) { lo[i] = (uint32_t)c[i]; hi[i] = (uint32_t)(c[i] >> 32); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (OP == 0) {  // IMAD.WIDE.U32 reg form, 64-bit accumulate (distinct multiplicands per chain)
#define X(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a[i]), "r"(b[i]));
      REP8(X)
#undef X
      // multiplicands change every round (otherwise ptxas hoists the products and the loop is IADD3 only)
#define Y(i) a[i] = (a[i] << 1) | (a[i] >> 31);
      REP8(Y)
#undef Y
    }
    if (OP == 1) {  // IMAD.WIDE.U32 immediate form
#define X(i) asm volatile("mad.wide.u32 %0, %1, 0x3c208c16, %0;" : "+l"(c[i]) : "r"(a[i]));
      REP8(X)
#undef X
#define Y(i) a[i] = (a[i] << 1) | (a[i] >> 31);
      REP8(Y)
#undef Y
    }
    if (OP == 2) {  // IMAD (lo) reg form
#define X(i) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a[i]), "r"(b[i]));
      REP8(X)
#undef X
    }
    if (OP == 3) {  // IMAD (lo) immediate form
#define X(i) asm volatile("mad.lo.u32 %0, %1, 0x3c208c16, %0;" : "+r"(lo[i]) : "r"(a[i]));
      REP8(X)
#undef X
    }
    if (OP == 4) {  // IMAD.HI reg form
#define X(i) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a[i]), "r"(b[i]));
      REP8(X)
#undef X
    }
    if (OP == 5) {  // IMAD.HI immediate form
#define X(i) asm volatile("mad.hi.u32 %0, %1, 0x3c208c16, %0;" : "+r"(hi[i]) : "r"(a[i]));
      REP8(X)
#undef X
    }
    if (OP == 6) {  // wide with carry chain (the form fp_mul uses): 4-long chains, reg form
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                   "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                   : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]));
      asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                   "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                   : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                   : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[1]));
    }
    if (OP == 7) {  // the same with immediate multiplicands (reduction rows)
      asm volatile("mad.lo.cc.u32 %0, %8, 0x3c208c16, %0; madc.hi.cc.u32 %1, %8, 0x3c208c16, %1; madc.lo.cc.u32 %2, %8, 0x97816a91, %2; madc.hi.cc.u32 %3, %8, 0x97816a91, %3;"
                   "madc.lo.cc.u32 %4, %8, 0xb85045b6, %4; madc.hi.cc.u32 %5, %8, 0xb85045b6, %5; madc.lo.cc.u32 %6, %8, 0x30644e72, %6; madc.hi.u32 %7, %8, 0x30644e72, %7;"
                   : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3])
                   : "r"(a[0]));
      asm volatile("mad.lo.cc.u32 %0, %8, 0xd87cfd47, %0; madc.hi.cc.u32 %1, %8, 0xd87cfd47, %1; madc.lo.cc.u32 %2, %8, 0x6871ca8d, %2; madc.hi.cc.u32 %3, %8, 0x6871ca8d, %3;"
                   "madc.lo.cc.u32 %4, %8, 0x8181585d, %4; madc.hi.cc.u32 %5, %8, 0x8181585d, %5; madc.lo.cc.u32 %6, %8, 0xe131a029, %6; madc.hi.u32 %7, %8, 0xe131a029, %7;"
                   : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7])
                   : "r"(a[1]));
    }
    if (OP == 8) {  // IADD3 with carry chain (alu pipe)
      asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11; addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %10; addc.u32 %7, %7, %11;"
                   : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]));
    }
    if (OP == 9) {  // mix: 8 IMAD imm (lo) + 8 IADD3 carry adds, independent
#define X(i) asm volatile("mad.lo.u32 %0, %1, 0x3c208c16, %0;" : "+r"(hi[i]) : "r"(a[i]));
      REP8(X)
#undef X
      asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %10; addc.cc.u32 %3, %3, %11; addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %10; addc.u32 %7, %7, %11;"
                   : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
                   : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s ^= lo[i] ^ hi[i] ^ (uint32_t)c[i] ^ (uint32_t)(c[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, uint32_t* d, long long* dc) {
  printf("%-44s", name);
  const int iters = 2048;
  for (int warps : {4, 8, 16, 32}) {
    k<OP><<<1, 32 * warps>>>(d, iters, dc);
    k<OP><<<1, 32 * warps>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    // cycles per warp instruction per sub-partition: warps / 4 warps share a scheduler
    printf("  w/SMSP=%d: %5.2f", warps / 4, (double)c / iters / per_iter / (warps / 4.0));
  }
  printf("   cycles per warp-instruction per scheduler\n");
}

int main() {
  uint32_t* d; long long* dc;
  cudaMalloc(&d, 1 << 20); cudaMalloc(&dc, 8);
  run<0>("IMAD.WIDE reg, 64-bit accumulate (+1 rotate)", 8, d, dc);
  run<1>("IMAD.WIDE imm, 64-bit accumulate (+1 rotate)", 8, d, dc);
  run<2>("IMAD lo reg", 8, d, dc);
  run<3>("IMAD lo imm", 8, d, dc);
  run<4>("IMAD.HI reg", 8, d, dc);
  run<5>("IMAD.HI imm", 8, d, dc);
  run<6>("IMAD.WIDE.X carry chains reg (fp_mul rows)", 8, d, dc);
  run<7>("IMAD.WIDE.X carry chains imm (reduce rows)", 8, d, dc);
  run<8>("IADD3.X carry chain", 8, d, dc);
  run<9>("8 IMAD imm + 8 IADD3.X (both pipes)", 16, d, dc);
  return 0;
}
