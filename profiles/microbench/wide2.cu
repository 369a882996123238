#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// 8x8 schoolbook pattern of IMAD.WIDE.U32 with distinct register operands and 64-bit accumulators (no carries)
__global__ void k_sb(uint64_t* out, int iters){
  uint32_t a[8], b[8]; uint64_t c[15];
  for(int i=0;i<8;i++){ a[i]=threadIdx.x*2654435761u+i*40503u; b[i]=blockIdx.x*2246822519u+i*3266489917u; }
  for(int k=0;k<15;k++) c[k]=k;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++)
      #pragma unroll
      for(int j=0;j<8;j++) asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(c[i+j]):"r"(a[i]),"r"(b[j]));
    #pragma unroll
    for(int i=0;i<8;i++){ a[i]^=(uint32_t)c[i]; b[i]+=(uint32_t)(c[i+7]>>32); }
  }
  uint64_t s=0; for(int k=0;k<15;k++) s^=c[k];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// same with only 2 distinct multiplicand registers (reuse-friendly)
__global__ void k_reuse(uint64_t* out, int iters){
  uint32_t a=threadIdx.x*2654435761u, b=blockIdx.x*2246822519u+1; uint64_t c[15];
  for(int k=0;k<15;k++) c[k]=k;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++)
      #pragma unroll
      for(int j=0;j<8;j++) asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(c[i+j]):"r"(a),"r"(b));
    a^=(uint32_t)c[3]; b+=(uint32_t)(c[9]>>32);
  }
  uint64_t s=0; for(int k=0;k<15;k++) s^=c[k];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> double run(F f, double ops){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e9;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return ops/best*1e-6; }
int main(){ uint64_t* out; cudaMalloc(&out,148*8*256*8); int iters=512;
  for(int bps: {2,4,8}){ int blocks=148*bps; double n=(double)blocks*256*iters*64;
    printf("warps/SM=%d schoolbook distinct regs: %.0f G lane-ops/s ; 2 regs reused: %.0f G lane-ops/s\n", bps*8, run([&]{k_sb<<<blocks,256>>>(out,iters);},n), run([&]{k_reuse<<<blocks,256>>>(out,iters);},n)); }
  return 0; }
