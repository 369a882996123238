#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// column-wise (product-scanning) order: the accumulator is the same for consecutive instructions, multiplicands change
__global__ void k_col(uint64_t* out, int iters){
  uint32_t a[8], b[8]; uint64_t c[15];
  for(int i=0;i<8;i++){ a[i]=threadIdx.x*2654435761u+i*40503u; b[i]=blockIdx.x*2246822519u+i*3266489917u; }
  for(int k=0;k<15;k++) c[k]=k;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int k=0;k<15;k++)
      #pragma unroll
      for(int i=0;i<8;i++){ int j=k-i; if(j<0||j>7) continue; asm volatile("mad.wide.u32 %0,%1,%2,%0;":"+l"(c[k]):"r"(a[i]),"r"(b[j])); }
    #pragma unroll
    for(int i=0;i<8;i++){ a[i]^=(uint32_t)c[i]; b[i]+=(uint32_t)(c[i+7]>>32); }
  }
  uint64_t s=0; for(int k=0;k<15;k++) s^=c[k];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// row-wise order (a[i] reused, b[j] and accumulator change) -- as before
__global__ void k_row(uint64_t* out, int iters){
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
// multiply by immediates (modulus-like), accumulator changes, one register multiplicand reused
__global__ void k_imm(uint64_t* out, int iters){
  uint32_t m=threadIdx.x*2654435761u; uint64_t c[8];
  for(int k=0;k<8;k++) c[k]=k;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int r=0;r<8;r++){
      asm volatile("mad.wide.u32 %0,%1,0x3c208c16,%0;":"+l"(c[0]):"r"(m)); asm volatile("mad.wide.u32 %0,%1,0x6871ca8d,%0;":"+l"(c[1]):"r"(m));
      asm volatile("mad.wide.u32 %0,%1,0x97816a91,%0;":"+l"(c[2]):"r"(m)); asm volatile("mad.wide.u32 %0,%1,0x8181585d,%0;":"+l"(c[3]):"r"(m));
      asm volatile("mad.wide.u32 %0,%1,0xb85045b6,%0;":"+l"(c[4]):"r"(m)); asm volatile("mad.wide.u32 %0,%1,0xe131a029,%0;":"+l"(c[5]):"r"(m));
      asm volatile("mad.wide.u32 %0,%1,0x30644e72,%0;":"+l"(c[6]):"r"(m)); asm volatile("mad.wide.u32 %0,%1,0xd87cfd47,%0;":"+l"(c[7]):"r"(m));
      m^=(uint32_t)c[r];
    }
  }
  uint64_t s=0; for(int k=0;k<8;k++) s^=c[k];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<class F> double run(F f, double ops){ cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e9;
  for(int r=0;r<3;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return ops/best*1e-6; }
int main(){ uint64_t* out; cudaMalloc(&out,148*8*256*8); int iters=512;
  for(int bps: {2,8}){ int blocks=148*bps; double n=(double)blocks*256*iters*64;
    printf("warps/SM=%d  column-order %.0f | row-order %.0f | imm %.0f  G lane-ops/s\n", bps*8, run([&]{k_col<<<blocks,256>>>(out,iters);},n), run([&]{k_row<<<blocks,256>>>(out,iters);},n), run([&]{k_imm<<<blocks,256>>>(out,iters);},n)); }
  return 0; }
